"""`hydra.utils.instantiate` / `get_class` as the reference calls them (core/utils.py:94-104,129,138,157,180-196):
`_target_` dotted paths, positional args, keyword overrides, `_recursive_` (instantiate nested `_target_` nodes first)
and `_convert_` ("none": containers stay DictConfig/ListConfig, "partial"/"all": plain dict / list)."""
from __future__ import annotations

import importlib
from typing import Any

from omegaconf import DictConfig, ListConfig, OmegaConf

_RESERVED = ("_target_", "_recursive_", "_convert_", "_partial_", "_args_")


def get_class(path: str):
    mod, _, name = path.rpartition(".")
    obj = importlib.import_module(mod)
    return getattr(obj, name)


get_method = get_object = get_class


def _convert(v: Any, mode: str):
    if isinstance(v, (DictConfig, ListConfig)):
        if mode in ("partial", "all", "object"):
            return OmegaConf.to_container(v, resolve=True)
        return OmegaConf.create(OmegaConf.to_container(v, resolve=True))     # detached, resolved copy
    return v


def _build(node: Any, recursive: bool, convert: str):
    if isinstance(node, DictConfig):
        if recursive and "_target_" in node:
            return instantiate(node, _recursive_=True, _convert_=convert)
        if recursive:
            built = {k: _build(node[k], True, convert) for k in node}
            return built if convert in ("partial", "all", "object") else built
        return _convert(node, convert)
    if isinstance(node, ListConfig):
        if recursive:
            return [_build(v, True, convert) for v in node]
        return _convert(node, convert)
    return node


def instantiate(config: Any, *args: Any, **kwargs: Any):
    if config is None:
        return None
    if isinstance(config, dict):
        config = OmegaConf.create(config)
    recursive = kwargs.pop("_recursive_", config.get("_recursive_", True))
    convert = kwargs.pop("_convert_", config.get("_convert_", "none"))
    partial = kwargs.pop("_partial_", config.get("_partial_", False))
    if "_target_" not in config:
        raise ValueError("instantiate: config has no _target_")
    target = config["_target_"]
    fn = get_class(target) if isinstance(target, str) else target
    params = {k: _build(config[k], recursive, convert) for k in config if k not in _RESERVED}
    params.update(kwargs)
    pos = list(config.get("_args_", [])) + list(args)
    if partial:
        import functools
        return functools.partial(fn, *pos, **params)
    return fn(*pos, **params)


call = instantiate
