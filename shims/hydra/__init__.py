"""Minimal `hydra` for running the reference's entry scripts where hydra-core cannot be installed (no network).

Covers what santurini/vsrlab uses: `@hydra.main(config_path, config_name, version_base)` with defaults-list composition
(config groups, nested defaults relative to the including file, `_self_`, `# @package _global_` overlays such as
`+experiment=basic`, their `override /group: option` lines), command-line overrides (`a.b=v`, `+a.b=v`, `~a.b`, `group=option`, `group/sub=option`), the
`hydra.job.env_set` block of conf/hydra/default.yaml, and `hydra.utils.instantiate` (see utils.py).  The composed config
is an `omegaconf.DictConfig` of the sibling shim.  Configs may be `.yaml` or `.json` files."""
from __future__ import annotations

import functools
import json
import os
import re
import sys
from pathlib import Path
from typing import List, Optional, Tuple

import yaml
from omegaconf import DictConfig, OmegaConf
from omegaconf import _Loader, _merge_into  # type: ignore

from . import utils  # noqa: F401

__all__ = ["main", "compose", "utils"]
_PACKAGE = re.compile(r"^\s*#\s*@package\s+(\S+)", re.M)


def _load_file(cfg_dir: Path, rel: str) -> Tuple[dict, Optional[str]]:
    """(content, package directive) of `<cfg_dir>/<rel>.yaml|.json`."""
    for ext in (".yaml", ".yml", ".json"):
        p = cfg_dir / (rel + ext)
        if p.exists():
            text = p.read_text()
            if ext == ".json":
                obj = json.loads(text)
                return obj, obj.pop("__package__", None)
            m = _PACKAGE.search("\n".join(text.splitlines()[:5]))           # `# @package <where>` header
            return (yaml.load(text, Loader=_Loader) or {}), (m.group(1) if m else None)
    raise FileNotFoundError(f"config '{rel}' not found under {cfg_dir} (missing config group file)")


def _place(content: dict, package: str) -> dict:
    if package in ("", "_global_"):
        return content
    out: dict = {}
    cur = out
    parts = package.split(".")
    for p in parts[:-1]:
        cur = cur.setdefault(p, {})
    cur[parts[-1]] = content
    return out


def _compose_file(cfg_dir: Path, rel: str, package: str, group_choice: dict, extra_defaults: List[dict]) -> dict:
    """Merge `rel` and everything its defaults list pulls in; `package` is where this file's own content lands."""
    content, directive = _load_file(cfg_dir, rel)
    if directive is not None:
        package = "" if directive == "_global_" else directive.replace("_group_", package)
    defaults = list(content.pop("defaults", []) or []) + extra_defaults
    if not any(d == "_self_" for d in defaults):
        defaults.append("_self_")
    group_dir = str(Path(rel).parent) if "/" in rel else ""
    out: dict = {}
    for d in defaults:
        if d == "_self_":
            _merge_into(out, _place(content, package))
            continue
        if isinstance(d, str):
            d = {d.rsplit("/", 1)[0]: d.rsplit("/", 1)[1]} if "/" in d else {d: None}
        (grp, opt), = d.items()
        if grp.startswith("override "):
            continue                                      # changes a choice made elsewhere; collected by compose()
        grp = grp.replace("optional ", "").strip()
        absolute = grp.startswith("/")
        grp_path = grp.lstrip("/") if absolute else (f"{group_dir}/{grp}" if group_dir else grp)
        opt = group_choice.get(grp_path, opt)
        if opt is None or opt == "null":
            continue
        child_pkg = grp_path.replace("/", ".")
        _merge_into(out, _compose_file(cfg_dir, f"{grp_path}/{opt}", child_pkg, group_choice, []))
    return out


def _set_path(cfg: dict, key: str, value, must_exist: bool, add: bool):
    cur = cfg
    parts = key.split(".")
    for p in parts[:-1]:
        if p not in cur or not isinstance(cur[p], dict):
            if must_exist:
                raise KeyError(f"override '{key}': no such key (use +{key}=... to add it)")
            cur[p] = {}
        cur = cur[p]
    if must_exist and parts[-1] not in cur:
        raise KeyError(f"override '{key}': no such key (use +{key}=... to add it)")
    if add and parts[-1] in cur and not isinstance(cur[parts[-1]], dict):
        pass
    cur[parts[-1]] = value


def compose(config_path: str, config_name: str, overrides: Optional[List[str]] = None) -> DictConfig:
    cfg_dir = Path(config_path)
    group_choice, extra, value_ops = {}, [], []
    for ov in overrides or []:
        if ov.startswith("~"):
            value_ops.append(("del", ov[1:].split("=")[0], None))
            continue
        key, _, raw = ov.partition("=")
        add = key.startswith("+")
        key = key.lstrip("+")
        is_group = (cfg_dir / key).is_dir()
        if is_group:
            if add:
                extra.append({key: raw})
            else:
                group_choice[key] = raw
        else:
            value_ops.append(("add" if add else "set", key, yaml.load(raw, Loader=_Loader) if raw != "" else ""))
    for e in extra:                                       # `override /group: option` lines of appended overlays (+experiment=...)
        (grp, opt), = e.items()
        content, _ = _load_file(cfg_dir, f"{grp}/{opt}")
        for d in content.get("defaults", []) or []:
            if isinstance(d, dict):
                (k, v), = d.items()
                if k.startswith("override "):
                    group_choice.setdefault(k[len("override "):].strip().lstrip("/"), v)
    merged = _compose_file(cfg_dir, config_name, "", group_choice, extra)
    for op, key, val in value_ops:
        if op == "del":
            cur = merged
            parts = key.split(".")
            for p in parts[:-1]:
                cur = cur[p]
            cur.pop(parts[-1])
        else:
            _set_path(merged, key, val, must_exist=(op == "set"), add=(op == "add"))
    return DictConfig(merged)


def main(config_path: Optional[str] = None, config_name: Optional[str] = None, version_base: Optional[str] = None):
    def deco(fn):
        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            if args and isinstance(args[0], DictConfig):
                return fn(*args, **kwargs)
            cfg = compose(config_path, config_name, [a for a in sys.argv[1:] if not a.startswith("--")])
            hyd = cfg.pop("hydra", None)
            if hyd is not None:
                # interpolations inside the hydra node resolve against the job config
                tmp = OmegaConf.merge(cfg, {"hydra": OmegaConf.to_container(hyd)})
                env = OmegaConf.select(tmp, "hydra.job.env_set")
                for k, v in (OmegaConf.to_container(env, resolve=True) if env is not None else {}).items():
                    os.environ[str(k)] = str(v)
            return fn(cfg)
        return wrapper
    return deco
