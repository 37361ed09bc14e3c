"""Minimal `omegaconf` for running the reference's scripts where the real package cannot be installed (no network).

Implements what santurini/vsrlab's code touches (core/utils.py:13,64; core/loggers.py:25; test.py:80; hydra-style
composition in the sibling `hydra` shim): `DictConfig` / `ListConfig` containers with attribute access and lazy
`${a.b.c}` / `${oc.env:VAR[,default]}` interpolation against the root config, and `OmegaConf.{create, load, merge,
to_container, to_yaml, resolve, select}`.  Not a general replacement."""
from __future__ import annotations

import copy
import json
import os
import re
from typing import Any

import yaml

__all__ = ["OmegaConf", "DictConfig", "ListConfig", "MISSING"]
MISSING = "???"
_INTERP = re.compile(r"\$\{([^{}]+)\}")


class _Loader(yaml.SafeLoader):
    pass


# YAML 1.1 misses floats like 1e-4 (no dot); OmegaConf / YAML 1.2 read them as floats
_Loader.add_implicit_resolver(
    "tag:yaml.org,2002:float",
    re.compile(r"""^(?:[-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?
                |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
                |\.[0-9_]+(?:[eE][-+][0-9]+)?
                |[-+]?\.(?:inf|Inf|INF)|\.(?:nan|NaN|NAN))$""", re.X), list("-+0123456789."))


def _wrap(value: Any, parent, key):
    if isinstance(value, (DictConfig, ListConfig)):
        value = value._plain()
    if isinstance(value, dict):
        return DictConfig(value, parent, key)
    if isinstance(value, (list, tuple)):
        return ListConfig(list(value), parent, key)
    return value


class _Node:
    _parent = None
    _key = None

    def _root(self):
        n = self
        while n._parent is not None:
            n = n._parent
        return n

    def _resolve_value(self, v):
        if isinstance(v, str) and "${" in v:
            return _resolve_str(v, self)
        return v


def _select(root, path: str, node=None):
    """`a.b.0.c` from the root; a leading dot walks up from `node` (relative interpolation)."""
    cur = root
    if path.startswith("."):
        cur = node
        while path.startswith("."):
            path = path[1:]
            if path.startswith("."):
                cur = cur._parent
    for part in [p for p in path.split(".") if p != ""]:
        if isinstance(cur, ListConfig):
            cur = cur[int(part)]
        elif isinstance(cur, DictConfig):
            if part not in cur._data:
                raise KeyError(f"interpolation key '{path}' not found")
            cur = cur[part]
        else:
            raise KeyError(f"interpolation key '{path}' not found")
    return cur


def _resolve_str(s: str, node):
    root = node._root()

    def one(expr: str):
        expr = expr.strip()
        if expr.startswith("oc.env:"):
            name, _, default = expr[len("oc.env:"):].partition(",")
            if name.strip() in os.environ:
                return os.environ[name.strip()]
            if default != "":
                return yaml.load(default.strip(), Loader=_Loader)
            raise KeyError(f"environment variable '{name}' not set")
        return _select(root, expr, node)

    m = _INTERP.fullmatch(s.strip())
    if m:                                        # the whole value is one interpolation: keep the referenced type
        return one(m.group(1))
    prev = None
    while prev != s and "${" in s:
        prev = s
        s = _INTERP.sub(lambda mm: str(one(mm.group(1))), s)
    return s


class DictConfig(_Node):
    def __init__(self, content=None, parent=None, key=None):
        object.__setattr__(self, "_parent", parent)
        object.__setattr__(self, "_key", key)
        object.__setattr__(self, "_data", {})
        for k, v in (content or {}).items():
            self._data[str(k) if not isinstance(k, str) else k] = _wrap(v, self, k)

    # ---- mapping protocol ----
    def __getitem__(self, k):
        v = self._data[k]
        return self._resolve_value(v)

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        try:
            return self[k]
        except KeyError:
            raise AttributeError(f"Missing key {k}") from None

    def __setitem__(self, k, v):
        self._data[k] = _wrap(v, self, k)

    def __setattr__(self, k, v):
        self[k] = v

    def __delitem__(self, k):
        del self._data[k]

    def __contains__(self, k):
        return k in self._data

    def __iter__(self):
        return iter(self._data)

    def __len__(self):
        return len(self._data)

    def keys(self):
        return self._data.keys()

    def values(self):
        return [self[k] for k in self._data]

    def items(self):
        return [(k, self[k]) for k in self._data]

    def get(self, k, default=None):
        return self[k] if k in self._data else default

    def pop(self, k, *default):
        if k in self._data:
            v = self[k]
            del self._data[k]
            return v
        if default:
            return default[0]
        raise KeyError(k)

    def __repr__(self):
        return repr(self._plain())

    def __eq__(self, other):
        return OmegaConf.to_container(self) == (OmegaConf.to_container(other) if isinstance(other, _Node) else other)

    def __deepcopy__(self, memo):
        return DictConfig(copy.deepcopy(self._plain(), memo))

    def __getstate__(self):
        return {"data": OmegaConf.to_container(self, resolve=True)}

    def __setstate__(self, st):
        DictConfig.__init__(self, st["data"])

    def _plain(self, resolve=False):
        out = {}
        for k, v in self._data.items():
            if isinstance(v, _Node):
                out[k] = v._plain(resolve)
            else:
                r = self._resolve_value(v) if resolve else v
                out[k] = r._plain(True) if isinstance(r, _Node) else r
        return out


class ListConfig(_Node):
    def __init__(self, content=None, parent=None, key=None):
        self._parent, self._key = parent, key
        self._data = [_wrap(v, self, i) for i, v in enumerate(content or [])]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._resolve_value(v) for v in self._data[i]]
        return self._resolve_value(self._data[i])

    def __setitem__(self, i, v):
        self._data[i] = _wrap(v, self, i)

    def __iter__(self):
        return (self._resolve_value(v) for v in self._data)

    def __len__(self):
        return len(self._data)

    def __contains__(self, v):
        return v in list(self)

    def append(self, v):
        self._data.append(_wrap(v, self, len(self._data)))

    def __repr__(self):
        return repr(self._plain())

    def __eq__(self, other):
        return self._plain(True) == (other._plain(True) if isinstance(other, ListConfig) else other)

    def __deepcopy__(self, memo):
        return ListConfig(copy.deepcopy(self._plain(), memo))

    def _plain(self, resolve=False):
        out = []
        for v in self._data:
            if isinstance(v, _Node):
                out.append(v._plain(resolve))
            else:
                r = self._resolve_value(v) if resolve else v
                out.append(r._plain(True) if isinstance(r, _Node) else r)
        return out


def _merge_into(dst: dict, src: dict) -> dict:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge_into(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)
    return dst


class OmegaConf:
    @staticmethod
    def create(obj=None):
        if obj is None:
            obj = {}
        if isinstance(obj, str):
            obj = yaml.load(obj, Loader=_Loader)
        return _wrap(obj, None, None)

    @staticmethod
    def load(path):
        text = open(path).read()
        obj = json.loads(text) if str(path).endswith(".json") else yaml.load(text, Loader=_Loader)
        return OmegaConf.create(obj if obj is not None else {})

    @staticmethod
    def merge(*cfgs):
        out: dict = {}
        for c in cfgs:
            _merge_into(out, c._plain() if isinstance(c, _Node) else dict(c))
        return DictConfig(out)

    @staticmethod
    def to_container(cfg, resolve: bool = False, **_ignored):
        return cfg._plain(resolve) if isinstance(cfg, _Node) else cfg

    @staticmethod
    def to_yaml(cfg, resolve: bool = False, sort_keys: bool = False) -> str:
        return yaml.safe_dump(OmegaConf.to_container(cfg, resolve=resolve), default_flow_style=False, sort_keys=sort_keys)

    @staticmethod
    def resolve(cfg) -> None:
        plain = cfg._plain(True)
        if isinstance(cfg, DictConfig):
            DictConfig.__init__(cfg, plain, cfg._parent, cfg._key)
        else:
            ListConfig.__init__(cfg, plain, cfg._parent, cfg._key)

    @staticmethod
    def select(cfg, key: str, default=None):
        try:
            return _select(cfg, key)
        except (KeyError, IndexError, ValueError):
            return default

    @staticmethod
    def is_config(obj) -> bool:
        return isinstance(obj, _Node)

    @staticmethod
    def is_dict(obj) -> bool:
        return isinstance(obj, DictConfig)

    @staticmethod
    def is_list(obj) -> bool:
        return isinstance(obj, ListConfig)
