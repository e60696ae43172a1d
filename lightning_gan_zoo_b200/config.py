"""A small stand-in for the slice of Hydra / OmegaConf the reference's HoloGAN experiment uses
(`@hydra.main(config_path="conf", config_name="config")` + `+expt=hologan`, run_network.py:25; `instantiate` of
`_target_` nodes, core/lightning_module.py:38-49,75-87).  hydra / omegaconf are not installed in this image;
where they are, the reference's own loader works unchanged on the same YAML because the key layout is the reference's.

* `Config`: nested mapping with attribute access and `${a.b.c}` interpolation resolved at access time against the root.
* `load_hologan_config()`: the shipped flat document (conf/hologan.yaml) or, with `conf_dir=<reference>/conf`, the
  composition Hydra performs for `+expt=hologan`: config.yaml, its `defaults` groups, the experiment overlay at the root
  (`# @package _global_`) and the experiment's `override /group: option` entries.
* `instantiate(node, *args, **kwargs)`: import `_target_`, call it with the node's remaining keys (nested `_target_`
  nodes are instantiated first), like `hydra.utils.instantiate`.
"""
from __future__ import annotations

import importlib
import os
import re
from typing import Any, Dict, Iterable, Optional

import yaml

_INTERP = re.compile(r"\$\{([^}]+)\}")
CONF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "conf")


class Config:
    """Mapping node with attribute access; string values `${path}` resolve against the root when read."""

    def __init__(self, data: Dict[str, Any], root: Optional["Config"] = None):
        object.__setattr__(self, "_data", data)
        object.__setattr__(self, "_root", root if root is not None else self)

    # ---- access -----------------------------------------------------------------------------------
    def _wrap(self, v, depth=0):
        if depth > 16:
            raise ValueError("interpolation cycle in the configuration")
        if isinstance(v, dict):
            return Config(v, self._root)
        if isinstance(v, list):
            return [self._wrap(x, depth) for x in v]
        if isinstance(v, str):
            m = _INTERP.fullmatch(v.strip())
            if m:                                              # whole-value interpolation keeps the target's type
                return self._wrap(self._root._lookup(m.group(1)), depth + 1)
            if _INTERP.search(v):
                return _INTERP.sub(lambda mm: str(self._wrap(self._root._lookup(mm.group(1)), depth + 1)), v)
        return v

    def _lookup(self, path: str):
        node: Any = self._data
        for part in path.strip().split("."):
            if not isinstance(node, dict) or part not in node:
                raise KeyError(f"interpolation '${{{path}}}' does not resolve (missing '{part}')")
            node = node[part]
        return node

    def __getattr__(self, key):
        try:
            return self._wrap(self._data[key])
        except KeyError:
            raise AttributeError(key) from None

    __getitem__ = __getattr__

    def __setattr__(self, key, value):
        self._data[key] = value._data if isinstance(value, Config) else value

    __setitem__ = __setattr__

    def __contains__(self, key):
        return key in self._data

    def keys(self):
        return self._data.keys()

    def get(self, key, default=None):
        return self._wrap(self._data[key]) if key in self._data else default

    def to_dict(self) -> Dict[str, Any]:
        """Fully resolved plain dict."""
        def res(v):
            v = self._wrap(v) if not isinstance(v, Config) else v
            if isinstance(v, Config):
                return {k: res(v._data[k]) for k in v._data}
            if isinstance(v, list):
                return [res(x) for x in v]
            return v
        return {k: res(self._data[k]) for k in self._data}

    def __repr__(self):
        return f"Config({self._data!r})"


def _merge(dst: Dict[str, Any], src: Dict[str, Any]) -> Dict[str, Any]:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v
    return dst


_FLOAT = re.compile(r"[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)")


def _numbers(v):
    """PyYAML (YAML 1.1) reads `1e-4` as a string; OmegaConf reads it as a float.  Follow OmegaConf."""
    if isinstance(v, dict):
        return {k: _numbers(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_numbers(x) for x in v]
    if isinstance(v, str) and _FLOAT.fullmatch(v.strip()):
        return float(v)
    return v


def _load_yaml(path: str) -> Dict[str, Any]:
    with open(path) as f:
        return _numbers(yaml.safe_load(f) or {})


def _set_path(tree: Dict[str, Any], dotted: str, value):
    parts = dotted.split(".")
    for p in parts[:-1]:
        tree = tree.setdefault(p, {})
    if isinstance(value, dict) and isinstance(tree.get(parts[-1]), dict):
        _merge(tree[parts[-1]], value)
    else:
        tree[parts[-1]] = value


def compose_reference_conf(conf_dir: str, expt: str = "hologan", groups: Optional[Dict[str, str]] = None) -> Dict[str, Any]:
    """What Hydra builds for `python run_network.py +expt=<expt> [group=option ...]` from a reference-style conf tree:
    root config.yaml; every `defaults` group placed under its group name; the experiment file merged at the root; its
    `override /group: option` defaults replacing the group's node.  Groups whose file is missing (e.g. `filepaths:
    local`, which the reference does not ship) are skipped; `hydra/...` entries are ignored.  `/pkg@dest: option`
    entries (the figure callbacks) are placed at `dest`."""
    root = _load_yaml(os.path.join(conf_dir, "config.yaml"))
    chosen: Dict[str, str] = {}
    for d in root.pop("defaults", []) or []:
        if isinstance(d, dict):
            for k, v in d.items():
                if not k.startswith("override hydra") and not k.startswith("hydra"):
                    chosen[k.replace("override ", "").lstrip("/")] = v
    overlay = _load_yaml(os.path.join(conf_dir, "expt", expt + ".yaml"))
    placed = {}
    for d in overlay.pop("defaults", []) or []:
        if isinstance(d, dict):
            for k, v in d.items():
                k = k.strip()
                if k.startswith("override "):
                    chosen[k[len("override "):].lstrip("/")] = v
                elif "@" in k:
                    grp, dest = k.lstrip("/").split("@", 1)
                    placed[dest] = (grp, v)
    chosen.update(groups or {})
    tree: Dict[str, Any] = {}
    for grp, opt in chosen.items():
        path = os.path.join(conf_dir, grp, f"{opt}.yaml")
        if os.path.exists(path):
            tree[grp] = _load_yaml(path)
    _merge(tree, root)
    _merge(tree, overlay)
    for dest, (grp, opt) in placed.items():
        path = os.path.join(conf_dir, grp, f"{opt}.yaml")
        if os.path.exists(path):
            _set_path(tree, dest, _load_yaml(path))
    return tree


def load_hologan_config(conf_dir: Optional[str] = None, overrides: Iterable[str] = (), groups: Optional[Dict[str, str]] = None) -> Config:
    """The HoloGAN experiment configuration.  `conf_dir=None`: the document shipped with this package; else a
    reference-style Hydra conf tree (composed like `+expt=hologan`).  `overrides`: `a.b.c=value` strings (YAML values),
    like Hydra's command-line overrides."""
    tree = _load_yaml(os.path.join(CONF_DIR, "hologan.yaml")) if conf_dir is None else compose_reference_conf(conf_dir, "hologan", groups)
    for ov in overrides:
        key, _, val = ov.partition("=")
        _set_path(tree, key.strip().lstrip("+"), yaml.safe_load(val))
    return Config(tree)


def _locate(dotted: str):
    module, _, attr = dotted.rpartition(".")
    while module:
        try:
            obj = importlib.import_module(module)
            break
        except ModuleNotFoundError:
            module, _, head = module.rpartition(".")
            attr = head + "." + attr
    else:
        raise ImportError(f"cannot locate '{dotted}'")
    for part in attr.split("."):
        obj = getattr(obj, part)
    return obj


def instantiate(node, *args, **kwargs):
    """`hydra.utils.instantiate` for the nodes this experiment has: `_target_` is imported and called with the node's
    other keys as keyword arguments (recursively for nested `_target_` nodes); extra positional / keyword arguments
    are passed through.  Mapping values stay `Config` objects (attribute access, like a DictConfig)."""
    if not isinstance(node, Config) or "_target_" not in node:
        raise ValueError("instantiate() needs a configuration node with a _target_")
    target = _locate(node._target_)
    kw = {}
    for k in node.keys():
        if k == "_target_":
            continue
        v = node[k]
        kw[k] = instantiate(v) if isinstance(v, Config) and "_target_" in v else v
    kw.update(kwargs)
    return target(*args, **kw)
