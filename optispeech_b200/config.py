"""A small Hydra-compatible composer / instantiator for the `configs/model` group.

Hydra and OmegaConf are not dependencies of this implementation; the subset the reference's model configs use is
re-implemented here: `defaults` lists (with `_self_`, nested groups, same-group includes and `override group: option`),
`${a.b}` interpolation against the composed root, and `_target_` / `_partial_` instantiation
(reference: configs/model/**.yaml, optispeech/train.py:54-60 `hydra.utils.instantiate(cfg.model)`).
"""
from __future__ import annotations

import functools
import importlib
import os
import re
from typing import Any, Dict, Optional

import yaml

_YAML_FLOAT = re.compile(r"^[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)?$")


class AttrDict(dict):
    """dict with attribute access (what OmegaConf's DictConfig gives the reference code)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _load(path: str) -> dict:
    with open(path) as f:
        data = yaml.safe_load(f) or {}

    def fix(v):  # PyYAML reads "2e-4" / "1e-2" as strings (YAML 1.1); OmegaConf reads them as floats
        if isinstance(v, str) and _YAML_FLOAT.match(v):
            return float(v)
        if isinstance(v, dict):
            return {k: fix(x) for k, x in v.items()}
        if isinstance(v, list):
            return [fix(x) for x in v]
        return v

    return fix(data)


def _merge(dst: dict, src: dict) -> dict:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v
    return dst


def _set_path(d: dict, path: str, value):
    keys = path.split("/")
    for k in keys[:-1]:
        d = d.setdefault(k, {})
    d[keys[-1]] = value


def compose_group(root: str, group: str, name: str, overrides: Optional[Dict[str, str]] = None) -> dict:
    """Compose `<root>/<group>/<name>.yaml`.  `overrides` maps a group path relative to `group` (e.g. "generator/encoder")
    to the option to use instead of the one named in the defaults list."""
    overrides = dict(overrides or {})
    cfg = _load(os.path.join(root, group, name + ".yaml"))
    defaults = cfg.pop("defaults", ["_self_"])
    # `override x: y` entries of this file apply to everything it includes
    for entry in defaults:
        if isinstance(entry, dict):
            for k, v in entry.items():
                if k.startswith("override "):
                    overrides.setdefault(k[len("override "):].strip(), v)
    out: dict = {}
    for entry in defaults:
        if entry == "_self_":
            _merge(out, cfg)
        elif isinstance(entry, str):  # same-group include
            _merge(out, compose_group(root, group, entry, overrides))
        else:
            for k, v in entry.items():
                if k.startswith("override "):
                    continue
                option = overrides.get(k, v)
                sub_over = {p[len(k) + 1:]: o for p, o in overrides.items() if p.startswith(k + "/")}
                sub = compose_group(root, os.path.join(group, k), option, sub_over)
                tmp: dict = {}
                _set_path(tmp, k, sub)
                _merge(out, tmp)
    if "_self_" not in defaults:
        _merge(out, cfg)
    return out


_INTERP = re.compile(r"^\$\{([^}]+)\}$")


def resolve(node: Any, root: dict) -> Any:
    """Replace whole-value `${a.b.c}` interpolations with the referenced node of `root`."""
    if isinstance(node, dict):
        return {k: resolve(v, root) for k, v in node.items()}
    if isinstance(node, list):
        return [resolve(v, root) for v in node]
    if isinstance(node, str):
        m = _INTERP.match(node)
        if m:
            cur: Any = root
            for key in m.group(1).split("."):
                cur = cur[key]
            return resolve(cur, root)
    return node


_ALIASES = (("optispeech.", "optispeech_b200."),)


def locate(target: str):
    """Import `pkg.mod.attr`; `optispeech.*` targets resolve to this implementation."""
    for old, new in _ALIASES:
        if target.startswith(old):
            target = new + target[len(old):]
    mod_name, _, attr = target.rpartition(".")
    try:
        return getattr(importlib.import_module(mod_name), attr)
    except (ImportError, AttributeError) as e:
        raise ImportError(f"cannot resolve _target_ {target!r}: {e} (modules outside the ConvNeXt hot path are not "
                          f"part of this implementation, see DESIGN.md)") from e


def instantiate(node: Any) -> Any:
    """hydra.utils.instantiate for plain dict configs: `_target_` (+ `_partial_`) nodes become objects / partials,
    other mappings become AttrDicts; already-instantiated objects pass through."""
    if isinstance(node, dict):
        if "_target_" in node:
            kwargs = {k: instantiate(v) for k, v in node.items() if k not in ("_target_", "_partial_")}
            fn = locate(node["_target_"])
            return functools.partial(fn, **kwargs) if node.get("_partial_", False) else fn(**kwargs)
        return AttrDict({k: instantiate(v) for k, v in node.items()})
    if isinstance(node, list):
        return [instantiate(v) for v in node]
    return node


def default_config_root() -> str:
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs")


def compose_model(name: str = "optispeech", data: Optional[dict] = None, config_root: Optional[str] = None,
                  overrides: Optional[Dict[str, str]] = None) -> dict:
    """-> the resolved `model` config dict (what `cfg.model` is in the reference's train.py)."""
    root = config_root or default_config_root()
    model = compose_group(root, "model", name, overrides)
    data = data if data is not None else _load(os.path.join(root, "data", "synthetic.yaml"))
    return resolve(model, {"model": model, "data": data})


def build_from_config(name: str = "optispeech", data: Optional[dict] = None, config_root: Optional[str] = None,
                      overrides: Optional[Dict[str, str]] = None):
    """compose + instantiate: returns the OptiSpeech module exactly as `hydra.utils.instantiate(cfg.model)` would."""
    return instantiate(compose_model(name, data, config_root, overrides))
