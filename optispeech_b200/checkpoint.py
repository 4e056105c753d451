"""Reading checkpoints written by the reference (Lightning + Hydra + OmegaConf) without those packages.

Reference: `OptiSpeech.__init__` calls `save_hyperparameters(logger=False)` (optispeech/model/optispeech.py:26), so a
Lightning checkpoint's `hyper_parameters` hold what Hydra's `instantiate` passed in:

  * `functools.partial`s of `optispeech.model...` classes (`_partial_: true` in configs/model/**) — resolved by the
    repository's `optispeech/` alias package to this implementation (same class paths, same constructor arguments);
  * OmegaConf containers (`train_args`, `data_args`, `inference_args`, nested keyword arguments such as `loss_coeffs`):
    `omegaconf.dictconfig.DictConfig` / `omegaconf.listconfig.ListConfig` whose pickled state is their `__dict__`
    (`_content`: key -> node, `_metadata`, `_parent`, `_flags_cache`) with leaf values wrapped in
    `omegaconf.nodes.*Node` objects (`_val`);
  * live helper objects (`data_args.text_processor`: optispeech.text.TextProcessor, `data_args.feature_extractor`:
    optispeech.dataset.feature_extractors.CommonFeatureExtractor) whose classes (and piper_phonemize / librosa behind
    them) are not part of this package.

`load_checkpoint_file` unpickles such a file with a `find_class` that substitutes attribute-bag placeholders for every
class it cannot import (anything under `omegaconf.`, the text / dataset helpers), and `hyper_parameters_from_checkpoint`
turns the result into plain Python: containers -> `AttrDict` / list, nodes -> their value, helper objects -> attribute
bags (`feature_extractor.sample_rate`, `text_processor.languages` ... keep working; calling a placeholder raises).
Checkpoints written by `OptiSpeech.save_checkpoint` contain none of these and take the same path unchanged.
"""
from __future__ import annotations

import functools
import pickle
from types import SimpleNamespace
from typing import Any, Dict

import torch

_OMEGA_CONTAINERS = {"DictConfig": "dict", "ListConfig": "list"}


class AttrDict(dict):
    """dict with attribute access (what the model code does with train_args / data_args / inference_args)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self[k] = v


class Placeholder:
    """Stand-in for an object whose class is not importable here: keeps the pickled attributes."""

    _osb_module = ""
    _osb_name = ""

    def __init__(self, *args, **kwargs):
        self.__dict__["_osb_args"] = (args, kwargs)

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):   # (dict state, slots state)
            for part in state:
                if isinstance(part, dict):
                    self.__dict__.update(part)
        elif isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["_osb_state"] = state

    # pickled subclasses of dict / list / set arrive through SETITEM(S) / APPEND(S) / ADDITEMS
    def __setitem__(self, k, v):
        self.__dict__.setdefault("_osb_items", {})[k] = v

    def append(self, v):
        self.__dict__.setdefault("_osb_list", []).append(v)

    def extend(self, vs):
        self.__dict__.setdefault("_osb_list", []).extend(vs)

    def add(self, v):
        self.__dict__.setdefault("_osb_list", []).append(v)

    def __call__(self, *a, **k):
        raise RuntimeError(f"{self._osb_module}.{self._osb_name} came from a reference checkpoint and is only an attribute "
                           f"bag here (its implementation is outside this package's scope)")

    def __repr__(self):
        keys = [k for k in self.__dict__ if not k.startswith("_osb")]
        return f"<Placeholder {self._osb_module}.{self._osb_name} {keys}>"


_PLACEHOLDER_TYPES: Dict[tuple, type] = {}


def _placeholder_type(module: str, name: str) -> type:
    key = (module, name)
    t = _PLACEHOLDER_TYPES.get(key)
    if t is None:
        t = type(name, (Placeholder,), {"_osb_module": module, "_osb_name": name})
        _PLACEHOLDER_TYPES[key] = t
    return t


class _ShimUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "omegaconf" or module.startswith("omegaconf."):
            return _placeholder_type(module, name)
        try:
            return super().find_class(module, name)
        except (ImportError, AttributeError):
            return _placeholder_type(module, name)


class _ShimPickleModule:
    """The `pickle_module` interface torch.load expects."""

    __name__ = "optispeech_b200.checkpoint.shim_pickle"
    Unpickler = _ShimUnpickler
    Pickler = pickle.Pickler
    HIGHEST_PROTOCOL = pickle.HIGHEST_PROTOCOL
    dump = staticmethod(pickle.dump)
    dumps = staticmethod(pickle.dumps)

    @staticmethod
    def load(f, **kw):
        return _ShimUnpickler(f, **kw).load()

    @staticmethod
    def loads(b, **kw):
        import io

        return _ShimUnpickler(io.BytesIO(b), **kw).load()


def load_checkpoint_file(path, map_location="cpu") -> Dict[str, Any]:
    """torch.load with the placeholder-substituting unpickler (full pickle: only load checkpoints you trust)."""
    return torch.load(path, map_location=map_location, weights_only=False, pickle_module=_ShimPickleModule)


def to_plain(obj, _memo=None):
    """OmegaConf containers / nodes -> plain Python, recursively (partials are rebuilt around converted arguments)."""
    memo = {} if _memo is None else _memo
    oid = id(obj)
    if oid in memo:
        return memo[oid]
    if isinstance(obj, Placeholder):
        mod, name = obj._osb_module, obj._osb_name
        if mod.startswith("omegaconf"):
            if name in _OMEGA_CONTAINERS:
                content = obj.__dict__.get("_content")
                if _OMEGA_CONTAINERS[name] == "dict":
                    out = AttrDict()
                    memo[oid] = out
                    if isinstance(content, dict):
                        for k, v in content.items():
                            out[to_plain(k, memo)] = to_plain(v, memo)
                    return out
                out = []
                memo[oid] = out
                if isinstance(content, (list, tuple)):
                    out.extend(to_plain(v, memo) for v in content)
                return out
            if "_val" in obj.__dict__:                       # omegaconf.nodes.*Node
                val = to_plain(obj.__dict__["_val"], memo)
                memo[oid] = val
                return val
            memo[oid] = None                                  # metadata objects carry nothing the model needs
            return None
        if "_osb_items" in obj.__dict__ and not [k for k in obj.__dict__ if not k.startswith("_osb")]:   # a dict subclass
            out = AttrDict()
            memo[oid] = out
            for k, v in obj.__dict__["_osb_items"].items():
                out[to_plain(k, memo)] = to_plain(v, memo)
            return out
        memo[oid] = obj                                        # helper object: convert its attributes in place
        for k, v in list(obj.__dict__.items()):
            if not k.startswith("_osb"):
                obj.__dict__[k] = to_plain(v, memo)
        return obj
    if isinstance(obj, functools.partial):
        out = functools.partial(obj.func, *[to_plain(a, memo) for a in obj.args], **{k: to_plain(v, memo) for k, v in obj.keywords.items()})
        memo[oid] = out
        return out
    if isinstance(obj, dict) and not isinstance(obj, AttrDict):
        out = type(obj)() if type(obj) is not dict else {}
        memo[oid] = out
        for k, v in obj.items():
            out[to_plain(k, memo)] = to_plain(v, memo)
        return out
    if isinstance(obj, list):
        out = []
        memo[oid] = out
        out.extend(to_plain(v, memo) for v in obj)
        return out
    if isinstance(obj, tuple):
        return tuple(to_plain(v, memo) for v in obj)
    if isinstance(obj, SimpleNamespace):
        for k, v in list(vars(obj).items()):
            setattr(obj, k, to_plain(v, memo))
        return obj
    return obj


def hyper_parameters_from_checkpoint(ckpt: Dict[str, Any]) -> Dict[str, Any]:
    """`cls(**hparams)` arguments of OptiSpeech from a checkpoint dictionary (reference or ours)."""
    hp = to_plain(ckpt["hyper_parameters"])
    if isinstance(hp, Placeholder):
        hp = {k: v for k, v in hp.__dict__.items() if not k.startswith("_osb")}
    hp = dict(hp)
    for k in ("train_args", "data_args", "inference_args"):
        v = hp.get(k)
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            hp[k] = AttrDict(v)
    keep = ("dim", "generator", "vocoder", "discriminator", "train_args", "data_args", "inference_args", "optimizer", "scheduler")
    return {k: hp[k] for k in keep if k in hp}
