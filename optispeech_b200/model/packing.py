"""Packed tensor-core operands derived from fp32 master parameters.

The CUDA kernels consume fp16 weights in (2, taps, N, K) layout ([0] = hi = fp16(w), [1] = lo =
fp16(w - hi), the second only read in split-precision mode); nn.Parameters stay fp32 with the
reference's shapes so `state_dict` is interchangeable.  Packed copies are cached per module and
rebuilt when any source parameter changes (tracked through the tensors' in-place version
counters and storage pointers), e.g. after an optimizer step or `load_state_dict`.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, Tuple

import weakref

import torch


_CACHES = weakref.WeakSet()


def live_packs() -> list:
    """Every packed tensor currently cached by any module (a captured CUDA graph keeps them alive: model/graphed.py)."""
    return [hit[1] for cache in list(_CACHES) for hit in list(cache._store.values())]


class PackedCache:
    def __init__(self):
        self._store: Dict[str, Tuple[tuple, object]] = {}
        _CACHES.add(self)

    @staticmethod
    def _key(params: Iterable[torch.Tensor]) -> tuple:
        # _osb_epoch is bumped by FlatAdamW, whose kernels update parameter storage behind autograd's version counter
        return tuple((p.data_ptr(), p._version, p.device.index, getattr(p, "_osb_epoch", 0)) for p in params)

    def get(self, name: str, params: Iterable[torch.Tensor], build: Callable[[], object]):
        params = list(params)
        key = self._key(params)
        hit = self._store.get(name)
        if hit is not None and hit[0] == key:
            return hit[1]
        with torch.no_grad():
            val = build()
        self._store[name] = (key, val)
        return val

    def clear(self):
        self._store.clear()


def pack_linear(weight: torch.Tensor, col_scale: torch.Tensor | None = None, k_pad: int | None = None) -> torch.Tensor:
    """(N, K) fp32 -> (2, 1, N, Kp) fp16 hi/lo; column k optionally scaled by col_scale[k] first."""
    from .. import ops

    N, K = weight.shape
    Kp = k_pad or K
    w = weight.detach().contiguous()
    out = torch.empty((2, 1, N, Kp), device=w.device, dtype=torch.float16)
    ops.pack_h16(w, rows=N, cols=K, src_ld=K, dst_cols=Kp, col_scale=col_scale, out=out[0, 0], out_lo=out[1, 0])
    return out


def pack_conv(weight: torch.Tensor, k_pad: int | None = None) -> torch.Tensor:
    """Conv1d weight (N, Cin, k) fp32 -> (2, k, N, Cin_pad) fp16 hi/lo (one K-major matrix per tap)."""
    from .. import ops

    return ops.pack_conv_h16(weight.detach(), k_pad=k_pad, split=True)
