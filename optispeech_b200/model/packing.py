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


# --------------------------------------------------------------------------------------------------
# per-step weight packs of the training path, batched into one launch
# --------------------------------------------------------------------------------------------------
class StepPacker:
    """Trainable weights change every step, so their fp16 tensor-core copies are per-step work: ~44 small pack launches in
    the ConvNeXt pre-training step.  The packer records the sequence of pack requests of one training forward (source
    parameter, layout, column scale), then serves the same sequence on later steps from ONE `osb_pack_multi` launch at the
    start of the forward: the job table (device) and the destination buffers are built once, the sources are parameter
    storage (stable addresses: views into the optimizer's flat bucket).  Any deviation from the recorded sequence (another
    training phase, other shapes, re-allocated parameters) drops the plan; the requests of that step fall back to one
    launch each and the next step records again.  Requests are served in order, so the plan is also what a captured CUDA
    graph replays."""

    def __init__(self):
        self.plan = None          # dict(specs, outs, table, total)
        self.recording = None     # list of (spec, out) while recording
        self.cursor = 0
        self.active = False

    @staticmethod
    def _spec(kind, srcs, col_scale, k_pad, aux=None):
        return (kind, tuple((s.data_ptr(), tuple(s.shape)) for s in srcs), col_scale.data_ptr() if col_scale is not None else 0, k_pad or 0,
                aux.data_ptr() if aux is not None else 0)

    def begin(self, device):
        from .. import _lib, ops

        self.cursor = 0
        if self.plan is not None:
            if self.plan["device"] != device:
                self.plan = None
            else:
                _lib.check(_lib.load().osb_pack_multi(self.plan["table"].data_ptr(), self.plan["n_jobs"], self.plan["total"], ops._stream()),
                           "osb_pack_multi")
                self.active = True
                self.recording = None
                return
        self.active = False
        self.recording = []

    def end(self):
        if self.recording:
            self._build(self.recording)
        self.recording = None
        self.active = False

    def drop(self):
        self.plan, self.recording, self.active = None, None, False

    def request(self, kind, srcs, col_scale, k_pad, direct, aux=None):
        """-> packed tensor for (kind, srcs, col_scale, k_pad); `direct()` packs it with one launch (fallback / recording).
        kind "matvec": fp32 (rows,) = aux + srcs[0] @ col_scale."""
        if self.active:
            spec = self._spec(kind, srcs, col_scale, k_pad, aux)
            if self.cursor < len(self.plan["specs"]) and self.plan["specs"][self.cursor] == spec:
                out = self.plan["outs"][self.cursor]
                self.cursor += 1
                return out
            self.drop()                       # the step deviates from the recorded one
            return direct()
        out = direct()
        if self.recording is not None:
            self.recording.append((self._spec(kind, srcs, col_scale, k_pad, aux), kind, list(srcs), col_scale, k_pad, out, aux))
        return out

    def _build(self, rec):
        import ctypes as C

        from .. import _lib

        jobs, outs, specs, first = [], [], [], 0
        for spec, kind, srcs, col_scale, k_pad, sample, aux in rec:
            out = torch.empty_like(sample)
            flat = out.view(-1)
            off = 0
            if kind == "matvec":
                rows, cols = srcs[0].shape
                j = _lib.PackJob()
                j.src, j.col_scale, j.aux = srcs[0].data_ptr(), col_scale.data_ptr(), (aux.data_ptr() if aux is not None else None)
                j.dst, j.first_elem = flat.data_ptr(), first
                j.kind, j.rows, j.cols, j.k, j.dst_cols = 2, rows, cols, 1, 8
                jobs.append(j)
                first += 8 * rows
                outs.append(out)
                specs.append(spec)
                continue
            for s in srcs:
                j = _lib.PackJob()
                j.src, j.col_scale = s.data_ptr(), (col_scale.data_ptr() if col_scale is not None else None)
                j.dst = flat.data_ptr() + 2 * off
                j.first_elem = first
                if kind == "nk":
                    rows, cols = s.shape
                    j.kind, j.rows, j.cols, j.k, j.dst_cols = 0, rows, cols, 1, (k_pad or cols)
                    n = rows * j.dst_cols
                else:
                    N, Cin, k = s.shape
                    j.kind, j.rows, j.cols, j.k, j.dst_cols = 1, N, Cin, k, (k_pad or Cin)
                    n = k * N * j.dst_cols
                if j.dst_cols % 8 or n % 8:   # the launch writes 16-byte groups: keep such a step on the per-request path
                    self.plan = None
                    return
                jobs.append(j)
                first += n
                off += n
            assert off == flat.numel(), (off, flat.numel(), kind)
            outs.append(out)
            specs.append(spec)
        arr = (_lib.PackJob * len(jobs))(*jobs)
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        dev = rec[0][5].device
        self.plan = dict(specs=specs, outs=outs, table=raw.to(dev), n_jobs=len(jobs), total=first, device=dev,
                         keepalive=[(srcs, cs, ax) for _, _, srcs, cs, _, _, ax in rec])


_CURRENT: "StepPacker | None" = None   # the packer of the training forward in progress (set by generator_training_forward)


def current_packer() -> "StepPacker | None":
    return _CURRENT


class step_packs:
    """Context of one training forward: `with step_packs(generator, device): ...` routes the pack requests of the autograd
    Functions to the generator's own StepPacker (one plan per model: two models in one process do not disturb each other)."""

    def __init__(self, owner, device):
        self.packer = owner.__dict__.setdefault("_step_packer", StepPacker())
        self.device = device

    def __enter__(self):
        global _CURRENT
        self.prev, _CURRENT = _CURRENT, self.packer
        self.packer.begin(self.device)
        return self.packer

    def __exit__(self, *exc):
        global _CURRENT
        self.packer.end()
        _CURRENT = self.prev
