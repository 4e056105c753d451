"""Flat-bucket AdamW for the B200 path.

`FlatAdamW` is a torch.optim.Optimizer (so `transformers.get_cosine_schedule_with_warmup` and any LR scheduler
work on it unchanged) whose step is two kernel launches over ONE contiguous fp32 bucket per parameter group:
osb_grad_sumsq (global norm + non-finite detection) and osb_adamw_step (unscale by the static loss scale, clip
by global norm, decoupled-weight-decay Adam).  Under data parallelism the same bucket is what gets all-reduced
(one NCCL call per optimizer per step).

Bucket membership is static per training phase: parameters that never receive a gradient (the decoder and the
energy embedding at reference commit 3bdde20 — generator/__init__.py:161 feeds the vocoder `segment.detach()`)
are left out, exactly as torch.optim.AdamW skips parameters whose .grad is None (no weight decay on them either).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import _lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Bucket:
    def __init__(self, params: List[torch.nn.Parameter]):
        dev = params[0].device
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4          # keep every segment 16-byte aligned
        self.params = params
        self.ids = {id(p) for p in params}
        self.offsets = offs
        self.offset_of = {id(p): o for p, o in zip(params, offs)}
        self.numel = total
        self.flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        self.m = torch.zeros(total, device=dev, dtype=torch.float32)
        self.v = torch.zeros(total, device=dev, dtype=torch.float32)
        self.stats = torch.zeros(2, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(params, offs):
                n = p.numel()
                self.flat_p[o:o + n].copy_(p.detach().reshape(-1))
                if p.grad is not None:
                    self.flat_g[o:o + n].copy_(p.grad.detach().reshape(-1))
                p.data = self.flat_p[o:o + n].view(p.shape)
                p.grad = self.flat_g[o:o + n].view(p.shape)

    def view_of(self, flat: torch.Tensor, p: torch.nn.Parameter) -> torch.Tensor:
        o = self.offset_of[id(p)]
        return flat[o:o + p.numel()].view(p.shape)


class FlatAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm: float = 0.0,
                 loss_scale: float = 1.0, process_group=None, world_size: int = 1):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.max_grad_norm = float(max_grad_norm)
        self.loss_scale = float(loss_scale)
        self.process_group = process_group
        self.world_size = int(world_size)
        self._buckets: Dict[int, _Bucket] = {}
        self._steps: Dict[int, int] = {}
        self.last_stats: Optional[torch.Tensor] = None
        # CUDA-graph mode (model/graphed.py): lr and the bias corrections are read from `hyper_dev`, which the host refreshes
        # with stage_hyper() before every replay; step() then has no per-step host value baked into its launches.
        self.graph_mode = False
        self.hyper_host: Optional[torch.Tensor] = None
        self.hyper_dev: Optional[torch.Tensor] = None

    # -- bucket management -----------------------------------------------------------------------
    def _bucket_for(self, gi: int, group) -> Optional[_Bucket]:
        live = [p for p in group["params"] if p.grad is not None]
        if not live:
            return None
        b = self._buckets.get(gi)
        if b is not None and b.ids == {id(p) for p in live}:
            return b
        # (re)build, carrying Adam moments over for parameters that were already tracked
        old = b
        saved = {}
        if old is not None:
            for p in old.params:
                saved[id(p)] = (old.view_of(old.m, p).clone(), old.view_of(old.v, p).clone())
        b = _Bucket(live)
        for p in live:
            if id(p) in saved:
                b.view_of(b.m, p).copy_(saved[id(p)][0])
                b.view_of(b.v, p).copy_(saved[id(p)][1])
        self._buckets[gi] = b
        return b

    def buckets(self) -> List[_Bucket]:
        return list(self._buckets.values())

    def zero_grad(self, set_to_none: bool = False):
        for gi, group in enumerate(self.param_groups):
            b = self._buckets.get(gi)
            if b is not None:
                b.flat_g.zero_()
            for p in group["params"]:
                if b is None or id(p) not in b.ids:
                    p.grad = None

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm: Optional[float] = None):
        max_norm = self.max_grad_norm if max_grad_norm is None else float(max_grad_norm)
        for gi, group in enumerate(self.param_groups):
            b = self._bucket_for(gi, group)
            if b is None:
                continue
            if self.world_size > 1:
                torch.distributed.all_reduce(b.flat_g, group=self.process_group)
            # the all-reduce SUMS the ranks' gradients; the 1/world factor is folded into the unscale factor
            inv_scale = 1.0 / (self.loss_scale * self.world_size)
            if self.graph_mode:
                self._kernel_step_dev(b, group, gi, max_norm, inv_scale)
                self.last_stats = b.stats
                continue  # step counters and weight epochs are advanced by stage_hyper(), once per replay
            step = self._steps.get(gi, 0) + 1
            self._steps[gi] = step
            self._kernel_step(b, group, step, max_norm, inv_scale)
            self.last_stats = b.stats
            for p in b.params:  # invalidate packed fp16 copies of these weights (model/packing.py)
                p._osb_epoch = getattr(p, "_osb_epoch", 0) + 1
        return None

    def _kernel_step(self, b: _Bucket, group, step: int, max_norm: float, inv_scale: float) -> None:
        """Two launches over the flat bucket: global norm (+ non-finite flag), then unscale + clip + AdamW."""
        lib = _lib.load()
        b.stats.zero_()
        _lib.check(lib.osb_grad_sumsq(b.flat_g.data_ptr(), b.numel, b.stats.data_ptr(), _stream()), "osb_grad_sumsq")
        beta1, beta2 = group["betas"]
        _lib.check(lib.osb_adamw_step(b.flat_p.data_ptr(), b.flat_g.data_ptr(), b.m.data_ptr(), b.v.data_ptr(), b.numel,
                                      b.stats.data_ptr(), float(group["lr"]), float(beta1), float(beta2), float(group["eps"]),
                                      float(group["weight_decay"]), step, max_norm, inv_scale, _stream()), "osb_adamw_step")

    def _ensure_hyper(self, device) -> None:
        if self.hyper_dev is None:
            n = len(self.param_groups)
            self.hyper_host = torch.zeros(n, 4, dtype=torch.float32).pin_memory()
            self.hyper_dev = torch.zeros(n, 4, dtype=torch.float32, device=device)

    def stage_hyper(self) -> None:
        """Host side of one graph-mode step: advance the step counters, stage [lr, 1-b1^t, sqrt(1-b2^t)] of every group
        for the captured osb_adamw_step_dev launches (async copy on the current stream) and invalidate packed weights."""
        for gi, group in enumerate(self.param_groups):
            b = self._buckets.get(gi)
            if b is None:
                continue
            self._ensure_hyper(b.flat_p.device)
            step = self._steps.get(gi, 0) + 1
            self._steps[gi] = step
            beta1, beta2 = group["betas"]
            self.hyper_host[gi, 0] = float(group["lr"])
            self.hyper_host[gi, 1] = 1.0 - beta1 ** step
            self.hyper_host[gi, 2] = (1.0 - beta2 ** step) ** 0.5
            for p in b.params:
                p._osb_epoch = getattr(p, "_osb_epoch", 0) + 1
        if self.hyper_dev is not None:
            self.hyper_dev.copy_(self.hyper_host, non_blocking=True)

    def _kernel_step_dev(self, b: _Bucket, group, gi: int, max_norm: float, inv_scale: float) -> None:
        lib = _lib.load()
        self._ensure_hyper(b.flat_p.device)
        b.stats.zero_()
        _lib.check(lib.osb_grad_sumsq(b.flat_g.data_ptr(), b.numel, b.stats.data_ptr(), _stream()), "osb_grad_sumsq")
        beta1, beta2 = group["betas"]
        _lib.check(lib.osb_adamw_step_dev(b.flat_p.data_ptr(), b.flat_g.data_ptr(), b.m.data_ptr(), b.v.data_ptr(), b.numel,
                                          b.stats.data_ptr(), self.hyper_dev[gi].data_ptr(), float(beta1), float(beta2),
                                          float(group["eps"]), float(group["weight_decay"]), max_norm, inv_scale, _stream()),
                   "osb_adamw_step_dev")

    def grad_norm(self) -> float:
        """Unscaled global gradient norm of the last step (reads a device scalar: call outside the hot loop)."""
        if self.last_stats is None:
            return float("nan")
        s = self.last_stats.tolist()
        return float("inf") if s[1] else (s[0] ** 0.5) / (self.loss_scale * self.world_size)
