"""Flat-bucket AdamW for the B200 path.

`FlatAdamW` is a torch.optim.Optimizer (so `transformers.get_cosine_schedule_with_warmup` and any LR scheduler
work on it unchanged) whose step is two kernel launches over ONE contiguous fp32 bucket per parameter group:
osb_grad_gather (multi-tensor copy of the step's gradients into the bucket, fused with the global norm and
non-finite detection) and osb_adamw_step (unscale by the static loss scale, clip by global norm,
decoupled-weight-decay Adam).  Under data parallelism the bucket is what gets all-reduced (one NCCL call per
optimizer per step; the norm is then taken after the all-reduce by osb_grad_sumsq).

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


GATHER_CHUNK = 2048  # floats per work item of osb_grad_gather (csrc/osb_optim.cu)


class _Bucket:
    """`cohorts` = [(n_params, first_step)]: consecutive runs of `params` that joined the bucket together.  `first_step` is the
    group's step count when the run's first update happened minus one, so the run's own Adam step (bias correction) is
    `group_step - first_step` — what torch.optim.AdamW keeps per parameter (parameters that start receiving gradients later,
    e.g. the vocoder after `pretraining_steps`, begin at step 1)."""

    def __init__(self, params: List[torch.nn.Parameter], cohorts=None):
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
        self.cohorts = []   # (element start, element end, first_step)
        cohorts = cohorts or [(len(params), 0)]
        assert sum(n for n, _ in cohorts) == len(params)
        i = 0
        for n, first in cohorts:
            if n == 0:
                continue
            start = offs[i]
            end = offs[i + n] if i + n < len(params) else total
            self.cohorts.append((start, end, int(first)))
            i += n
        self.cohort_of = {}
        i = 0
        for ci, (n, _) in enumerate([c for c in cohorts if c[0] > 0]):
            for p in params[i:i + n]:
                self.cohort_of[id(p)] = ci
            i += n
        self.flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        self.m = torch.zeros(total, device=dev, dtype=torch.float32)
        self.v = torch.zeros(total, device=dev, dtype=torch.float32)
        self.stats = torch.zeros(2, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(params, offs):
                n = p.numel()
                self.flat_p[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[o:o + n].view(p.shape)
        # work list of the gather kernel: (tensor, chunk) pairs; depends on the sizes only
        chunks = [(t, c) for t, p in enumerate(params) for c in range((p.numel() + GATHER_CHUNK - 1) // GATHER_CHUNK)]
        self.n_chunks = len(chunks)
        self.chunks_dev = torch.tensor(chunks, dtype=torch.int32).to(dev)
        # pointer tables [src, dst offset, numel]: a ring of pinned host buffers for eager steps (each guarded by the event of
        # its last upload), and one dedicated, never rewritten (host, device) pair per captured CUDA graph
        self._ring = []
        self._ring_pos = 0
        self._graph_tables = []

    def view_of(self, flat: torch.Tensor, p: torch.nn.Parameter) -> torch.Tensor:
        o = self.offset_of[id(p)]
        return flat[o:o + p.numel()].view(p.shape)

    def _fill_table(self, host: torch.Tensor) -> None:
        rows = []
        for p, o in zip(self.params, self.offsets):
            g = p.grad
            if g is not None and (g.dtype != torch.float32 or not g.is_contiguous()):
                g = g.float().contiguous()
                p.grad = g
            rows.append((g.data_ptr() if g is not None else 0, o, p.numel()))
        host.copy_(torch.tensor(rows, dtype=torch.int64))

    def stage_table(self, for_graph: bool) -> torch.Tensor:
        """Upload this step's gradient pointers; returns the device table the gather launch reads."""
        dev = self.flat_p.device
        n = len(self.params)
        if for_graph:
            host = torch.empty((n, 3), dtype=torch.int64).pin_memory()
            table = torch.empty((n, 3), dtype=torch.int64, device=dev)
            self._fill_table(host)
            table.copy_(host, non_blocking=True)   # captured as a memcpy node: re-reads `host`, which is never rewritten
            self._graph_tables.append((host, table))
            return table
        if len(self._ring) < 4:
            self._ring.append((torch.empty((n, 3), dtype=torch.int64).pin_memory(),
                               torch.empty((n, 3), dtype=torch.int64, device=dev), torch.cuda.Event()))
        host, table, done = self._ring[self._ring_pos % len(self._ring)]
        self._ring_pos += 1
        done.synchronize()                          # the upload that last read this host buffer has finished
        self._fill_table(host)
        table.copy_(host, non_blocking=True)
        done.record()
        return table


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
        self._pending_state: Dict[int, tuple] = {}   # id(param) -> (exp_avg, exp_avg_sq, step) from load_state_dict
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
        live_ids = {id(p) for p in live}
        if b is not None and b.ids == live_ids:
            return b
        # (re)build: parameters already tracked keep their order, run ("cohort") and Adam moments; newcomers form new runs whose
        # bias-correction step starts at 1 (or continues from a loaded state_dict)
        old = b
        gstep = self._steps.get(gi, 0)
        params, cohorts, saved = [], [], {}
        if old is not None:
            for ci, (_, _, first) in enumerate(old.cohorts):
                keep = [p for p in old.params if old.cohort_of[id(p)] == ci and id(p) in live_ids]
                for p in keep:
                    saved[id(p)] = (old.view_of(old.m, p).clone(), old.view_of(old.v, p).clone())
                params += keep
                cohorts.append((len(keep), first))
        new = [p for p in live if id(p) not in saved]
        by_first: Dict[int, list] = {}
        for p in new:
            pend = self._pending_state.pop(id(p), None)
            first = gstep
            if pend is not None:
                saved[id(p)] = (pend[0], pend[1])
                first = gstep - int(pend[2])
            by_first.setdefault(first, []).append(p)
        for first in sorted(by_first):
            params += by_first[first]
            cohorts.append((len(by_first[first]), first))
        b = _Bucket(params, cohorts)
        for p in params:
            if id(p) in saved:
                b.view_of(b.m, p).copy_(saved[id(p)][0].to(b.m.device).reshape(p.shape))
                b.view_of(b.v, p).copy_(saved[id(p)][1].to(b.v.device).reshape(p.shape))
        self._buckets[gi] = b
        return b

    def buckets(self) -> List[_Bucket]:
        return list(self._buckets.values())

    def zero_grad(self, set_to_none: bool = True):
        """Gradients are dropped, not zero-filled: autograd then hands each parameter its gradient tensor as is (no
        accumulate kernel per parameter), and step() gathers them into the flat bucket in one launch."""
        for group in self.param_groups:
            for p in group["params"]:
                p.grad = None

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm: Optional[float] = None):
        max_norm = self.max_grad_norm if max_grad_norm is None else float(max_grad_norm)
        # The device-side hyper-parameter path is for launches that are being CAPTURED (their lr / bias corrections must not be
        # baked into the graph; stage_hyper() refreshes them and advances the counters once per replay).  Every other step —
        # warm-up steps of a new batch shape, accumulation, the phase switch, cuda_graph=False — is a normal eager step even
        # after a capture has happened: it reads lr from the param group, advances the step counter and the weight epochs.
        capturing = self.graph_mode and torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
        for gi, group in enumerate(self.param_groups):
            b = self._bucket_for(gi, group)
            if b is None:
                continue
            b.stats.zero_()
            self._gather_grads(b, fused_norm=self.world_size == 1)
            if self.world_size > 1:
                torch.distributed.all_reduce(b.flat_g, group=self.process_group)
                self._grad_norm(b)
            # the all-reduce SUMS the ranks' gradients; the 1/world factor is folded into the unscale factor
            inv_scale = 1.0 / (self.loss_scale * self.world_size)
            if capturing:
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

    def _gather_grads(self, b: _Bucket, fused_norm: bool) -> None:
        """One launch: every member's .grad -> b.flat_g (+ sum of squares / non-finite flag into b.stats when fused_norm)."""
        lib = _lib.load()
        table = b.stage_table(for_graph=self.graph_mode and torch.cuda.is_current_stream_capturing())
        _lib.check(lib.osb_grad_gather(table.data_ptr(), b.chunks_dev.data_ptr(), b.n_chunks, b.flat_g.data_ptr(),
                                       b.stats.data_ptr() if fused_norm else None, _stream()), "osb_grad_gather")

    def _grad_norm(self, b: _Bucket) -> None:
        _lib.check(_lib.load().osb_grad_sumsq(b.flat_g.data_ptr(), b.numel, b.stats.data_ptr(), _stream()), "osb_grad_sumsq")

    def _kernel_step(self, b: _Bucket, group, step: int, max_norm: float, inv_scale: float) -> None:
        """Unscale + clip by the global norm (b.stats) + AdamW over the flat bucket, one launch per cohort (its own Adam step)."""
        lib = _lib.load()
        beta1, beta2 = group["betas"]
        for start, end, first in b.cohorts:
            o = 4 * start
            _lib.check(lib.osb_adamw_step(b.flat_p.data_ptr() + o, b.flat_g.data_ptr() + o, b.m.data_ptr() + o, b.v.data_ptr() + o,
                                          end - start, b.stats.data_ptr(), float(group["lr"]), float(beta1), float(beta2),
                                          float(group["eps"]), float(group["weight_decay"]), max(step - first, 1), max_norm, inv_scale,
                                          _stream()), "osb_adamw_step")

    HYPER_ROWS = 32

    def _ensure_hyper(self, device) -> None:
        if self.hyper_dev is None:
            self.hyper_host = torch.zeros(self.HYPER_ROWS, 4, dtype=torch.float32)
            if torch.cuda.is_available():
                self.hyper_host = self.hyper_host.pin_memory()
            self.hyper_dev = torch.zeros(self.HYPER_ROWS, 4, dtype=torch.float32, device=device)

    def _hyper_row(self, gi: int, ci: int) -> int:
        """Row of the device hyper-parameter table for cohort `ci` of group `gi` (stable for the life of a captured graph:
        rows are only ever appended)."""
        rows = self.__dict__.setdefault("_hyper_row_ids", {})
        key = (gi, ci)
        if key not in rows:
            if len(rows) >= self.HYPER_ROWS:
                raise RuntimeError("FlatAdamW: too many (group, cohort) pairs for the device hyper-parameter table")
            rows[key] = len(rows)
        return rows[key]

    def stage_hyper(self) -> None:
        """Host side of one graph-mode step: advance the step counters, stage [lr, 1-b1^t, sqrt(1-b2^t)] of every cohort
        for the captured osb_adamw_step_dev launches (async copy on the current stream) and invalidate packed weights."""
        for gi, group in enumerate(self.param_groups):
            b = self._buckets.get(gi)
            if b is None:
                continue
            self._ensure_hyper(b.flat_p.device)
            step = self._steps.get(gi, 0) + 1
            self._steps[gi] = step
            beta1, beta2 = group["betas"]
            for ci, (_, _, first) in enumerate(b.cohorts):
                r = self._hyper_row(gi, ci)
                t = max(step - first, 1)
                self.hyper_host[r, 0] = float(group["lr"])
                self.hyper_host[r, 1] = 1.0 - beta1 ** t
                self.hyper_host[r, 2] = (1.0 - beta2 ** t) ** 0.5
            for p in b.params:
                p._osb_epoch = getattr(p, "_osb_epoch", 0) + 1
        if self.hyper_dev is not None:
            self.hyper_dev.copy_(self.hyper_host, non_blocking=True)

    def _kernel_step_dev(self, b: _Bucket, group, gi: int, max_norm: float, inv_scale: float) -> None:
        lib = _lib.load()
        self._ensure_hyper(b.flat_p.device)
        beta1, beta2 = group["betas"]
        for ci, (start, end, _) in enumerate(b.cohorts):
            o = 4 * start
            _lib.check(lib.osb_adamw_step_dev(b.flat_p.data_ptr() + o, b.flat_g.data_ptr() + o, b.m.data_ptr() + o, b.v.data_ptr() + o,
                                              end - start, b.stats.data_ptr(), self.hyper_dev[self._hyper_row(gi, ci)].data_ptr(),
                                              float(beta1), float(beta2), float(group["eps"]), float(group["weight_decay"]), max_norm,
                                              inv_scale, _stream()), "osb_adamw_step_dev")

    # -- checkpointing: torch.optim.AdamW layout ------------------------------------------------------
    def state_dict(self):
        """Same layout as torch.optim.AdamW.state_dict(): per-parameter `step`, `exp_avg`, `exp_avg_sq` (copies of the bucket
        views) for every parameter that has been updated, plus the param groups — what Lightning stores under
        `optimizer_states` (reference checkpoints; base_lightning_module.py relies on torch.optim.AdamW round-tripping)."""
        self.state.clear()
        for gi, group in enumerate(self.param_groups):
            b = self._buckets.get(gi)
            gstep = self._steps.get(gi, 0)
            for p in group["params"]:
                if b is not None and id(p) in b.ids:
                    first = b.cohorts[b.cohort_of[id(p)]][2]
                    self.state[p] = {"step": torch.tensor(float(max(gstep - first, 0))), "exp_avg": b.view_of(b.m, p).clone(),
                                     "exp_avg_sq": b.view_of(b.v, p).clone()}
                elif id(p) in self._pending_state:
                    m, v, st = self._pending_state[id(p)]
                    self.state[p] = {"step": torch.tensor(float(st)), "exp_avg": m.clone(), "exp_avg_sq": v.clone()}
        try:
            sd = super().state_dict()
        finally:
            self.state.clear()
        sd["osb_group_steps"] = {int(k): int(v) for k, v in self._steps.items()}
        return sd

    def load_state_dict(self, state_dict):
        """Accepts FlatAdamW.state_dict() and torch.optim.AdamW.state_dict() alike.  The moments are parked until the buckets
        are (re)built on the next step (bucket membership depends on which parameters receive gradients)."""
        sd = dict(state_dict)
        group_steps = sd.pop("osb_group_steps", None)
        super().load_state_dict(sd)
        self._pending_state = {}
        self._buckets = {}
        self._steps = {}
        for gi, group in enumerate(self.param_groups):
            best = 0
            for p in group["params"]:
                st = self.state.get(p)
                if not st or "exp_avg" not in st:
                    continue
                step = int(float(st["step"]))
                self._pending_state[id(p)] = (st["exp_avg"].detach().clone(), st["exp_avg_sq"].detach().clone(), step)
                best = max(best, step)
            if group_steps is not None and (gi in group_steps or str(gi) in group_steps):
                best = int(group_steps.get(gi, group_steps.get(str(gi))))
            if best:
                self._steps[gi] = best
        self.state.clear()

    def grad_norm(self) -> float:
        """Unscaled global gradient norm of the last step (reads a device scalar: call outside the hot loop)."""
        if self.last_stats is None:
            return float("nan")
        s = self.last_stats.tolist()
        return float("inf") if s[1] else (s[0] ** 0.5) / (self.loss_scale * self.world_size)
