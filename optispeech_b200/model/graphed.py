"""CUDA-graph replay of OptiSpeech.training_step.

The eager step is ~200 library launches plus ~350 small torch kernels for ~7 ms of device work: the host cannot issue
them as fast as a B200 retires them.  `GraphedTrainingStep` captures one whole step — forward, losses, backward, gradient
all-reduce, clip + AdamW — into a `torch.cuda.CUDAGraph` per (batch shapes, training phase) and replays it afterwards.

What makes the captured step reusable:
  * inputs live in static device buffers refreshed by async copies before each replay;
  * every per-step scalar is read from device memory: lr / bias corrections (`FlatAdamW.stage_hyper`,
    osb_adamw_step_dev) and the dropout seeds (host seed + the device step counter, include/osb200.h
    `dropout_seed_dev`); torch's own RNG ops advance their Philox offset per replay (torch.cuda.graph does that);
  * host bookkeeping the reference does per step (scheduler.step, batch counters — base_lightning_module.py:86-130)
    runs on the host around the replay, so `global_step`, the LR schedule and the pre-training gate behave as in eager;
  * packed fp16 weight copies of trainable parameters are rebuilt inside the graph (their epochs are bumped before the
    capture); packs of frozen parameters the graph reads are kept alive by the entry.

Gradient accumulation (`gradient_accumulate_batches` > 1) alternates two different steps and stays eager.

Graph mode needs a BOUNDED set of batch shapes: one graph (and its private memory pool) is kept per distinct
(shapes, phase) key, least-recently-used entries beyond `max_entries` are dropped, and every key costs `warmup` eager
steps plus a capture.  Pad or bucket variable-length batches (e.g. to multiples of 64 frames) before enabling it.
Steps that run eagerly while graphs exist (warm-up of a new key, the phase switch) are ordinary optimizer steps:
FlatAdamW takes its device-side hyper-parameter path only while a capture is in progress.
"""
from __future__ import annotations

from typing import Any, Dict

import numpy as np
import torch

from .. import _lib
from ..optim import FlatAdamW


class _Entry:
    def __init__(self):
        self.graph: torch.cuda.CUDAGraph = None
        self.static: Dict[str, torch.Tensor] = {}
        self.passthrough: Dict[str, Any] = {}
        self.keepalive = []
        self.launches = 0
        self.train_discriminator = False


def _as_tensor(v):
    return torch.from_numpy(v) if isinstance(v, np.ndarray) else v


class GraphedTrainingStep:
    def __init__(self, module, warmup: int = 3, max_entries: int = 8, step_fn=None, train_discriminator=None):
        """`step_fn(batch, batch_idx)` (default: the module's `_training_step_eager`) is what gets captured; a custom step must
        do its optimizer steps through the module's FlatAdamW optimizers.  `train_discriminator` pins the phase (and with it
        the optimizers whose hyper-parameters are staged per replay) instead of reading it from `global_step`."""
        self.module = module
        self.step_fn = step_fn
        self.force_phase = train_discriminator
        self.warmup = int(warmup)
        self.max_entries = int(max_entries)
        self._entries: Dict[tuple, _Entry] = {}   # insertion order = recency (re-inserted on every hit)
        self._warm: Dict[tuple, int] = {}
        self.replays = 0
        self.last_entry: _Entry = None
        self._capture_stream = None

    @staticmethod
    def _key(batch, train_discriminator: bool) -> tuple:
        sig = []
        for k in sorted(batch):
            v = batch[k]
            if isinstance(v, (torch.Tensor, np.ndarray)):
                sig.append((k, tuple(v.shape), str(v.dtype)))
            else:
                sig.append((k, repr(v)))
        return (bool(train_discriminator),) + tuple(sig)

    def __call__(self, batch, batch_idx):
        m = self.module
        acc = m.train_args.gradient_accumulate_batches
        step_fn = self.step_fn or m._training_step_eager
        if m.device.type != "cuda" or (acc is not None and acc != 1):
            return step_fn(batch, batch_idx)
        train_discriminator = (m.global_step >= m.train_args.pretraining_steps) if self.force_phase is None else bool(self.force_phase)
        key = self._key(batch, train_discriminator)
        entry = self._entries.get(key)
        if entry is None:
            seen = self._warm.get(key, 0)
            if seen < self.warmup:  # eager steps build the flat buckets, tables and kernel attributes the capture relies on
                self._warm[key] = seen + 1
                return step_fn(batch, batch_idx)
            entry = self._capture(batch, batch_idx, train_discriminator)
            while len(self._entries) >= self.max_entries:   # LRU: drop the stalest graph and its memory pool
                old = self._entries.pop(next(iter(self._entries)))
                old.graph, old.keepalive, old.static = None, [], {}
        else:
            self._entries.pop(key)
        self._entries[key] = entry
        self._replay(entry, batch)
        return None

    def release(self) -> None:
        """Drop every captured graph (and its private memory pool).  Call before tearing down a process group whose
        collectives were captured."""
        for e in self._entries.values():
            e.graph = None
            e.keepalive = []
            e.static = {}
        self._entries.clear()
        self.last_entry = None
        if self.module._optimizers is not None:
            for opt in self.module._optimizers:
                if isinstance(opt, FlatAdamW):
                    opt.graph_mode = False

    # ------------------------------------------------------------------------------------------------
    def _optimizers(self, train_discriminator: bool):
        opt_g, opt_d = self.module.optimizers()
        sched_g, sched_d = self.module.lr_schedulers()
        pairs = [(opt_g, sched_g)]
        if train_discriminator:
            pairs.append((opt_d, sched_d))
        return pairs

    def _capture(self, batch, batch_idx, train_discriminator: bool) -> _Entry:
        m = self.module
        dev = m.device
        e = _Entry()
        e.train_discriminator = train_discriminator
        for k, v in batch.items():
            if isinstance(v, (torch.Tensor, np.ndarray)):
                e.static[k] = _as_tensor(v).to(dev).clone()
            else:
                e.passthrough[k] = v
        for opt, _ in self._optimizers(train_discriminator):
            if not isinstance(opt, FlatAdamW):
                raise RuntimeError("cuda_graph=True needs the FlatAdamW optimizer (torch.optim.AdamW in the config maps to it)")
            opt.graph_mode = True
            opt._ensure_hyper(dev)
            for b in opt.buckets():  # force in-graph re-packing of every trainable weight
                for p in b.params:
                    p._osb_epoch = getattr(p, "_osb_epoch", 0) + 1
        lib = _lib.load()
        static_batch = dict(e.passthrough)
        static_batch.update(e.static)
        world = torch.distributed.get_world_size() if torch.distributed.is_available() and torch.distributed.is_initialized() else 1
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = lib.osb_launch_count()
        m._capturing = True
        try:
            # the NCCL watchdog thread polls events while we capture: only this thread's calls may invalidate the capture
            # captured on a high-priority stream: the step's main chain outranks the low-priority decoder / vocoder and
            # weight-gradient branches (ops.side_stream) when CTAs compete for SMs
            if self._capture_stream is None:
                self._capture_stream = torch.cuda.Stream(device=dev, priority=-1)
            with torch.cuda.graph(graph, stream=self._capture_stream, capture_error_mode="thread_local" if world > 1 else "global"):
                (self.step_fn or m._training_step_eager)(static_batch, batch_idx)
        finally:
            m._capturing = False
        e.launches = int(lib.osb_launch_count() - n0)
        e.graph = graph
        from .packing import live_packs
        e.keepalive = live_packs()
        packer = getattr(m.generator, "_step_packer", None)
        if packer is not None and packer.plan is not None:
            e.keepalive.append(packer.plan)   # the graph replays osb_pack_multi on this plan's job table and buffers
        return e

    def _replay(self, e: _Entry, batch) -> None:
        m = self.module
        for k, buf in e.static.items():
            buf.copy_(_as_tensor(batch[k]), non_blocking=True)
        pairs = self._optimizers(e.train_discriminator)
        for opt, _ in pairs:
            opt.stage_hyper()
        e.graph.replay()
        for _, sched in pairs:
            sched.step()
        m._fit.total_batch_idx += 1
        self.replays += 1
        self.last_entry = e
