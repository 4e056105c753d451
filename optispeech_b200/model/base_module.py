"""Training-step orchestration without Lightning.

Restates optispeech/model/base_lightning_module.py:24-186,294-303 (reference @ 3bdde20) in plain PyTorch:
manual optimisation of generator and discriminator, gradient-accumulation scaling, clip-by-norm + optimizer +
scheduler step, generator pre-training gate.  Lightning's services are replaced by small equivalents with the
same names (`optimizers`, `lr_schedulers`, `toggle_optimizer`, `manual_backward`, `clip_gradients`, `log_dict`),
so a `lightning.Trainer` could still drive the module (it only needs `training_step` / `configure_optimizers`).

Departures, all to keep the hot loop free of host synchronisation:
  * logged values stay on the device (`self.logged`), nothing calls .item() per step;
  * the ground-truth waveform crop is a device gather (the reference slices a numpy array per sample, :38-43);
  * `cache_generator_outputs=True` is honoured with `wav_hat.detach()` for the discriminator turn: the reference
    hands over the attached tensor, which would back-propagate through an already-freed graph (:112-119,161).
"""
from __future__ import annotations

import contextlib
import math
from types import SimpleNamespace
from typing import Any, Dict, Optional

import numpy as np
import torch
from torch import nn

from ..optim import FlatAdamW


OVERLAP_TURNS = True    # GAN phase: the discriminator turn is queued next to the generator's backward pass (see _training_step_eager)
DISC_TURN_SLOT = 19     # ops.side_stream slot of the discriminator turn
PREFETCH_REAL = True    # GAN phase: the discriminators' pass over the real signals starts before the generator forward (_process_batch)


class _FitLoopState:
    def __init__(self):
        self.total_batch_idx = 0


class BaseModule(nn.Module):
    """The subset of LightningModule behaviour the reference relies on."""

    def __init__(self):
        super().__init__()
        self.hparams = SimpleNamespace()
        self.automatic_optimization = False
        self.trainer: Any = None
        self.logged: Dict[str, torch.Tensor] = {}
        self._optimizers = None
        self._schedulers = None
        self._fit = _FitLoopState()
        self.loss_scale = 1024.0  # static loss scale for the fp16 tensor-core operands of the backward pass
        self.cuda_graph = False   # True: training_step replays a captured CUDA graph of itself (model/graphed.py)
        self._graphed = None
        self._capturing = False
        self.ckpt_loaded_epoch = -1

    # -- Lightning-compatible helpers ------------------------------------------------------------
    def save_hyperparameters(self, hparams: Dict[str, Any]):
        self.hparams = SimpleNamespace(**hparams)

    @property
    def device(self) -> torch.device:
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def optimizers(self):
        if self._optimizers is None:
            opts, scheds = self.configure_optimizers()
            self._optimizers = opts
            self._schedulers = [s["scheduler"] if isinstance(s, dict) else s for s in scheds]
        return self._optimizers

    def lr_schedulers(self):
        self.optimizers()
        return self._schedulers

    def toggle_optimizer(self, opt):
        """Freeze every parameter the optimizer does not own (Lightning semantics)."""
        owned = {id(p) for g in opt.param_groups for p in g["params"]}
        self._toggled = []
        for p in self.parameters():
            if id(p) not in owned and p.requires_grad:
                p.requires_grad_(False)
                self._toggled.append(p)

    def untoggle_optimizer(self, opt):
        for p in getattr(self, "_toggled", []):
            p.requires_grad_(True)
        self._toggled = []

    def manual_backward(self, loss, **kwargs):
        (loss * self.loss_scale).backward(**kwargs)

    def clip_gradients(self, opt, gradient_clip_val=None, gradient_clip_algorithm="norm"):
        """FlatAdamW: recorded and applied inside the fused optimizer step (osb_adamw_step unscales and clips by the global
        norm).  Any other optimizer (configure_optimizers then runs with loss_scale = 1, so there is nothing to unscale):
        data-parallel mean of the gradients, then torch's clip_grad_norm_ — what Lightning's DDP strategy + clip_gradients do
        (reference base_lightning_module.py:100-105)."""
        assert gradient_clip_algorithm == "norm"
        if isinstance(opt, FlatAdamW):
            opt.max_grad_norm = float(gradient_clip_val or 0.0)
            return
        params = [p for g in opt.param_groups for p in g["params"] if p.grad is not None]
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            ws = torch.distributed.get_world_size()
            for p in params:
                torch.distributed.all_reduce(p.grad)
                p.grad.div_(ws)
        if gradient_clip_val:
            torch.nn.utils.clip_grad_norm_(params, gradient_clip_val)

    def log_dict(self, values: Dict[str, Any], **kwargs):
        for k, v in values.items():
            self.logged[k] = v.detach() if isinstance(v, torch.Tensor) else v

    @property
    def global_step(self) -> int:
        """Batches processed, divided by the accumulation factor (reference :294-303)."""
        total = self.trainer.fit_loop.total_batch_idx if self.trainer is not None else self._fit.total_batch_idx
        acc = self.train_args.gradient_accumulate_batches
        return int(total // acc) if acc is not None else int(total)

    def on_load_checkpoint(self, checkpoint: Dict[str, Any]) -> None:
        self.ckpt_loaded_epoch = checkpoint.get("epoch", -1)

    # -- reference training logic -------------------------------------------------------------------
    def _upload_early(self, batch):
        """Enqueue the host -> device copies of the batch's (pinned) host tensors NOW, on a copy stream, into a ring of
        persistent device buffers, and return the batch with those entries replaced.  The DMA transfers run while the host
        cuts the waveform crop — and while the PREVIOUS step still computes: the compute stream only waits for the copy
        event, and a buffer set is overwritten only after the step that read it has passed its release event."""
        dev = self.device
        if dev.type != "cuda":
            return batch
        st = self.__dict__.setdefault("_upload_state", {"stream": torch.cuda.Stream(device=dev), "sets": {}, "turn": 0})
        copy_stream = st["stream"]
        turn = st["turn"] % 3
        st["turn"] += 1
        out = dict(batch)
        main = torch.cuda.current_stream(dev)
        rel = st.get(("released", turn))
        if rel is not None:
            copy_stream.wait_event(rel)           # the step that last read this buffer set has been enqueued and passed
        with torch.cuda.stream(copy_stream):
            for k, v in batch.items():
                if k == "wav" or not isinstance(v, torch.Tensor) or v.is_cuda:
                    continue
                key = (turn, k, tuple(v.shape), v.dtype)
                buf = st["sets"].get(key)
                if buf is None:
                    buf = torch.empty(v.shape, dtype=v.dtype, device=dev)
                    st["sets"][key] = buf
                buf.copy_(v, non_blocking=True)
                out[k] = buf
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        main.wait_event(ev)
        self._upload_turn = turn
        return out

    def stage_batch(self, batch, upload: bool = False):
        """Host half of the reference's `_process_batch` (base_lightning_module.py:38-43): when the collate function hands the
        waveform and the lengths over in HOST memory — as the reference's does (`wav` is a numpy array,
        text_wav_datamodule.py:253-266) — the segment start indices are drawn from the CPU generator exactly as
        `get_random_segments` does (utils/segments.py:29-35: they depend on `mel_lengths` only) and the ground-truth crop is
        cut on the host (`get_segments_numpy`, utils/segments.py:63-72).  Only the (B, segment*hop) crop and the (B,) draw
        travel to the device instead of the whole (B, Tw) waveform (2 MB instead of 28 MB at B=32 x 864 frames); the device
        recomputes the same start indices from the same fp32 draw.  Batches already on the device are returned unchanged."""
        wav, ml = batch.get("wav"), batch.get("mel_lengths")
        if wav is None or "wav_segment" in batch:
            return batch
        wav_on_host = isinstance(wav, np.ndarray) or (isinstance(wav, torch.Tensor) and not wav.is_cuda)
        if not (wav_on_host and isinstance(ml, torch.Tensor) and not ml.is_cuda):
            return batch
        B = int(ml.shape[0])
        seg = min(int(self.generator.segment_size), int(batch["mel"].shape[-1]))
        if upload:   # (training_step) start the transfers of the other tensors before the host-side work below
            batch = self._upload_early(batch)
        rand = torch.rand(B)
        start = (rand * (ml.to(torch.float32) - 4 - seg).clamp(min=0)).to(torch.long)
        hop = int(self.hop_length)
        n = seg * hop
        slot = self._stage_slot(B, n)
        crop, rand_buf = slot["crop"], slot["rand"]
        rand_buf.copy_(rand)
        w = wav if isinstance(wav, np.ndarray) else wav.numpy()
        if w.ndim == 3:
            w = w[:, 0]
        dst = crop.numpy()
        for b, s in enumerate((start * hop).tolist()):
            chunk = w[b, s: s + n]
            dst[b, : chunk.shape[0]] = chunk
            if chunk.shape[0] < n:
                dst[b, chunk.shape[0]:] = 0.0
        staged = {k: v for k, v in batch.items() if k != "wav"}
        staged["wav_segment"] = crop
        staged["seg_rand"] = rand_buf
        self._stage_live = slot
        return staged

    def _stage_slot(self, B: int, n: int):
        """Pinned staging buffers, a ring of four per (B, n): a slot is reused only after the step that copied from it has
        passed the event recorded behind its copies (allocating pinned memory per step would cost more than the copy)."""
        ring = self.__dict__.setdefault("_stage_rings", {}).setdefault((B, n), {"slots": [], "next": 0})
        pin = torch.cuda.is_available()
        if len(ring["slots"]) < 4:
            slot = {"crop": torch.zeros((B, n), dtype=torch.float32, pin_memory=pin), "rand": torch.zeros(B, pin_memory=pin), "event": None}
            ring["slots"].append(slot)
            return slot
        slot = ring["slots"][ring["next"] % 4]
        ring["next"] += 1
        if slot["event"] is not None:
            slot["event"].synchronize()
        return slot

    def _stage_release(self):
        slot = self.__dict__.pop("_stage_live", None)
        if slot is not None and self.device.type == "cuda":
            ev = slot["event"] or torch.cuda.Event()
            ev.record()
            slot["event"] = ev
        turn = self.__dict__.pop("_upload_turn", None)
        if turn is not None and self.device.type == "cuda":   # the uploaded buffer set may be overwritten once this point has passed
            st = self._upload_state
            ev = st.get(("released", turn)) or torch.cuda.Event()
            ev.record()
            st[("released", turn)] = ev

    @staticmethod
    def batch_h2d_bytes(batch) -> int:
        """Bytes a (staged) batch moves host -> device per step."""
        return int(sum(v.numel() * v.element_size() if isinstance(v, torch.Tensor) else v.nbytes
                       for v in batch.values() if isinstance(v, (torch.Tensor, np.ndarray)) and not (isinstance(v, torch.Tensor) and v.is_cuda)))

    def _process_batch(self, batch, vocoder_grad: bool = True, prefetch_real: bool = False):
        """Generator forward + ground-truth crop (reference base_lightning_module.py:24-45).  `prefetch_real` (GAN phase): the
        ground-truth crop depends on the batch alone, so it is cut BEFORE the generator runs and the discriminators' forward
        pass over the real signals is queued on their streams right away (VocosDiscriminator.prefetch_real) — it then runs next
        to the generator's forward pass, whose kernels leave most of the machine idle, instead of after it."""
        dev = self.device
        sids, lids = batch.get("sids"), batch.get("lids")
        seg_rand = batch.get("seg_rand")
        mel_lengths = batch["mel_lengths"].to(dev, non_blocking=True)
        wav_real = None
        prefetch_real = prefetch_real and dev.type == "cuda" and PREFETCH_REAL
        if "wav_segment" in batch:  # host-staged crop (stage_batch)
            wav_real = batch["wav_segment"].to(dev, non_blocking=True).to(torch.float32)
        elif prefetch_real:
            from .. import ops
            seg_frames = min(int(self.generator.segment_size), int(batch["mel"].shape[-1]))
            if seg_rand is None:   # the draw the generator would make (generator/training.py), made here and handed over
                seg_rand = torch.rand([mel_lengths.shape[0]], device=dev) if torch.cuda.is_current_stream_capturing() else torch.rand([mel_lengths.shape[0]])
            seg_rand = seg_rand.to(dev, non_blocking=True).float()
            start_idx = ops.segment_starts(seg_rand, mel_lengths, seg_frames, margin=4)
            wav_real = ops.crop_segments(self._device_wav(batch["wav"]), start_idx, seg_frames * self.hop_length, self.hop_length)
        if prefetch_real and wav_real is not None:
            self.discriminator.prefetch_real(wav_real)
        gen_outputs = self.generator(
            x=batch["x"].to(dev, non_blocking=True), x_lengths=batch["x_lengths"].to(dev, non_blocking=True),
            mel=batch["mel"].to(dev, non_blocking=True), mel_lengths=mel_lengths,
            pitches=batch["pitches"].to(dev, non_blocking=True), energies=batch["energies"].to(dev, non_blocking=True),
            sids=sids.to(dev) if sids is not None else None, lids=lids.to(dev) if lids is not None else None,
            **({"seg_rand": seg_rand.to(dev, non_blocking=True)} if seg_rand is not None else {}),
        )
        if wav_real is not None:
            gen_outputs["wav"] = wav_real.type_as(gen_outputs["wav_hat"])
            return gen_outputs
        seg = gen_outputs["segment_size"] * self.hop_length
        wav = self._device_wav(batch["wav"])
        from .. import ops
        # ground-truth crop wav[b, start*hop : start*hop + seg], zero-filled past the end like get_segments_numpy's pre-zeroed
        # array (utils/segments.py:63-72) and the host path (stage_batch).  start_idx is produced by the decoder / vocoder branch:
        # when that branch has not been joined yet (pre-training, deferred join) the crop is queued on the branch's stream, behind
        # the kernel that writes start_idx — on the current stream it would read the previous step's indices.
        pending = gen_outputs.get("_pending_streams") or []
        if pending:
            pending[0].wait_stream(torch.cuda.current_stream())      # `wav` was made contiguous / fp32 on the current stream
            with torch.cuda.stream(pending[0]):
                gen_outputs["wav"] = ops.crop_segments(wav, gen_outputs["start_idx"], seg, self.hop_length).type_as(gen_outputs["wav_hat"])
            gen_outputs.setdefault("_pending_keepalive", []).append(wav)
        else:
            gen_outputs["wav"] = ops.crop_segments(wav, gen_outputs["start_idx"], seg, self.hop_length).type_as(gen_outputs["wav_hat"])
        return gen_outputs

    def _device_wav(self, wav):
        wav = torch.from_numpy(wav) if isinstance(wav, np.ndarray) else wav
        if wav.dim() == 3:      # (B, 1, Tw) as the reference's collate function hands it over (text_wav_datamodule.py:253-266)
            wav = wav[:, 0]
        return wav.to(self.device, non_blocking=True).to(torch.float32).contiguous()

    def configure_optimizers(self):
        gen_params = [{"params": list(self.generator.parameters())}]
        disc_params = [{"params": list(self.discriminator.parameters())}]
        make = self.hparams.optimizer
        target = getattr(make, "func", make)
        if target is torch.optim.AdamW:  # the fused flat-bucket implementation, same hyper-parameters
            kw = dict(getattr(make, "keywords", {}))
            kw["betas"] = tuple(kw.get("betas", (0.9, 0.999)))
            ws = torch.distributed.get_world_size() if torch.distributed.is_available() and torch.distributed.is_initialized() else 1
            opt_gen = FlatAdamW(gen_params, loss_scale=self.loss_scale, world_size=ws, **kw)
            opt_disc = FlatAdamW(disc_params, loss_scale=self.loss_scale, world_size=ws, **kw)
        else:
            # stock optimizers see plain gradients: the static loss scale is a FlatAdamW service (it unscales in its fused step)
            self.loss_scale = 1.0
            opt_gen, opt_disc = make(gen_params), make(disc_params)
        max_steps = (self.trainer.max_steps if self.trainer is not None else getattr(self, "max_steps", 2_000_000)) // 2
        acc = self.train_args.gradient_accumulate_batches
        if acc is not None:
            max_epochs = getattr(self.trainer, "max_epochs", None) if self.trainer is not None else None
            max_steps = math.ceil(max_steps / acc) * max(max_epochs if max_epochs is not None else -1, 1)
        sched = self.hparams.scheduler
        if hasattr(sched, "keywords") and "num_training_steps" in sched.keywords:
            sched.keywords["num_training_steps"] = max_steps
        # reference :66-67 passes getattr("self", ...) (a string), i.e. last_epoch is always -1
        scheduler_gen = sched(opt_gen, last_epoch=-1)
        scheduler_disc = sched(opt_disc, last_epoch=-1)
        return ([opt_gen, opt_disc],
                [{"scheduler": scheduler_gen, "interval": "step"}, {"scheduler": scheduler_disc, "interval": "step"}])

    def training_step(self, batch, batch_idx, **kwargs):
        """Reference base_lightning_module.py:86-130.  With `self.cuda_graph = True` the same step is captured once per
        batch shape and training phase and replayed afterwards (one graph launch instead of ~550 kernel launches)."""
        if self._capturing:
            return self._training_step_eager(batch, batch_idx, **kwargs)
        batch = self.stage_batch(batch, upload=True)
        try:
            if self.cuda_graph:
                if self._graphed is None:
                    from .graphed import GraphedTrainingStep
                    self._graphed = GraphedTrainingStep(self)
                return self._graphed(batch, batch_idx)
            return self._training_step_eager(batch, batch_idx, **kwargs)
        finally:
            self._stage_release()

    def _training_step_eager(self, batch, batch_idx, **kwargs):
        acc = self.train_args.gradient_accumulate_batches
        if acc is not None:
            loss_scaling_factor = float(acc)
            should_apply = (batch_idx + 1) % acc == 0
        else:
            loss_scaling_factor, should_apply = 1.0, True
        train_discriminator = self.global_step >= self.train_args.pretraining_steps
        opt_g, opt_d = self.optimizers()
        sched_g, sched_d = self.lr_schedulers()

        if self.device.type == "cuda":
            from .. import ops
            ops.step_counter(self.device).add_(1)  # device-side ingredient of the dropout seeds (graph-replay safe)
        self.toggle_optimizer(opt_g)
        loss_g, wav_outputs = self.training_step_g(batch, train_discriminator=train_discriminator)
        loss_g = loss_g / loss_scaling_factor
        # The discriminator turn reads the two waveforms, the discriminator weights and what the discriminators' forward
        # left behind (layer outputs, weight packs) — all complete here — and nothing the generator's backward pass or
        # optimizer writes.  With cached generator outputs (the reference default) it is therefore queued on its own stream
        # forked HERE, so its kernels run next to the generator's backward pass instead of behind the generator optimizer.
        overlap = (train_discriminator and OVERLAP_TURNS and self.device.type == "cuda" and self.train_args.cache_generator_outputs)
        if overlap:
            fork_event = torch.cuda.Event()
            fork_event.record()
        if should_apply:
            opt_g.zero_grad()
        self.manual_backward(loss_g)
        self.clip_gradients(opt_g, gradient_clip_val=self.train_args.gradient_clip_val, gradient_clip_algorithm="norm")
        if should_apply:
            opt_g.step()
            if not self._capturing:  # host bookkeeping of a captured step is done per replay (model/graphed.py)
                sched_g.step()
        self.untoggle_optimizer(opt_g)
        for s in self.__dict__.pop("_pending_streams", []):   # branches the generator turn left running (see training_step_g)
            torch.cuda.current_stream().wait_stream(s)
        self.__dict__.pop("_pending_keepalive", None)
        if not self._capturing:
            self._fit.total_batch_idx += 1
        if not train_discriminator:
            return

        self.toggle_optimizer(opt_d)
        if not self.train_args.cache_generator_outputs:
            wav_outputs = None
        turn_ctx = contextlib.nullcontext()
        if overlap:
            from .. import ops
            from .vocoder.wavenext.disc import native as disc_native
            main_stream = torch.cuda.current_stream()
            turn_stream = ops.side_stream(self.device, DISC_TURN_SLOT)
            if turn_stream != main_stream:
                turn_stream.wait_event(fork_event)
                turn_ctx = torch.cuda.stream(turn_stream)
                disc_native.STREAM_SET = 1
        try:
            with turn_ctx:
                loss_d = self.training_step_d(batch, wav_outputs=wav_outputs) / loss_scaling_factor
                if should_apply:
                    opt_d.zero_grad()
                self.manual_backward(loss_d)
                self.clip_gradients(opt_d, gradient_clip_val=self.train_args.gradient_clip_val, gradient_clip_algorithm="norm")
                if should_apply:
                    opt_d.step()
                    if not self._capturing:
                        sched_d.step()
        finally:
            if overlap:
                disc_native.STREAM_SET = 0
                if turn_stream != main_stream:
                    main_stream.wait_stream(turn_stream)
        self.untoggle_optimizer(opt_d)

    def training_step_g(self, batch, train_discriminator):
        log_outputs = {}
        if train_discriminator:
            gen_outputs = self._process_batch(batch, prefetch_real=True)
        else:
            # pre-training: the vocoder output feeds no loss, so its autograd graph is not built and its stream is only joined
            # at the end of the step (the decoder / vocoder forward overlaps the backward pass)
            self.generator.vocoder_needs_grad = False
            self.generator.defer_vocoder_join = True
            try:
                gen_outputs = self._process_batch(batch)
            finally:
                self.generator.vocoder_needs_grad = True
                self.generator.defer_vocoder_join = False
        self._pending_streams = list(gen_outputs.pop("_pending_streams", []) or [])
        self._pending_keepalive = gen_outputs.pop("_pending_keepalive", [])
        gen_am_loss = gen_outputs["loss"]
        log_outputs.update({
            "total_loss/train_am_loss": gen_am_loss,
            "gen_subloss/train_alighn_loss": gen_outputs["align_loss"],
            "gen_subloss/train_duration_loss": gen_outputs["duration_loss"],
            "gen_subloss/train_pitch_loss": gen_outputs["pitch_loss"],
            "gen_subloss/train_energy_loss": gen_outputs["energy_loss"],
        })
        wav, wav_hat = gen_outputs["wav"], gen_outputs["wav_hat"]
        if train_discriminator:
            gen_adv_loss, log_dict = self.discriminator.forward_gen(wav, wav_hat)
            log_outputs["total_loss/train_gen_adv_loss"] = gen_adv_loss
            log_outputs.update({f"gen_adv_loss/train_{k}": v for k, v in log_dict.items()})
        else:
            gen_adv_loss = 0.0
        loss = gen_am_loss + gen_adv_loss
        log_outputs["total_loss/generator"] = loss
        self.log_dict(log_outputs)
        return loss, (wav.detach(), wav_hat.detach())

    def training_step_d(self, batch, wav_outputs=None):
        if wav_outputs is None:
            with torch.no_grad():
                gen_outputs = self._process_batch(batch)
            wav, wav_hat = gen_outputs["wav"], gen_outputs["wav_hat"]
        else:
            wav, wav_hat = wav_outputs
        loss, log_dict = self.discriminator.forward_disc(wav, wav_hat)
        log_outputs = {"total_loss/discriminator": loss}
        log_outputs.update({f"discriminator/{k}": v for k, v in log_dict.items()})
        self.log_dict(log_outputs)
        return loss
