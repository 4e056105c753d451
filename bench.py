#!/usr/bin/env python
"""Benchmark of the OptiSpeech training hot path on B200 (contract: see the build prompt / DESIGN.md §Measurement).

    python bench.py --gpus N --steps K --warmup W              # this implementation (libosb200 kernels)
    python bench.py --impl reference --gpus N --steps K ...    # the UNMODIFIED reference modules on the host CPU (rank 0 only)
    python bench.py --impl torch-gpu --steps K ...             # the same reference modules .cuda(): the stock-PyTorch bar

Workload (BASELINE.json configs[1], SURVEY §8d "S-train"): ConvNeXt configuration, per-GPU batch of 32 synthetic
utterances, 192 phoneme positions, 864 mel frames (22.05 kHz, hop 256), random-initialised weights, train mode.  One step
is `training_step` in the generator pre-training phase (reference base_lightning_module.py:78-110: generator forward,
acoustic-model loss, backward, clip-by-norm 10, AdamW, cosine schedule).  The GAN-phase step and the synthesis path are
reported under "variants" / "synthesis" with their own rooflines.
Metric: mel frames processed per second (sum of mel_lengths over all ranks / step time).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_PER_GPU, TX, TM = 32, 192, 864
SEED = 1234
METRIC, UNIT = "mel_frames_per_sec_train_step", "mel-frames/s"

# SURVEY §8(d) algorithmic work
FLOP_PER_FRAME_SYNTH = 21_096_960          # decoder + vocoder embed + 8 vocoder blocks + head, per mel frame
FLOP_PER_PHONEME_SYNTH = 10_445_824        # encoder + duration / pitch / energy predictors, per phoneme
T1_FLOP_PER_SAMPLE = 13.33e9               # fwd 7.494 + bwd 5.837 GFLOP per utterance at Tx=192, Tm=864


def workload_config(world: int) -> dict:
    """The `config` object of the JSON line: identical for every arm (ours, reference, torch-gpu)."""
    return {"workload": f"ConvNeXt OptiSpeech training_step, generator pre-training phase (fwd + bwd + clip 10 + AdamW + cosine LR), "
                        f"B={B_PER_GPU}/GPU, Tx={TX}, Tm={TM}, 22.05 kHz hop 256, random-init weights, train mode (dropout + DropPath on)",
            "global_batch": B_PER_GPU * world, "parallelism": f"dp{world}"}


# --------------------------------------------------------------------------------------------------
# synthetic workload
# --------------------------------------------------------------------------------------------------
def make_batch(B: int, seed: int, n_feats: int = 100, hop: int = 256):
    g = torch.Generator().manual_seed(seed)
    x_lengths = torch.randint(TX // 2, TX + 1, (B,), generator=g)
    x_lengths[0] = TX
    x = torch.randint(1, 159, (B, TX), generator=g) * (torch.arange(TX)[None] < x_lengths[:, None])
    mel_lengths = torch.clamp((4.5 * x_lengths.float()).round().long(), max=TM)
    mel_lengths[0] = TM
    mmask = torch.arange(TM)[None] < mel_lengths[:, None]
    mel = torch.randn(B, n_feats, TM, generator=g) * mmask[:, None, :]
    pitches = torch.randn(B, TM, generator=g) * mmask
    energies = torch.randn(B, TM, generator=g) * mmask
    wav = torch.rand(B, TM * hop, generator=g) * 2 - 1
    return dict(x=x, x_lengths=x_lengths, mel=mel, mel_lengths=mel_lengths, pitches=pitches, energies=energies, wav=wav,
                sids=None, lids=None)


def synth_inputs(name: str):
    """SURVEY §8d S-synth-long (B=8 x 512 phonemes, ~10 s each) and S-synth-1 (B=1 x 120): ids, lengths, fixed integer durations
    (~2 frames / phoneme; random-init duration heads give ~1.6) so that every arm synthesises the same number of frames."""
    B, Tx = {"long_B8_Tx512": (8, 512), "single_B1_Tx120": (1, 120)}[name]
    g = torch.Generator().manual_seed(SEED + B)
    ids = torch.randint(1, 159, (B, Tx), generator=g)
    lens = torch.full((B,), Tx, dtype=torch.int64)
    durs = torch.randint(1, 4, (B, Tx), generator=g)
    return ids, lens, durs


# --------------------------------------------------------------------------------------------------
# clocks: NVML polled every 10 ms from a thread, from before the warm-up to the end of the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int, period_s: float = 0.004):
        self.gpu, self.period = gpu_index, period_s
        self.samples = []          # (t, sm_mhz, reasons bitmask, power_w)
        self.sm_max = None
        self._stop = threading.Event()
        self._thread = None
        self.window = None         # (t0, t1) of the timed region
        self.load_start = None     # start of the warm-up (everything before it is idle set-up)

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            # NVML enumerates physical GPUs: map through CUDA_VISIBLE_DEVICES when it is a plain index list
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except (ValueError, IndexError):
                    idx = self.gpu
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            return

        def loop():
            while not self._stop.is_set():
                try:
                    sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                    try:
                        rs = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                    except Exception:
                        rs = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                    self.samples.append((time.perf_counter(), sm, rs, pw))
                except Exception:
                    pass
                self._stop.wait(self.period)

        self._thread = threading.Thread(target=loop, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        out = {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0, "samples_in_timed_region": 0, "power_w_max": None,
               "how": "NVML polled every 4 ms from before the warm-up to the end of the timed region; sm_mhz = median over the samples "
                      "taken under load (timed region, else warm-up + timed region)"}
        if not self.samples:
            return out
        inside = [s for s in self.samples if self.window and self.window[0] <= s[0] <= self.window[1]]
        loaded = [s for s in self.samples if self.load_start is None or s[0] >= self.load_start]
        use = inside if len(inside) >= 5 else (loaded or self.samples)
        mask = 0
        for s in use:
            mask |= s[2]
        out.update(sm_mhz=float(np.median([s[1] for s in use])), reasons=sorted(n for b, n in self.REASONS.items() if mask & b),
                   samples=len(use), samples_in_timed_region=len(inside), power_w_max=float(max(s[3] for s in use)))
        return out


# --------------------------------------------------------------------------------------------------
# reference arms: the unmodified reference modules (oracle/_ref or /root/reference) on the CPU / on the same GPU
# --------------------------------------------------------------------------------------------------
class ReferenceStep:
    """`training_step` of the reference in the pre-training phase, restated around its own modules
    (base_lightning_module.py:24-45,78-110): `_process_batch` = generator(**batch.to(device)) + numpy crop
    of the ground-truth waveform at start_idx*hop + H2D; manual_backward; clip_gradients(norm, 10); AdamW; cosine schedule."""

    def __init__(self, device: str, amp: bool = False):
        from oracle import ref_harness as RH
        from oracle.spec import ModelSpec, deterministic_state_dict, generator_shapes

        self.RH = RH
        RH.import_reference()
        self.spec = ModelSpec()
        self.device = torch.device(device)
        gen, _ = RH.build_reference_generator(self.spec)
        gen.load_state_dict(deterministic_state_dict(generator_shapes(self.spec), seed=0, frames_per_token=4.5), strict=True)
        self.gen = gen.to(self.device).train()
        self.opt, self.sched = RH.reference_optimizer(self.gen)
        self.batch = make_batch(B_PER_GPU, SEED)
        self.wav_np = self.batch["wav"].numpy()[:, None, :]     # (B, 1, Tw) numpy, as TextWavBatchCollate hands it over
        if self.device.type == "cuda":
            self.batch = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in self.batch.items()}
        self.frames = int(self.batch["mel_lengths"].sum())
        self.amp = amp
        self.scaler = torch.amp.GradScaler("cuda", enabled=amp) if self.device.type == "cuda" else None

    def step(self) -> float:
        import optispeech.utils.segments as seg

        dev = self.device if self.device.type == "cuda" else None
        ctx = torch.autocast("cuda", dtype=torch.float16) if self.amp else torch.autocast("cpu", enabled=False)
        with ctx:
            out = self.RH.run_reference_forward(self.gen, self.batch, dev)
        hop = self.spec.hop_length
        crop = seg.get_segments_numpy(self.wav_np, (out["start_idx"] * hop).cpu().numpy(), out["segment_size"] * hop)
        wav = torch.from_numpy(crop).squeeze(1).to(self.device).type_as(out["wav_hat"])   # the (unused in this phase) GT crop
        del wav
        loss = out["loss"]
        self.opt.zero_grad()
        params = [p for g in self.opt.param_groups for p in g["params"]]
        if self.scaler is not None and self.amp:
            self.scaler.scale(loss).backward()
            self.scaler.unscale_(self.opt)
            torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 10.0)
            self.scaler.step(self.opt)
            self.scaler.update()
        else:
            loss.backward()
            torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 10.0)
            self.opt.step()
        self.sched.step()
        return float(loss.detach())     # the step's loss read-back (a host sync on the GPU)

    def synthesise(self, name: str):
        ids, lens, durs = synth_inputs(name)
        gen = self.gen
        was_training = gen.training
        gen.eval()
        # hold the integer durations fixed (same frames as every other arm): DurationPredictor.infer -> the given durations
        dp = gen.duration_predictor
        orig = dp.infer
        # (the predictor still runs: its compute is part of the synthesis work)
        dp.infer = lambda x, mask, factor=1.0: (orig(x, mask, factor=factor) * 0 + durs.to(x.device)).masked_fill(mask, 0)
        try:
            with torch.inference_mode():
                t0 = time.perf_counter()
                out = gen.synthesise(ids.to(self.device), lens.to(self.device))
                if self.device.type == "cuda":
                    torch.cuda.synchronize()
                dt = time.perf_counter() - t0
        finally:
            dp.infer = orig
            gen.train(was_training)
        return int(out["wav_lengths"].sum()), dt


def _port_step_factory(sample_b: int):
    """Fallback when the reference tree is not on the box: the oracle port (kind "port")."""
    from oracle import model as O
    from oracle.spec import ModelSpec, deterministic_state_dict, generator_shapes

    spec = ModelSpec()
    sd = deterministic_state_dict(generator_shapes(spec), seed=0, frames_per_token=4.5)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.AdamW(list(sd.values()), lr=2e-4, betas=(0.8, 0.99), weight_decay=1e-2)
    full = make_batch(B_PER_GPU, SEED)
    b = {k: (v[:sample_b] if isinstance(v, torch.Tensor) else v) for k, v in full.items()}
    frames = int(b["mel_lengths"].sum())

    def step():
        out = O.generator_forward(sd, spec, b["x"], b["x_lengths"], b["mel"], b["mel_lengths"], b["pitches"], b["energies"], torch.rand(sample_b))
        opt.zero_grad()
        out["loss"].backward()
        torch.nn.utils.clip_grad_norm_([p for p in sd.values() if p.grad is not None], 10.0)
        opt.step()
        return float(out["loss"].detach())

    return step, frames


def run_reference_arm(args, rank: int, device: str):
    """--impl reference (device cpu) / --impl torch-gpu (device cuda, fp32 and autocast fp16)."""
    if rank != 0:
        return
    from oracle import ref_harness as RH

    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    have_ref = RH.reference_root() is not None
    line = {"impl": "reference" if device == "cpu" else "torch-gpu", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic",
            "config": workload_config(max(args.gpus, 1))}
    if device == "cpu":
        if have_ref:
            ref = ReferenceStep("cpu")
            step, frames = ref.step, ref.frames
            kind = "reference"
            sample = (f"the full B={B_PER_GPU} batch per step (Tx={TX}, Tm={TM}), UNMODIFIED reference modules (oracle/_ref, mush42/optispeech "
                      f"@ 3bdde20) in train mode, fp32, {cores} host threads; step = _process_batch + backward + clip 10 + AdamW + cosine LR")
        else:
            step, frames = _port_step_factory(2)
            kind = "port"
            sample = f"B=2 of the B={B_PER_GPU} batch per step (oracle port; reference tree not on this box), fp32, {cores} threads"
        # warm-up: numba JIT + prior cache need one step; a CPU step takes seconds, so cap the rest when it is slow
        t0 = time.perf_counter()
        step()
        first = time.perf_counter() - t0
        extra = max(0, args.warmup - 1) if first < 8.0 else min(1, max(0, args.warmup - 1))
        for _ in range(extra):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
        value = frames * args.steps / dt
        line.update(value=value, ms_per_step=1e3 * dt / args.steps, dtype="f32", warmup_done=1 + extra,
                    cpu_baseline={"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                    e2e={"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        if have_ref and not args.no_variants:
            synth = {}
            for name in ("single_B1_Tx120", "long_B8_Tx512"):
                try:
                    ref.synthesise(name)
                    n, dt_s = ref.synthesise(name)
                    synth[name] = {"audio_samples_per_s": n / dt_s, "ms": 1e3 * dt_s, "frames": n // 256, "cores": cores, "kind": "reference"}
                except Exception as exc:
                    synth[name] = {"error": repr(exc)[:200]}
            line["synthesis"] = synth
        print(json.dumps(line), flush=True)
        return
    # ---- stock PyTorch on the same GPU: reference modules .cuda(), fp32 and autocast(fp16) (the library bar)
    variants = {}
    for name, amp in (("fp32", False), ("autocast_fp16", True)):
        try:
            torch.manual_seed(SEED)
            ref = ReferenceStep("cuda", amp=amp)
            for _ in range(max(args.warmup, 2)):
                ref.step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            loss = 0.0
            for _ in range(args.steps):
                loss = ref.step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            variants[name] = {"value": ref.frames / (ms * 1e-3), "ms_per_step": ms, "loss": loss, "finite": bool(np.isfinite(loss))}
            if name == "fp32" and not args.no_variants:
                synth = {}
                for sname in ("single_B1_Tx120", "long_B8_Tx512"):
                    for _ in range(2):
                        ref.synthesise(sname)
                    n, dt_s = ref.synthesise(sname)
                    synth[sname] = {"audio_samples_per_s": n / dt_s, "ms": 1e3 * dt_s, "frames": n // 256}
                variants[name]["synthesis"] = synth
            del ref
            torch.cuda.empty_cache()
        except Exception as exc:
            variants[name] = {"error": repr(exc)[:300]}
    ok = {k: v for k, v in variants.items() if "value" in v and v.get("finite", True)}
    best = max(ok.values(), key=lambda v: v["value"]) if ok else None
    line.update(value=best["value"] if best else None, ms_per_step=best["ms_per_step"] if best else None, dtype="f32 / f16 autocast",
                variants=variants,
                note="the UNMODIFIED reference modules moved to the GPU (stock cuDNN / cuBLASLt / ATen kernels; its per-sample numba MAS, "
                     "scipy prior and CTC loops stay on the host as the reference wrote them); pinned host batch copied per step, loss read per step")
    print(json.dumps(line), flush=True)


def _run_sub(argv, timeout_s: int):
    """Run another arm of this script in a fresh process (the reference owns the module name `optispeech` there)."""
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True, timeout=timeout_s, env=env)
        for ln in reversed(res.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": (res.stderr or res.stdout)[-300:]}
    except Exception as exc:
        return {"error": repr(exc)[:300]}


def ncu_traffic(kernel: str):
    """DRAM bytes (read + write) per launch of `kernel` from the committed `ncu --set full` capture (profiles/ncu_traffic.json,
    written by tools/summarize_ncu.py), or None."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return table.get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


# --------------------------------------------------------------------------------------------------
# main arm
# --------------------------------------------------------------------------------------------------
def max_over_ranks(ms: float, world: int, dev) -> float:
    if world <= 1:
        return ms
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def timed_steps(fn, steps: int, world: int, dev, sampler=None) -> float:
    """K steps bracketed by barrier + synchronize on both sides; device time from CUDA events, max over ranks."""
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    if sampler is not None:
        sampler.window = (t0, time.perf_counter())
    if world > 1:
        torch.distributed.barrier()
    return max_over_ranks(e0.elapsed_time(e1), world, dev) / steps


def _event_time(fn, n=10, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def spectral_kernel_bench(model, dev, peaks):
    """The fused STFT / mel loss kernels timed alone (forward + gradient through `forward_val`), at the training segment size and
    at long-form size.  Bound: every input sample is read once for the prediction and once for the target (8 B per sample
    position per resolution pass) — reported as achieved GB/s of that algorithmic traffic against the measured HBM
    peak; the fused kernels are FP32-ALU / latency bound, so the fraction is the honest distance to the HBM bound."""
    out = {}
    disc = model.discriminator
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    flush = torch.empty(192 * 1024 * 1024 // 4, device=dev)   # > 126 MB L2, rewritten between timed iterations
    for name, (B, L) in {"segment_B32_16384": (32, 16384), "long_B8_221184": (8, 864 * 256)}.items():
        g = torch.Generator().manual_seed(3)
        y = (torch.rand(B, L, generator=g) * 2 - 1).to(dev)
        xh = (y + 0.1 * torch.randn(B, L, generator=g).to(dev)).clamp(-1, 1).requires_grad_(True)

        def run(_i):
            xh.grad = None
            loss, _ = disc.forward_val(y, xh)
            loss.backward()

        for i in range(3):
            run(i)
        times = []
        for i in range(5):
            flush.fill_(float(i))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run(i)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        # 4 spectral passes (3 STFT resolutions + the mel STFT), forward + backward each re-read x_hat and y: 2 * 4 * 8 B per sample
        alg_bytes = 2 * 4 * 8.0 * B * L
        out[name] = {"ms_fwd_bwd": ms, "samples_per_s": B * L / (ms * 1e-3), "algorithmic_bytes": alg_bytes,
                     "achieved_gbs": alg_bytes / (ms * 1e-3) / 1e9, "hbm_peak_gbs": hbm, "frac_of_hbm_bound": alg_bytes / (ms * 1e-3) / 1e9 / hbm,
                     "l2": "192 MB buffer rewritten between timed iterations", "launch_mode": "eager (8 loss kernels + glue)"}
    return out


def run_ours(args, rank: int, local_rank: int, world: int):
    from optispeech_b200 import _lib
    from optispeech_b200.factory import DEFAULT_MODEL, build_model

    lib = _lib.load()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.check(lib.osb_check_device(local_rank), "osb_check_device")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()     # before any warm-up: the line carries samples from the warm-up and the timed region

    torch.manual_seed(SEED)
    model = build_model(DEFAULT_MODEL, train_args=dict(pretraining_steps=10 ** 9)).to(dev).train()
    host_batch = make_batch(B_PER_GPU, SEED + rank)
    frames_local = int(host_batch["mel_lengths"].sum())
    frames_all = frames_local
    if world > 1:
        t = torch.tensor([frames_local], device=dev, dtype=torch.int64)
        torch.distributed.all_reduce(t)
        frames_all = int(t.item())
    dev_batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in host_batch.items()}
    pinned = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in host_batch.items()}
    h2d = model.batch_h2d_bytes(model.stage_batch(pinned))  # what training_step really copies: the waveform crop is cut on the host

    def step_resident(i):
        model.training_step(dev_batch, i)

    # Every step's loss is read on the host inside the timed region, one step late: the 4-byte device->host copy of step i is
    # enqueued behind step i and waited for after step i+1 has been enqueued (the last step's is waited for at once), so the
    # host stages / uploads batch i+1 while the device still computes step i.
    loss_pin = [torch.zeros(1, pin_memory=True) for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"n": 0, "losses": []}

    def step_e2e(i, last=False):
        model.training_step(pinned, i)  # the batch is staged and copied host->device inside the step
        n = e2e_state["n"]
        loss_pin[n % 2].copy_(model.logged["total_loss/generator"].reshape(1), non_blocking=True)
        loss_ev[n % 2].record()
        if n > 0:
            loss_ev[(n - 1) % 2].synchronize()
            e2e_state["losses"].append(float(loss_pin[(n - 1) % 2]))
        if last:
            loss_ev[n % 2].synchronize()
            e2e_state["losses"].append(float(loss_pin[n % 2]))
            e2e_state["n"] = 0
        else:
            e2e_state["n"] = n + 1

    model.cuda_graph = not args.eager
    sampler.load_start = time.perf_counter()
    if model.cuda_graph:
        # set-up, outside warm-up and timing: 3 eager steps build buckets / tables, the 4th call captures the step's CUDA graph
        for i in range(4):
            step_resident(i)
    for i in range(args.warmup):
        step_resident(i)
    n0 = _lib.launch_count()
    ms = timed_steps(step_resident, args.steps, world, dev, sampler)
    if model.cuda_graph:  # replayed launches do not pass through the C-ABI counter: count the graph's library kernel nodes
        assert model._graphed is not None and model._graphed.replays >= args.steps
        launches = model._graphed.last_entry.launches
    else:
        launches = (_lib.launch_count() - n0) // max(args.steps, 1)
    clocks = sampler.stop() if rank == 0 else {}
    if args.quick:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": frames_all / (ms * 1e-3), "unit": UNIT,
                              "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                              "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches),
                              "note": "--quick run (possibly under a profiler): not a bench value"}), flush=True)
        if model._graphed is not None:
            model._graphed.release()
        return
    # untimed: a host batch is staged (the waveform crop is cut on the host), i.e. it has its own tensor signature and its own
    # captured graph: 3 eager steps + the capture + replays
    for i in range(4 + args.warmup):
        step_e2e(i, last=(i == 3 + args.warmup))
    e2e_state["losses"].clear()
    ms_e2e = timed_steps(lambda i: step_e2e(i, last=(i == args.steps - 1)), args.steps, world, dev)
    assert len(e2e_state["losses"]) == args.steps and all(math.isfinite(v) for v in e2e_state["losses"]), e2e_state["losses"]

    # ---- per-kernel device times of one step (separate pass; event recording perturbs the step time) ----
    roofline, top = None, []
    was_graphed, model.cuda_graph = model.cuda_graph, False  # per-launch events need the eager launches
    if rank == 0:
        # park the GPU behind a ~20 ms spin first: the host then runs ahead and the step's launches are all queued, so the
        # event brackets measure device execution only (an idle GPU would make every bracket include the host's launch cost)
        from optispeech_b200 import ops as _ops
        torch.cuda._sleep(40_000_000)
        _ops.SIDE_STREAMS_ENABLED = False   # serial execution: a bracket must hold one kernel, not its concurrent neighbours
        try:
            with _lib.LaunchProfiler() as prof:
                step_resident(0)
            # a bracket around ONE launch also holds the event records and the launch latency of an empty stream (~10 us on a
            # 15-20 us kernel): every contraction launch of the step is re-issued 8x back to back between one event pair
            prof.replay_amortized(repeat=8)
        finally:
            _ops.SIDE_STREAMS_ENABLED = True
    else:
        step_resident(0)  # the step holds a gradient all-reduce: every rank has to take it (rank 0 takes it serially, see above)
    model.cuda_graph = was_graphed
    if rank == 0:
        summ = prof.summary()
        # shares of the step are taken from the single-launch brackets (every kernel of the step has one); the device time of a
        # contraction launch from the amortized replay
        total_ms = sum(a["bracket_ms"] for a in summ) or 1.0
        top = [{"kernel": a["key"], "launches": a["launches"], "total_ms": round(a["total_ms"], 4), "share_of_lib_time": round(a["bracket_ms"] / total_ms, 4),
                "avg_us": round(a["avg_us"], 2), "avg_us_single_bracket": round(a["avg_us_single_bracket"], 2), "tflops": round(a["flops_per_launch"] / (a["avg_us"] * 1e-6) / 1e12, 2) if a["flops"] else None}
               for a in summ[:14]]
        # dominant kernel = the kernel FUNCTION (all of its shapes in the step together) with the largest device time
        groups = {}
        for a in summ:
            if a["flops"] > 0:
                g = groups.setdefault(a["key"].split("[")[0], {"ms": 0.0, "bracket_ms": 0.0, "flops": 0.0, "launches": 0, "shapes": []})
                g["ms"] += a["total_ms"]
                g["bracket_ms"] += a["bracket_ms"]
                g["flops"] += a["flops"]
                g["launches"] += a["launches"]
                g["shapes"].append({"shape": a["key"], "launches": a["launches"], "avg_us": round(a["avg_us"], 2),
                                    "tflops": round(a["flops_per_launch"] / (a["avg_us"] * 1e-6) / 1e12, 2)})
        if groups:
            name, g = max(groups.items(), key=lambda kv: kv[1]["ms"])
            # each launch is timed alone between its own events (serial pass): the burst figure is the honest denominator
            peak = float(peaks.get("bf16_tflops", 1590.0))
            achieved = g["flops"] / (g["ms"] * 1e-3) / 1e12  # algorithmic flops of all its launches / their summed duration
            roofline = {"bound": "tensor", "kernel": name, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                        "frac": achieved / peak, "traffic": ncu_traffic(name),
                        "peak_source": "MEASURED_PEAKS.json bf16_tflops, burst (of measured)" if peaks else "fallback 1590 burst (of fallback)",
                        "launches_per_step": g["launches"], "avg_us": 1e3 * g["ms"] / g["launches"], "share_of_lib_time": g["bracket_ms"] / total_ms,
                        "timing": "CUDA events on the launching stream; every contraction launch of the step re-issued 8x back to back "
                                  "(same arguments) between one event pair with the GPU parked behind a spin, / 8; a bracket around a "
                                  "single launch (avg_us_single_bracket in top_kernels) adds ~10 us of event / launch latency",
                        "per_shape": g["shapes"],
                        "all_tensor_kernels": {k: {"tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1), "ms": round(v["ms"], 4),
                                                   "launches": v["launches"]} for k, v in groups.items()},
                        "step_level": {"note": "algorithmic = the contraction FLOPs the REFERENCE executes for this step (SURVEY 8d); this "
                                               "implementation does not compute the upsampler / decoder frames no output of the step "
                                               "depends on (it runs them on segment + 24 of the 864 frames per sample), so it executes "
                                               "about 0.30 TFLOP of the 0.427",
                                       "algorithmic_tflop_per_step": T1_FLOP_PER_SAMPLE * B_PER_GPU / 1e12,
                                       "achieved_tflops": T1_FLOP_PER_SAMPLE * B_PER_GPU / (ms * 1e-3) / 1e12,
                                       "frac_of_sustained_peak": T1_FLOP_PER_SAMPLE * B_PER_GPU / (ms * 1e-3) / 1e12 / float(peaks.get("bf16_tflops_sustained", 1400.0))}}

    # ---- synthesis (SURVEY §8d S-synth-long: B=8 x 512 phonemes, and S-synth-1): device time -> roofline, wall time -> e2e ----
    synth = {}
    if rank == 0:
        model.eval()
        peak_t = float(peaks.get("bf16_tflops", 1590.0))
        for name in ("long_B8_Tx512", "single_B1_Tx120"):
            ids, lens, durs = synth_inputs(name)
            ids_pin = ids.pin_memory()
            for _ in range(3):
                out = model.generator.synthesise(ids_pin.to(dev, non_blocking=True), lens, durations=durs)
            torch.cuda.synchronize()
            reps, dev_ms = 5, []
            t0 = time.perf_counter()
            for _ in range(reps):
                out = model.generator.synthesise(ids_pin.to(dev, non_blocking=True), lens, durations=durs)  # returns CPU tensors
                dev_ms.append(out["latency"])
            dt = (time.perf_counter() - t0) / reps
            nsamp = int(out["wav_lengths"].sum())
            frames = nsamp // 256
            flops = FLOP_PER_FRAME_SYNTH * frames + FLOP_PER_PHONEME_SYNTH * int(lens.sum()) + 2.0 * ids.shape[1] * 256 * frames
            dms = float(np.median(dev_ms))
            synth[name] = {"audio_samples_per_s_e2e": nsamp / dt, "rtf": dt / (nsamp / 22050.0), "ms": 1e3 * dt, "frames": int(frames),
                           "precision": "fp16x3 (three tensor-core passes per contraction)", "device_ms": dms,
                           "audio_samples_per_s_device": nsamp / (dms * 1e-3),
                           "roofline": {"bound": "tensor", "algorithmic_gflop": flops / 1e9, "achieved": flops / (dms * 1e-3) / 1e12,
                                        "peak": peak_t, "unit": "TFLOP/s", "frac": flops / (dms * 1e-3) / 1e12 / peak_t,
                                        "hbm_floor_bytes": 1024 * frames, "d2h_bytes": int(nsamp * 4)}}
        model.train()
    if world > 1:
        # BASELINE config 5 ("long-form B=8, 512 phonemes, N x B200"): synthesis does not shard inside an utterance — REPLICAS ONLY
        # (utils/sharding.py deals whole utterances to the ranks).  Every rank synthesises its own B=8 x 512 batch at the same
        # time, no collective on the data path; aggregate = sum of the ranks' samples / the slowest rank's wall time per call.
        model.eval()
        ids, lens, durs = synth_inputs("long_B8_Tx512")
        ids_pin = ids.pin_memory()
        for _ in range(3):
            out = model.generator.synthesise(ids_pin.to(dev, non_blocking=True), lens, durations=durs)
        torch.cuda.synchronize()
        torch.distributed.barrier()
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            out = model.generator.synthesise(ids_pin.to(dev, non_blocking=True), lens, durations=durs)
        dt = (time.perf_counter() - t0) / reps
        stat = torch.tensor([dt, float(int(out["wav_lengths"].sum()))], device=dev, dtype=torch.float64)
        mx = stat.clone()
        torch.distributed.all_reduce(mx, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(stat, op=torch.distributed.ReduceOp.SUM)
        model.train()
        if rank == 0:
            synth["long_B8_Tx512_replicas"] = {
                "n_gpus": world, "utterances": 8 * world, "ms_slowest_rank": 1e3 * float(mx[0]), "audio_samples_per_s_e2e": float(stat[1]) / float(mx[0]),
                "vs_one_gpu": (float(stat[1]) / float(mx[0])) / synth["long_B8_Tx512"]["audio_samples_per_s_e2e"] if "long_B8_Tx512" in synth else None,
                "scaling": "replicas only: every rank synthesises its own B=8 x 512 batch concurrently, no collective on the data path"}

    # ---- BASELINE config 4: the same step and synthesis with the Transformer backbone (self-attention kernel path) ----
    variants = {}
    if rank == 0 and world == 1 and not args.no_variants:
        try:
            from optispeech_b200.factory import transformer_model_config

            torch.manual_seed(SEED)
            tmodel = build_model(transformer_model_config(), train_args=dict(pretraining_steps=10 ** 9)).to(dev).train()
            tmodel.cuda_graph = True
            tms = _event_time(lambda i: tmodel.training_step(dev_batch, i), n=10, warm=4 + 3)
            tmodel.eval()
            ids, lens, durs = synth_inputs("long_B8_Tx512")
            for _ in range(3):
                out = tmodel.generator.synthesise(ids.to(dev), lens, durations=durs)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                out = tmodel.generator.synthesise(ids.to(dev), lens, durations=durs)
            dt = (time.perf_counter() - t0) / 5
            nsamp = int(out["wav_lengths"].sum())
            variants["transformer_backbone"] = {
                "config": "configs/model/transformer.yaml (2 heads, d_k 128, 4+4 blocks), same batch, pre-training step, CUDA-graph replay",
                "ms_per_step": tms, "mel_frames_per_s": frames_local / (tms * 1e-3),
                "synthesis_long_B8_Tx512": {"audio_samples_per_s_e2e": nsamp / dt, "ms": 1e3 * dt, "frames": nsamp // 256}}
            if tmodel._graphed is not None:
                tmodel._graphed.release()
            del tmodel
        except Exception as exc:  # the headline line must not depend on a variant
            variants["transformer_backbone"] = {"error": repr(exc)[:300]}

    # ---- SURVEY §8(d) step variants on the same batch, both as CUDA-graph replays: T2 = generator step with the spectral losses
    #      through the vocoder; T3 = the full GAN-phase training_step (generator turn + discriminator turn) ----
    if rank == 0 and world == 1 and not args.no_variants:
        try:
            from optispeech_b200.model.graphed import GraphedTrainingStep

            torch.manual_seed(SEED)
            gmodel = build_model(DEFAULT_MODEL, train_args=dict(pretraining_steps=0)).to(dev).train()
            opt_g = gmodel.optimizers()[0]

            def t2_step(batch, i):   # forward, AM + 45 mel + 2.5 MR-STFT losses, backward through vocoder and acoustic model, AdamW
                from optispeech_b200 import ops as _ops
                _ops.step_counter(dev).add_(1)
                out = gmodel._process_batch(batch)
                spec_loss, _ = gmodel.discriminator.forward_val(out["wav"], out["wav_hat"])
                opt_g.zero_grad()
                gmodel.manual_backward(out["loss"] + spec_loss)
                gmodel.clip_gradients(opt_g, gradient_clip_val=10, gradient_clip_algorithm="norm")
                opt_g.step()

            t2_graph = GraphedTrainingStep(gmodel, step_fn=t2_step, train_discriminator=False)
            t2 = _event_time(lambda i: t2_graph(dev_batch, i), n=10, warm=4 + 3)
            variants["T2_generator_step_with_spectral_losses"] = {
                "ms_per_step": t2, "mel_frames_per_s": frames_local / (t2 * 1e-3), "launch_mode": "CUDA-graph replay",
                "library_launches": t2_graph.last_entry.launches if t2_graph.last_entry else None,
                "algorithmic_tflop": 16.48e9 * B_PER_GPU / 1e12, "achieved_tflops": 16.48e9 * B_PER_GPU / (t2 * 1e-3) / 1e12}
            t2_graph.release()
            gmodel.cuda_graph = True
            t3 = _event_time(lambda i: gmodel.training_step(dev_batch, i), n=10, warm=4 + 3)
            ent = gmodel._graphed.last_entry if gmodel._graphed is not None else None
            variants["T3_full_gan_training_step"] = {
                "ms_per_step": t3, "mel_frames_per_s": frames_local / (t3 * 1e-3), "launch_mode": "CUDA-graph replay",
                "library_launches": ent.launches if ent else None,
                "algorithmic_tflop": (78.8e9 + 124.4e9) * B_PER_GPU / 1e12,
                "achieved_tflops": (78.8e9 + 124.4e9) * B_PER_GPU / (t3 * 1e-3) / 1e12,
                "frac_of_sustained_peak": (78.8e9 + 124.4e9) * B_PER_GPU / (t3 * 1e-3) / 1e12 / float(peaks.get("bf16_tflops_sustained", 1400.0)),
                "discriminators": "period (MPD) and resolution (MRD) stacks on this package's kernels (disc/native.py: tcgen05 implicit "
                                  "GEMMs over the flat sequence layout); the MRD's rectangular-window |STFT| is torch.stft (cuFFT)"}
            if gmodel._graphed is not None:
                gmodel._graphed.release()
            variants["spectral_loss_kernels"] = spectral_kernel_bench(gmodel, dev, peaks)
            del gmodel
        except Exception as exc:
            variants["gan_phase"] = {"error": repr(exc)[:300]}

    if model._graphed is not None:  # every rank: graphs go before the process group does (and before the baselines below)
        model._graphed.release()

    # ---- baselines measured in the same run (rank 0, N=1 only), each in its own process: the reference owns the module name
    #      `optispeech` there.  CPU: the unmodified reference modules on the host cores, bounded sample; GPU: the same modules
    #      .cuda() = stock PyTorch kernels on this very B200 (the bar of SURVEY §2.1) ----
    cpu_baseline, gpu_library_baseline = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.cuda.empty_cache()
        ref_line = _run_sub(["--impl", "reference", "--steps", "2", "--warmup", "1"] + (["--no-variants"] if args.no_variants else []), 900)
        cpu_baseline = ref_line.get("cpu_baseline") or {"error": ref_line.get("error", "no line")}
        if isinstance(cpu_baseline, dict) and "sample" in cpu_baseline:
            cpu_baseline["sample"] += "; 2 timed steps after 1 warm-up (bounded sample)"
        for name, s in (ref_line.get("synthesis") or {}).items():
            if name in synth:
                synth[name]["cpu_baseline"] = s
        gpu_line = _run_sub(["--impl", "torch-gpu", "--steps", "5", "--warmup", "3"] + (["--no-variants"] if args.no_variants else []), 900)
        gpu_library_baseline = {k: gpu_line.get(k) for k in ("value", "ms_per_step", "variants", "note", "error") if k in gpu_line}
        gpu_library_baseline["unit"] = UNIT

    # ---- the gradient all-reduce on its own (N > 1): device time of the captured collective's payload, max over ranks ----
    allreduce = None
    if world > 1:
        bucket = model.optimizers()[0].buckets()[0]
        payload = torch.zeros_like(bucket.flat_g)
        for _ in range(3):
            torch.distributed.all_reduce(payload)
        torch.cuda.synchronize()
        torch.distributed.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(10):
            torch.distributed.all_reduce(payload)
        a1.record()
        torch.cuda.synchronize()
        ar_ms = max_over_ranks(a0.elapsed_time(a1), world, dev) / 10
        nbytes = payload.numel() * 4
        allreduce = {"bytes": nbytes, "ms": ar_ms, "bus_GBps": 2 * (world - 1) / world * nbytes / (ar_ms * 1e-3) / 1e9,
                     "note": "fp32 flat gradient bucket of the generator optimizer, one NCCL all-reduce per step inside the captured graph"}

    if rank == 0:
        value = frames_all / (ms * 1e-3)
        e2e_value = frames_all / (ms_e2e * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": workload_config(world),
            "impl_detail": {"precision": "fp16 tensor-core operands / fp32 accumulate, static loss scale 1024 (reference GPU default: 16-mixed)",
                            "l2": "no explicit flush: one step streams > 1 GB of activations through HBM (> 126 MB L2)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e},
            "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches),
            "launch_mode": "one CUDA-graph replay per step (graph holds the step's library + torch kernels)" if model.cuda_graph else "eager",
            "clocks": clocks,
            "allreduce": allreduce,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "gpu_library_baseline": gpu_library_baseline,
            "vs_torch_gpu": (e2e_value / gpu_library_baseline["value"]) if gpu_library_baseline and gpu_library_baseline.get("value") else None,
            "top_kernels": top,
            "synthesis": synth,
            "variants": variants,
        }
        if _STDOUT_FD is not None:
            sys.stdout.flush()
            os.dup2(_STDOUT_FD, 1)
        print(json.dumps(line), flush=True)


_STDOUT_FD = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-gpu"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU reference / stock-PyTorch-GPU baselines (sub-processes)")
    ap.add_argument("--no-variants", action="store_true", help="skip the Transformer / T2 / T3 / spectral variants")
    ap.add_argument("--eager", action="store_true", help="launch every kernel from the host instead of replaying the step's CUDA graph")
    ap.add_argument("--quick", action="store_true", help="timed region only (for runs under ncu): no e2e / synthesis / roofline pass / baselines")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, "cpu")
        return
    if args.impl == "torch-gpu":
        run_reference_arm(args, rank, "cuda")
        return
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version banner to file descriptor 1 at communicator creation; stdout carries ONE JSON line, so fd 1
        # points at stderr until the line is printed (run_ours restores it)
        global _STDOUT_FD
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, local_rank, world)
    except BaseException:
        import traceback

        traceback.print_exc()      # the multi-rank teardown below leaves through os._exit: the traceback has to be out first
        sys.stderr.flush()
        if world > 1:
            os._exit(1)
        raise
    finally:
        if world > 1:
            # The captured step holds NCCL kernels: its graphs must be gone before the communicator is torn down, and a
            # teardown that still blocks must not hold the launcher (the line is already printed): bounded by a watchdog.
            import gc

            gc.collect()
            torch.cuda.synchronize()
            sys.stdout.flush()
            threading.Timer(30.0, lambda: os._exit(0)).start()
            torch.distributed.destroy_process_group()
            os._exit(0)


if __name__ == "__main__":
    main()
