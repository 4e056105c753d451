"""Golden vectors at BASELINE.json's own size (configs[1]: B=32, Tx=192, Tm=864) from the REAL reference modules.

Run in the build container only (needs /root/reference or oracle/_ref):

    python tests/golden/make_golden_fullsize.py

The batch is bench.py's `make_batch(32, 1234)` (the benchmark's workload) plus a seeded segment draw; weights are
`oracle.spec.deterministic_state_dict(seed=0, frames_per_token=4.5)` (what bench.py's reference arm uses), eval mode
(dropout / DropPath off — they cannot be bit-matched).  Stored in tests/golden/fullsize_train.npz (a few hundred KB):

  * forward: the five losses, `start_idx`, MAS durations (from the reference's numba search), the duration-averaged pitch /
    energy targets, `wav_hat[:, ::32]` (key wav_hat_s16);
  * backward: every parameter's gradient norm, full gradients of the small tensors (<= 512 elements), a strided slice of
    the large ones;
  * three `training_step`s (base_lightning_module.py:78-110 restated in oracle/ref_harness.reference_training_step:
    clip 10, AdamW(2e-4, (0.8, 0.99), wd 1e-2), cosine schedule with the warm-up shortened from 1000 to 2 steps so that the
    updates are well above fp32 resolution: lr = 0, 1e-4, 2e-4): per-step losses and the parameter deltas in the same
    norm / small-full / slice form.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_harness as RH  # noqa: E402

sys.path.remove(ROOT)

SMALL = 512
WARMUP = 2   # cosine warm-up steps of the three-step fixture (the reference's 1000 would leave lr <= 4e-7: deltas at fp32 noise level)
STRIDE = 127


def pack_tensors(prefix, named, fx):
    keys = sorted(named)
    fx[f"{prefix}_keys"] = np.array(keys)
    fx[f"{prefix}_norms"] = np.array([float(named[k].double().norm()) if named[k] is not None else -1.0 for k in keys], dtype=np.float64)
    for k in keys:
        t = named[k]
        if t is None:
            continue
        flat = t.detach().reshape(-1).float().numpy()
        if flat.size <= SMALL:
            fx[f"{prefix}_full/{k}"] = flat
        else:
            fx[f"{prefix}_slice/{k}"] = flat[::STRIDE].copy()


def main():
    RH.import_reference()
    import optispeech.model.generator as G
    from oracle.spec import ModelSpec, deterministic_state_dict, generator_shapes

    import bench  # noqa: E402  (make_batch only; the reference already owns the name `optispeech`)

    torch.manual_seed(1234)
    torch.set_num_threads(os.cpu_count() or 1)
    spec = ModelSpec()
    gen, _ = RH.build_reference_generator(spec)
    sd = deterministic_state_dict(generator_shapes(spec), seed=0, frames_per_token=4.5)
    gen.load_state_dict(sd, strict=True)
    gen.eval()
    batch = bench.make_batch(bench.B_PER_GPU, bench.SEED)
    batch["seg_rand"] = torch.rand(bench.B_PER_GPU, generator=torch.Generator().manual_seed(4321))

    captured = {}
    orig_vd, orig_avg = G.viterbi_decode, G.average_by_duration

    def vd(*a, **k):
        ds, bl = orig_vd(*a, **k)
        captured["durations"] = ds.detach().clone()
        captured["bin_loss"] = float(bl)
        return ds, bl

    def avg(*a, **k):
        out = orig_avg(*a, **k)
        captured.setdefault("avgs", []).append(out.detach().clone())
        return out

    G.viterbi_decode, G.average_by_duration = vd, avg
    fx = {"warmup_steps": WARMUP, "seg_rand": batch["seg_rand"].numpy(), "batch_seed": bench.SEED, "B": bench.B_PER_GPU, "Tx": bench.TX, "Tm": bench.TM}
    try:
        t0 = time.time()
        gen.zero_grad()
        o = RH.run_reference_forward(gen, batch)
        o["loss"].backward()
        print(f"forward+backward {time.time() - t0:.1f}s  loss {o['loss'].item():.6f}")
    finally:
        G.viterbi_decode, G.average_by_duration = orig_vd, orig_avg
    fx.update(loss=o["loss"].item(), align_loss=o["align_loss"].item(), duration_loss=o["duration_loss"].item(),
              pitch_loss=o["pitch_loss"].item(), energy_loss=o["energy_loss"].item(), bin_loss=captured["bin_loss"],
              start_idx=o["start_idx"].numpy().astype(np.int64), durations=captured["durations"].numpy().astype(np.int16),
              pitch_avg=captured["avgs"][0].reshape(bench.B_PER_GPU, -1).numpy(), energy_avg=captured["avgs"][1].reshape(bench.B_PER_GPU, -1).numpy(),
              wav_hat_s16=o["wav_hat"].detach()[:, ::32].numpy())
    pack_tensors("grad", {k: p.grad for k, p in gen.named_parameters()}, fx)

    # ---- three training steps (fresh weights) ----
    gen.load_state_dict(sd, strict=True)
    gen.zero_grad(set_to_none=True)
    opt, sched = RH.reference_optimizer(gen, warmup=WARMUP)
    before = {k: p.detach().clone() for k, p in gen.named_parameters()}
    losses = []
    for i in range(3):
        t0 = time.time()
        out = RH.reference_training_step(gen, opt, sched, batch)
        losses.append(out["loss"].item())
        print(f"step {i}: loss {losses[-1]:.6f}  lr-after {sched.get_last_lr()[0]:.3e}  {time.time() - t0:.1f}s")
    fx["step_losses"] = np.array(losses, dtype=np.float64)
    pack_tensors("delta", {k: (p.detach() - before[k]) for k, p in gen.named_parameters()}, fx)
    path = os.path.join(HERE, "fullsize_train.npz")
    np.savez_compressed(path, **fx)
    print("wrote", path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
