"""GPU parity at BASELINE.json's own size (configs[1]: B=32, Tx=192, Tm=864) against fixtures generated from the REAL
reference modules (tests/golden/make_golden_fullsize.py -> tests/golden/fullsize_train.npz): the benchmark's batch, the
benchmark's weights, eval-mode numerics.

  * training forward + backward: losses, integer decisions (segment starts, MAS durations), averaged targets, waveform,
    every parameter gradient (norm; full value for small tensors; a strided slice of the large ones);
  * three `OptiSpeech.training_step`s (forward, backward, clip 10, FlatAdamW, cosine schedule): losses per step and the
    parameter deltas against three steps of the reference modules under torch.optim.AdamW.

Tolerances are those of fp16 tensor-core operands with fp32 accumulation (the reference's own GPU default is `16-mixed`),
stated per check.
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle.spec import ModelSpec, deterministic_state_dict, generator_shapes

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "fullsize_train.npz")


@pytest.fixture(scope="module")
def fx():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def batch(fx):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import bench

    assert (int(fx["B"]), int(fx["Tx"]), int(fx["Tm"]), int(fx["batch_seed"])) == (bench.B_PER_GPU, bench.TX, bench.TM, bench.SEED)
    b = bench.make_batch(bench.B_PER_GPU, bench.SEED)
    b["seg_rand"] = torch.from_numpy(fx["seg_rand"])
    return b


def _named(fx, prefix):
    keys = [str(k) for k in fx[f"{prefix}_keys"]]
    return keys, {k: float(n) for k, n in zip(keys, fx[f"{prefix}_norms"])}


def _family(k):
    return k.split(".")[0]


def _compare_packed(fx, prefix, tensors, scale, tol_full, tol_norm, what):
    """tensors: name -> CUDA tensor (divided by `scale` here).  Checks norms for all, values for the stored ones.
    `tol_full` is a float or a dict family -> tolerance (key "" = default)."""
    keys, norms = _named(fx, prefix)
    stride = 127
    bad, worst, fam_worst = [], 0.0, {}
    for k in keys:
        ref_norm = norms[k]
        t = tensors.get(k)
        if ref_norm <= 0.0:   # -1: the reference has no gradient here; 0: all-zero
            assert t is None or float(t.abs().max()) == 0.0, f"{what} {k}: reference has none"
            continue
        assert t is not None, f"{what} {k}: missing"
        flat = (t.detach().float().reshape(-1) / scale).cpu().numpy()
        rel_norm = abs(float(np.linalg.norm(flat.astype(np.float64))) - ref_norm) / ref_norm
        if f"{prefix}_full/{k}" in fx.files:
            ref = fx[f"{prefix}_full/{k}"]
        else:
            ref, flat = fx[f"{prefix}_slice/{k}"], flat[::stride]
        rel = float(np.linalg.norm(flat - ref) / (np.linalg.norm(ref) + 1e-30))
        worst = max(worst, rel)
        fam_worst[_family(k)] = max(fam_worst.get(_family(k), 0.0), rel)
        tol = tol_full.get(_family(k), tol_full[""]) if isinstance(tol_full, dict) else tol_full
        if not (rel <= tol and rel_norm <= tol_norm):
            bad.append((k, rel, rel_norm))
    print(f"{what}: {len(keys)} tensors, worst relative error {worst:.3e}; per family: "
          + ", ".join(f"{f} {v:.2e}" for f, v in sorted(fam_worst.items())))
    for k, rel, rn in bad:
        print(f"BAD {what} {k}: rel {rel:.3e} norm-rel {rn:.3e}")
    assert not bad


def test_training_forward_backward_fullsize_vs_reference(cuda_device, fx, batch):
    from optispeech_b200.factory import build_generator, model_config_from_spec
    from optispeech_b200.model.generator.training import generator_training_forward

    spec = ModelSpec()
    gen = build_generator(model_config_from_spec(spec))
    gen.load_state_dict(deterministic_state_dict(generator_shapes(spec), seed=0, frames_per_token=4.5), strict=True)
    gen = gen.to(cuda_device).eval()
    dev = cuda_device
    out = generator_training_forward(gen, batch["x"].to(dev), batch["x_lengths"].to(dev), batch["mel"].to(dev), batch["mel_lengths"].to(dev),
                                     batch["pitches"].to(dev), batch["energies"].to(dev), None, None, seg_rand=batch["seg_rand"])
    aux = out["_aux"]
    # integer decisions
    assert np.array_equal(out["start_idx"].cpu().numpy(), fx["start_idx"]), "segment start indices differ"
    ds = aux["durations"].cpu().numpy().astype(np.int64)
    ref_ds = fx["durations"].astype(np.int64)
    assert np.array_equal(ds.sum(1), ref_ds.sum(1)) and np.array_equal(ds.sum(1), batch["mel_lengths"].numpy())
    n_diff = int((ds != ref_ds).sum())
    print(f"MAS durations: {n_diff} of {ds.size} entries differ from the reference's numba search")
    # the search itself is bit-exact on identical inputs (test_training_gpu.py); here its input (log_p_attn) carries the
    # split-precision distance, so a handful of near-tie decisions out of 6144 may move by one frame
    assert n_diff <= ds.size // 500
    for key, tol in (("loss", 2e-3), ("align_loss", 2e-3), ("duration_loss", 2e-3), ("pitch_loss", 5e-3), ("energy_loss", 5e-3)):
        a, b = float(out[key].detach()), float(fx[key])
        print(f"{key}: cuda {a:.6f} reference {b:.6f}")
        assert abs(a - b) <= tol * max(1.0, abs(b)), key
    assert abs(float(aux["bin_loss"]) - float(fx["bin_loss"])) <= 2e-3 * max(1.0, abs(float(fx["bin_loss"])))
    if n_diff == 0:
        assert np.abs(aux["pitch_avg"].cpu().numpy() - fx["pitch_avg"]).max() <= 1e-5
        assert np.abs(aux["energy_avg"].cpu().numpy() - fx["energy_avg"]).max() <= 1e-5
    wav_err = np.abs(out["wav_hat"].detach().cpu().numpy()[:, ::32] - fx["wav_hat_s16"]).max()
    print(f"wav_hat max-abs diff (single-pass fp16 operands, training path) {wav_err:.3e}")
    assert wav_err <= 1e-2
    gen.zero_grad(set_to_none=True)
    (out["loss"] * 1024.0).backward()
    grads = {k: p.grad for k, p in gen.named_parameters()}
    # per layer family: ReLU gates of the 5-layer pitch predictor flip between an fp16-operand forward and the fp32 reference
    # (error grows with depth); smooth GELU blocks and the alignment convolutions agree far better (worst values are printed)
    # measured (round 2): alignment 3.8e-2, encoder 3.1e-2 (it inherits the pitch predictor's input gradient), text_embedding
    # 2.9e-2, energy 2.2e-2, duration 7e-3, pitch 5.2e-2; an isolated ConvNeXt block agrees to 5e-4 (test_autograd_fn_gpu.py)
    tol = {"": 5e-2, "pitch_predictor": 7e-2, "duration_predictor": 2e-2, "energy_predictor": 4e-2}
    _compare_packed(fx, "grad", grads, 1024.0, tol_full=tol, tol_norm=3e-2, what="gradient")


def test_three_training_steps_fullsize_vs_reference(cuda_device, fx, batch):
    """OptiSpeech.training_step x3 (eval-mode numerics, pre-training phase) against three reference steps."""
    from functools import partial

    from transformers import get_cosine_schedule_with_warmup

    from optispeech_b200.factory import build_model, model_config_from_spec

    spec = ModelSpec()
    model = build_model(model_config_from_spec(spec), train_args=dict(pretraining_steps=10 ** 9))
    model.hparams.scheduler = partial(get_cosine_schedule_with_warmup, num_warmup_steps=int(fx["warmup_steps"]), num_training_steps=-1)
    sd = deterministic_state_dict(generator_shapes(spec), seed=0, frames_per_token=4.5)
    model.generator.load_state_dict(sd, strict=True)
    model = model.to(cuda_device).eval()
    dev = cuda_device
    dbatch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
    before = {k: p.detach().clone() for k, p in model.generator.named_parameters()}
    losses = []
    for i in range(3):
        model.training_step(dbatch, i)
        losses.append(float(model.logged["total_loss/train_am_loss"]))
    print("step losses:", losses, "reference:", fx["step_losses"].tolist())
    for a, b in zip(losses, fx["step_losses"]):
        assert abs(a - b) <= 3e-3 * abs(b)
    deltas = {k: (p.detach() - before[k]) for k, p in model.generator.named_parameters()}
    # Adam's first updates are ~ lr * sign(g): elements whose gradient is at fp16 noise level flip sign, so the bound is on
    # the relative L2 error per tensor; tensors the reference never updates (decoder, energy embedding) must not move at all
    _compare_packed(fx, "delta", deltas, 1.0, tol_full=0.25, tol_norm=0.05, what="3-step delta")
    # the schedule itself
    sched = model.lr_schedulers()[0]
    assert abs(sched.get_last_lr()[0] - 2e-4) <= 1e-9 and model.optimizers()[0]._steps[0] == 3
