"""Generate golden vectors from the REAL reference modules (mush42/optispeech @ 3bdde20).

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference is imported as the top-level package `optispeech` from /root/reference with stub modules for
the packages that are not installed here (lightning, hydra, omegaconf, matplotlib — imported at module top
by optispeech/utils and optispeech/model/base_lightning_module.py but unused on this path).  Weights come
from oracle.spec.deterministic_state_dict, so the fixtures hold inputs/outputs only (plus the state_dict of
the tiny configuration).  Outputs: tests/golden/*.npz, tests/golden/state_dict_shapes.json.
"""
import importlib.machinery
import json
import os
import sys
import types
from functools import partial
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


sys.path.insert(0, ROOT)
from oracle.ref_harness import (build_reference_discriminator, build_reference_generator, install_stubs,  # noqa: E402,F401
                                run_reference_forward)
sys.path.remove(ROOT)


def train_batch(spec, B, Tx, Tm, seed):
    """Synthetic batch with the layout of TextWavBatchCollate."""
    g = torch.Generator().manual_seed(seed)
    x_lengths = torch.randint(Tx // 2, Tx + 1, (B,), generator=g)
    x_lengths[0] = Tx
    x = torch.randint(1, min(159, spec.n_vocab), (B, Tx), generator=g) * (torch.arange(Tx)[None] < x_lengths[:, None])
    mel_lengths = torch.clamp((x_lengths.float() * (Tm / Tx)).round().long(), max=Tm)
    mel_lengths[0] = Tm
    mmask = (torch.arange(Tm)[None] < mel_lengths[:, None])
    mel = torch.randn(B, spec.n_feats, Tm, generator=g) * mmask[:, None, :]
    pitches = torch.randn(B, Tm, generator=g) * mmask
    energies = torch.randn(B, Tm, generator=g) * mmask
    wav = (torch.rand(B, Tm * spec.hop_length, generator=g) * 2 - 1).numpy().astype(np.float32)
    seg_rand = torch.rand(B, generator=g)
    return dict(x=x, x_lengths=x_lengths, mel=mel, mel_lengths=mel_lengths, pitches=pitches, energies=energies, wav=wav,
                seg_rand=seg_rand)


def main():
    install_stubs()
    sys.path.insert(0, REF)
    sys.path.append(ROOT)  # for `oracle` only; the reference must win the name `optispeech`
    from oracle.spec import ModelSpec, deterministic_state_dict, generator_shapes, tiny_spec

    import optispeech  # noqa: F401  (the reference)
    assert optispeech.__file__.startswith(REF), optispeech.__file__
    from optispeech.model.generator.alignments import (GaussianUpsampling, _monotonic_alignment_search, average_by_duration,
                                                       expand_by_duration, viterbi_decode)
    from optispeech.model.generator.loss import ForwardSumLoss

    torch.manual_seed(1234)
    shapes_json = {}

    for name, spec, fpt, synth_shape, train_shape in (
        ("tiny", tiny_spec(), 3.0, (3, 23), (3, 20, 90)),
        ("full", ModelSpec(), 3.0, (2, 37), (2, 24, 110)),
    ):
        gen, fe = build_reference_generator(spec)
        ref_shapes = {k: tuple(v.shape) for k, v in gen.state_dict().items()}
        assert ref_shapes == {k: tuple(v) for k, v in generator_shapes(spec).items()}, "oracle/spec.py shape table is out of date"
        shapes_json[name] = {k: list(v) for k, v in ref_shapes.items()}
        sd = deterministic_state_dict(ref_shapes, seed=0, frames_per_token=fpt)
        gen.load_state_dict(sd, strict=True)
        gen.eval()

        # ---- synthesise ------------------------------------------------------------------
        B, Tx = synth_shape
        g = torch.Generator().manual_seed(99)
        x_lengths = torch.randint(Tx // 2, Tx + 1, (B,), generator=g)
        x_lengths[0] = Tx
        x = torch.randint(1, min(159, spec.n_vocab), (B, Tx), generator=g) * (torch.arange(Tx)[None] < x_lengths[:, None])
        out = gen.synthesise(x, x_lengths, d_factor=1.1, p_factor=1.6, e_factor=1.2)
        fx = dict(x=x.numpy(), x_lengths=x_lengths.numpy(), wav=out["wav"].numpy(), wav_lengths=out["wav_lengths"].numpy(),
                  durations=out["durations"].numpy(), pitch=out["pitch"].numpy(), energy=out["energy"].numpy())

        # ---- training forward + backward -----------------------------------------------
        B, Tx, Tm = train_shape
        batch = train_batch(spec, B, Tx, Tm, seed=7)
        gen.zero_grad()
        o = run_reference_forward(gen, batch)
        o["loss"].backward()
        fx.update({f"train_{k}": (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()})
        fx.update(train_wav_hat=o["wav_hat"].detach().numpy(), train_start_idx=o["start_idx"].numpy(),
                  train_loss=o["loss"].item(), train_align_loss=o["align_loss"].item(), train_duration_loss=o["duration_loss"].item(),
                  train_pitch_loss=o["pitch_loss"].item(), train_energy_loss=o["energy_loss"].item())
        grad_norms = {k: (float(p.grad.norm()) if p.grad is not None else -1.0) for k, p in gen.named_parameters()}
        fx["train_grad_keys"] = np.array(sorted(grad_norms))
        fx["train_grad_norms"] = np.array([grad_norms[k] for k in sorted(grad_norms)], dtype=np.float64)
        if name == "tiny":
            fx.update({f"sd/{k}": v.numpy() for k, v in sd.items()})
            fx.update({f"grad/{k}": p.grad.numpy() for k, p in gen.named_parameters() if p.grad is not None})

        # ---- discriminator-side losses (mel + MR-STFT, forward_val semantics) -------------------------
        disc = build_reference_discriminator(spec, fe)
        wav_hat = o["wav_hat"].detach().clone().requires_grad_(True)
        import optispeech.utils.segments as seg
        wav_gt = torch.from_numpy(seg.get_segments_numpy(batch["wav"][:, None, :], (o["start_idx"] * spec.hop_length).numpy(),
                                                         o["segment_size"] * spec.hop_length)[:, 0])
        mel_l = disc._get_mel_loss(wav_gt, wav_hat)
        sc, mag = disc.mr_stft_loss(wav_hat, wav_gt)
        (mel_l + (sc + mag) * spec.lambda_mr_stft).backward()
        fx.update(val_wav_gt=wav_gt.numpy(), val_mel_loss=mel_l.item(), val_sc_loss=sc.item(), val_mag_loss=mag.item(),
                  val_dwav_hat=wav_hat.grad.numpy(), mel_fb=disc.melspec_loss.mel_spec.mel_scale.fb.numpy())
        np.savez_compressed(os.path.join(HERE, f"generator_{name}.npz"), **fx)
        print(name, "synth frames", out["durations"].sum(1).tolist(), "loss", o["loss"].item(), "mel", mel_l.item(), sc.item(), mag.item())

    # ---- stand-alone algorithm fixtures (numba MAS, averaging, upsampling, expand, forward-sum, prior) ---------------
    g = torch.Generator().manual_seed(5)
    fx = {}
    mas_cases = []
    for i, (T, N) in enumerate([(1, 1), (7, 1), (5, 5), (40, 13), (120, 57), (64, 64)]):
        lp = torch.log_softmax(torch.randn(T, N, generator=g) * 2.0, dim=-1).numpy().astype(np.float32)
        if i == 3:
            lp[:, :] = -1.0  # all ties: exercises the >= tie-break
        A = _monotonic_alignment_search(lp)
        fx[f"mas{i}_lp"], fx[f"mas{i}_A"] = lp, np.asarray(A, dtype=np.int64)
        mas_cases.append(i)
    # viterbi_decode + bin loss + average_by_duration on a ragged batch
    B, Tm, Tx = 3, 50, 17
    tl = torch.tensor([17, 9, 12]); fl = torch.tensor([50, 31, 12])
    lpa = torch.log_softmax(torch.randn(B, Tm, Tx, generator=g), dim=-1)
    ds, bin_loss = viterbi_decode(lpa, tl, fl)
    xs = torch.randn(B, Tm, 1, generator=g)
    avg = average_by_duration(ds, xs, tl, fl)
    fx.update(vd_lp=lpa.numpy(), vd_tl=tl.numpy(), vd_fl=fl.numpy(), vd_ds=ds.numpy(), vd_bin=bin_loss.item(), avg_xs=xs.numpy(),
              avg_out=avg.numpy())
    fsl = ForwardSumLoss()(lpa, tl, fl)
    fx["fs_loss"] = fsl.item()
    # Gaussian upsampling (float and integer durations) and hard expansion
    hs = torch.randn(2, 9, 6, generator=g)
    d_int = torch.tensor([[2, 0, 3, 1, 4, 1, 1, 2, 3], [1, 5, 0, 2, 0, 0, 0, 0, 0]])
    d_mask = torch.tensor([[True] * 9, [True] * 4 + [False] * 5])
    ylen = d_int.sum(1)
    h_mask = torch.arange(int(ylen.max()))[None] < ylen[:, None]
    up = GaussianUpsampling()(hs, d_int.clone(), h_mask, d_mask)
    upf = GaussianUpsampling()(hs, d_int.float(), h_mask, d_mask)
    ex, exl = expand_by_duration(hs, d_int)
    fx.update(gu_hs=hs.numpy(), gu_d=d_int.numpy(), gu_dmask=d_mask.numpy(), gu_hmask=h_mask.numpy(), gu_out=up.numpy(),
              gu_out_float=upf.numpy(), ex_out=ex.numpy(), ex_len=exl.numpy())
    # beta-binomial prior (scipy) for a few (T, N)
    from scipy.stats import betabinom
    for T, N in [(5, 3), (31, 9), (110, 24)]:
        alpha = np.arange(1, T + 1, dtype=float)
        beta = np.array([T - t + 1 for t in alpha])
        fx[f"prior_{T}_{N}"] = betabinom.logpmf(np.arange(N)[..., None], N, alpha, beta).T  # (T, N)
    np.savez_compressed(os.path.join(HERE, "algorithms.npz"), **fx)
    with open(os.path.join(HERE, "state_dict_shapes.json"), "w") as f:
        json.dump(shapes_json, f, indent=0, sort_keys=True)
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
