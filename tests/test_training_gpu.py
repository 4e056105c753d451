"""GPU parity of the training forward/backward (eval-mode numerics) against the CPU oracle, full-size ConvNeXt
configuration with deterministic weights.  fp16 tensor-core operands + fp32 accumulation: tolerances are those
of the reference's own `16-mixed` GPU path, stated per check."""
import numpy as np
import pytest
import torch

from oracle import losses as OL
from oracle import model as O
from oracle.spec import ModelSpec, deterministic_state_dict, generator_shapes

pytestmark = pytest.mark.gpu


def make_batch(spec, B, Tx, Tm, seed):
    g = torch.Generator().manual_seed(seed)
    x_lengths = torch.randint(Tx // 2, Tx + 1, (B,), generator=g)
    x_lengths[0] = Tx
    x = torch.randint(1, 159, (B, Tx), generator=g) * (torch.arange(Tx)[None] < x_lengths[:, None])
    mel_lengths = torch.clamp((x_lengths.float() * (Tm / Tx)).round().long(), max=Tm)
    mel_lengths[0] = Tm
    mmask = torch.arange(Tm)[None] < mel_lengths[:, None]
    mel = torch.randn(B, spec.n_feats, Tm, generator=g) * mmask[:, None, :]
    pitches = torch.randn(B, Tm, generator=g) * mmask
    energies = torch.randn(B, Tm, generator=g) * mmask
    wav = (torch.rand(B, Tm * spec.hop_length, generator=g) * 2 - 1).numpy().astype(np.float32)
    seg_rand = torch.rand(B, generator=g)
    return dict(x=x, x_lengths=x_lengths, mel=mel, mel_lengths=mel_lengths, pitches=pitches, energies=energies, wav=wav,
                seg_rand=seg_rand)


@pytest.fixture(scope="module")
def setup(cuda_device):
    from optispeech_b200.factory import build_generator, model_config_from_spec

    spec = ModelSpec()
    sd = deterministic_state_dict(generator_shapes(spec), seed=0, frames_per_token=3.0)
    gen = build_generator(model_config_from_spec(spec))
    gen.load_state_dict(sd, strict=True)
    gen = gen.to(cuda_device).eval()
    return spec, sd, gen


def test_mas_and_average_bit_exact(cuda_device):
    from optispeech_b200 import ops

    g = torch.Generator().manual_seed(11)
    B, Tm, Tx = 4, 150, 41
    tl = torch.tensor([41, 23, 1, 30])
    fl = torch.tensor([150, 77, 9, 30])
    lp = torch.log_softmax(torch.randn(B, Tm, Tx, generator=g) * 3, dim=-1)
    lp[3] = -1.0  # all ties
    path, ds = ops.mas(lp.to(cuda_device), tl.to(cuda_device), fl.to(cuda_device))
    ref_ds, _ = O.viterbi_decode(lp, tl, fl)
    assert torch.equal(ds.cpu(), ref_ds)
    for b in range(B):
        A = O.monotonic_alignment_search(lp[b, : fl[b], : tl[b]].numpy())
        assert np.array_equal(path[b, : fl[b]].cpu().numpy(), A)
    xs = torch.randn(B, Tm, generator=g)
    avg = ops.average_by_duration(ds, xs.to(cuda_device), tl.to(cuda_device), fl.to(cuda_device)).cpu()
    ref = O.average_by_duration(ref_ds, xs, tl, fl)
    assert (avg - ref).abs().max().item() <= 1e-6


@pytest.mark.parametrize("B,Tx,Tm", [(3, 48, 200)])
def test_training_forward_backward_matches_oracle(setup, cuda_device, B, Tx, Tm):
    spec, sd, gen = setup
    batch = make_batch(spec, B, Tx, Tm, seed=5)
    sd_ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = O.generator_forward(sd_ref, spec, batch["x"], batch["x_lengths"], batch["mel"], batch["mel_lengths"], batch["pitches"],
                              batch["energies"], batch["seg_rand"])
    ref["loss"].backward()

    from optispeech_b200.model.generator.training import generator_training_forward

    dev = cuda_device
    gen.zero_grad(set_to_none=True)
    out = generator_training_forward(gen, batch["x"].to(dev), batch["x_lengths"].to(dev), batch["mel"].to(dev),
                                     batch["mel_lengths"].to(dev), batch["pitches"].to(dev), batch["energies"].to(dev), None, None,
                                     seg_rand=batch["seg_rand"])
    aux = out["_aux"]
    # integer decisions: identical
    assert torch.equal(out["start_idx"].cpu(), ref["start_idx"])
    dur_equal = torch.equal(aux["durations"].cpu(), ref["durations"])
    lp_err = (aux["log_p_attn"].detach().cpu() - ref["log_p_attn"].detach())
    lp_err = lp_err[torch.isfinite(ref["log_p_attn"].detach())].abs().max().item()
    print(f"log_p_attn max-abs diff {lp_err:.3e}; durations identical: {dur_equal}")
    assert lp_err <= 5e-2
    assert dur_equal, "MAS durations differ from the oracle's"
    for key, tol in (("loss", 2e-3), ("align_loss", 2e-3), ("duration_loss", 2e-3), ("pitch_loss", 5e-3), ("energy_loss", 5e-3)):
        a, b = float(out[key].detach()), float(ref[key].detach())
        print(f"{key}: cuda {a:.6f} oracle {b:.6f}")
        assert abs(a - b) <= tol * max(1.0, abs(b)), key
    print("forwardsum", float(aux["forwardsum_loss"]), float(ref["forwardsum_loss"].detach()), "bin", float(aux["bin_loss"]),
          float(ref["bin_loss"].detach()))
    wav_err = (out["wav_hat"].detach().cpu() - ref["wav_hat"].detach()).abs().max().item()
    print(f"wav_hat max-abs diff (fp16 single-pass operands) {wav_err:.3e}")
    assert wav_err <= 1e-2

    for scale in (1024.0, 65536.0):
        gen.zero_grad(set_to_none=True)
        (out["loss"] * scale).backward(retain_graph=True)
        errs = []
        for name, p in gen.named_parameters():
            rg = sd_ref[name].grad
            if rg is None or float(rg.abs().max()) == 0.0 or p.grad is None:
                continue
            errs.append(float((p.grad.detach().cpu() / scale - rg).norm() / (rg.norm() + 1e-12)))
        e = torch.tensor(errs)
        print(f"loss scale {scale}: params {len(errs)} nan {int(torch.isnan(e).sum())} median rel err {float(e.nanmedian()):.3e} max {float(e[~torch.isnan(e)].max()):.3e}")
    gen.zero_grad(set_to_none=True)
    (out["loss"] * 1024.0).backward()   # static loss scale, removed below
    worst, bad = 0.0, []
    for name, p in gen.named_parameters():
        rg = sd_ref[name].grad
        if rg is None or float(rg.abs().max()) == 0.0:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, f"{name}: reference has no gradient here"
            continue
        assert p.grad is not None, name
        g = p.grad.detach().cpu() / 1024.0
        rel = float((g - rg).norm() / (rg.norm() + 1e-12))
        if not (rel <= 6e-2):
            bad.append((name, rel))
        else:
            worst = max(worst, rel)
    print(f"worst relative gradient error among passing params {worst:.3e}")
    for name, rel in bad:
        print(f"BAD {name}: relative gradient error {rel:.3e}")
    assert not bad


def test_windowed_decoder_equals_full_length_decoder(setup, cuda_device):
    """Training runs the Gaussian upsampler and the ConvNeXt decoder on the frames the vocoder segment depends on only
    (training.WINDOWED_DECODER: segment + 12 frames of context on either side, osb_gaussian_upsample_window).  The segment the
    vocoder receives — and with it wav_hat — must equal the one cut from the full-length computation, including segments at
    the very start / end of an utterance (window rows outside the sequence) and next to the padding of a short sample."""
    from optispeech_b200.model.generator import training
    from optispeech_b200.model.generator.training import generator_training_forward

    spec, sd, gen = setup
    dev = cuda_device
    B, Tx, Tm = 5, 48, 200
    batch = make_batch(spec, B, Tx, Tm, seed=9)
    # draws that put segments at frame 0, at the last admissible start, and in the middle
    batch["seg_rand"] = torch.tensor([0.0, 0.999999, 0.999999, 0.0, 0.73])   # sample 2 is full length: its window runs past Tm

    def run(windowed):
        training.WINDOWED_DECODER = windowed
        try:
            with torch.no_grad():
                out = generator_training_forward(gen, batch["x"].to(dev), batch["x_lengths"].to(dev), batch["mel"].to(dev),
                                                 batch["mel_lengths"].to(dev), batch["pitches"].to(dev), batch["energies"].to(dev),
                                                 None, None, seg_rand=batch["seg_rand"])
            torch.cuda.synchronize()
        finally:
            training.WINDOWED_DECODER = True
        return out

    full, win = run(False), run(True)
    assert torch.equal(full["start_idx"], win["start_idx"])
    assert full["_aux"]["decoder_out"].shape[1] == Tm and win["_aux"]["decoder_out"].shape[1] == gen.segment_size + 24
    # The decoder rows of the segment: the same per-row arithmetic.  Not bit-identical: with few row tiles the fused block splits
    # the intermediate dimension over CTAs and sums the partial results in a different order, and an fp32 ulp in a block's output
    # now and then flips the fp16 rounding of the next block's tensor-core operand (2^-11 of that element).  A wrong window edge
    # (padding, mask, offset) would show up as O(0.1 .. 1) errors in the samples whose segment touches the sequence ends.
    S = gen.segment_size
    errs = []
    for b in range(B):
        s = int(full["start_idx"][b])
        a = full["_aux"]["decoder_out"][b, s:s + S]
        w = win["_aux"]["decoder_out"][b, 12:12 + S]
        errs.append((a - w).abs().max().item() / max(1.0, a.abs().max().item()))
        print(f"sample {b}: start {s}, length {int(batch['mel_lengths'][b])}: decoder rows max diff {errs[-1]:.2e} of the largest value")
    assert max(errs) <= 2e-4, errs
    err = (full["wav_hat"] - win["wav_hat"]).abs().max().item()
    print(f"wav_hat: windowed vs full-length decoder, max-abs diff {err:.2e}")
    # the 5e-5 differences of the segment go through eight vocoder blocks on single-pass fp16 operands (measured 6.5e-4 .. 6.8e-4;
    # the same operands against the fp32 oracle: 2.6e-3 .. 3.7e-3); a wrong window edge is >= 1e-2
    assert err <= 3e-3
