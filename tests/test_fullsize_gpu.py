"""BASELINE.json's full sizes on the GPU, checked through size-independent properties (the CPU oracle would need minutes at
these shapes): batch / padding invariance of synthesis at the long-form shape (config 5: B=8, 512 phonemes, ~10 s), alignment
invariants and batch-permutation equivariance of the training forward at the bench shape (config 2: B=32, Tx=192, Tm=864), and
the fused attention on a long ragged sequence against fp32 torch.  One long-form utterance is also compared with the oracle."""
import math

import numpy as np
import pytest
import torch

from oracle import model as O
from oracle.spec import ModelSpec, deterministic_state_dict, generator_shapes

pytestmark = pytest.mark.gpu


def _generator(backbone, cuda_device):
    from optispeech_b200.factory import build_generator, model_config_from_spec

    spec = ModelSpec(backbone=backbone)
    sd = deterministic_state_dict(generator_shapes(spec), seed=0, frames_per_token=3.0)
    gen = build_generator(model_config_from_spec(spec))
    gen.load_state_dict(sd, strict=True)
    return spec, sd, gen.to(cuda_device).eval()


@pytest.fixture(scope="module")
def convnext(cuda_device):
    return _generator("convnext", cuda_device)


@pytest.fixture(scope="module")
def transformer(cuda_device):
    return _generator("transformer", cuda_device)


def _longform_inputs(B=8, Tx=512, seed=11):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(Tx // 2, Tx + 1, (B,), generator=g)
    lens[0] = Tx
    x = torch.randint(1, 159, (B, Tx), generator=g) * (torch.arange(Tx)[None] < lens[:, None])
    durs = torch.randint(1, 4, (B, Tx), generator=g) * (torch.arange(Tx)[None] < lens[:, None])   # ~2 frames / phoneme
    return x, lens, durs


@pytest.mark.parametrize("which", ["convnext", "transformer"])
def test_longform_synthesis_is_batch_and_padding_invariant(which, convnext, transformer, cuda_device):
    """Utterances are independent through the whole generator.  The longest utterance of the padded B=8 x 512-phoneme batch
    (no padding of its own) must come out exactly as when it is synthesised alone.  Shorter ones differ near their END in the
    reference too: its backbones apply the final LayerNorm to padded positions as well (LN(0) = bias != 0) and the predictor
    convolutions (receptive field 10 tokens for the 5-layer pitch predictor) read those positions, while a stand-alone run sees
    zero padding there; away from the end the predictions must agree."""
    spec, sd, gen = convnext if which == "convnext" else transformer
    x, lens, durs = _longform_inputs()
    out = gen.synthesise(x.to(cuda_device), lens, durations=durs)
    frames = durs.sum(1)
    assert torch.equal(out["wav_lengths"], frames * spec.hop_length) and int(frames.max()) > 900
    worst = 0.0
    for b in (0, 3, 7):
        n = int(lens[b])
        single = gen.synthesise(x[b: b + 1, :n].to(cuda_device), lens[b: b + 1], durations=durs[b: b + 1, :n])
        m = int(frames[b]) * spec.hop_length
        assert int(single["wav_lengths"][0]) == m
        if b == 0:
            worst = float((single["wav"][0, :m] - out["wav"][b, :m]).abs().max())
            assert float((single["pitch"][0, :n] - out["pitch"][b, :n]).abs().max()) <= 1e-4
        elif which == "convnext":   # local receptive fields only: away from the end nothing can see the padding
            assert float((single["pitch"][0, : n - 12] - out["pitch"][b, : n - 12]).abs().max()) <= 1e-4
        assert torch.isfinite(out["wav"][b, :m]).all() and float(out["wav"][b, :m].abs().max()) <= 1.0
    print(f"  [{which}] batch-vs-single waveform max-abs diff of the unpadded utterance {worst:.3e}")
    assert worst <= 2e-4


def test_longform_utterance_matches_oracle(convnext, cuda_device):
    spec, sd, gen = convnext
    x, lens, durs = _longform_inputs(B=1, Tx=512, seed=5)
    ref = O.synthesise(sd, spec, x, lens, 1.0, 1.0, 1.0, durations=durs)
    out = gen.synthesise(x.to(cuda_device), lens, durations=durs)
    n = int(ref["wav_lengths"][0])
    err = float((out["wav"][0, :n] - ref["wav"][0, :n]).abs().max())
    print(f"  512 phonemes / {n // spec.hop_length} frames: waveform max-abs diff vs oracle {err:.3e}")
    assert n // spec.hop_length > 900 and err <= 1e-3


def _bench_batch(spec, B=32, Tx=192, Tm=864, seed=1234):
    g = torch.Generator().manual_seed(seed)
    xl = torch.randint(Tx // 2, Tx + 1, (B,), generator=g); xl[0] = Tx
    x = torch.randint(1, 159, (B, Tx), generator=g) * (torch.arange(Tx)[None] < xl[:, None])
    ml = torch.clamp((4.5 * xl.float()).round().long(), max=Tm); ml[0] = Tm
    mm = torch.arange(Tm)[None] < ml[:, None]
    return dict(x=x, x_lengths=xl, mel=torch.randn(B, spec.n_feats, Tm, generator=g) * mm[:, None, :], mel_lengths=ml,
                pitches=torch.randn(B, Tm, generator=g) * mm, energies=torch.randn(B, Tm, generator=g) * mm,
                seg_rand=torch.rand(B, generator=g))


def test_bench_shape_training_forward_invariants(convnext, cuda_device):
    """B=32, Tx=192, Tm=864 (the bench batch): alignment invariants that hold at any size, and equivariance under a permutation
    of the batch (per-sample quantities follow their sample)."""
    from optispeech_b200.model.generator.training import generator_training_forward

    spec, sd, gen = convnext
    dev = cuda_device
    batch = _bench_batch(spec)

    def run(order):
        b = {k: v[order].to(dev) for k, v in batch.items() if k != "seg_rand"}
        with torch.no_grad():
            return generator_training_forward(gen, b["x"], b["x_lengths"], b["mel"], b["mel_lengths"], b["pitches"], b["energies"], None,
                                              None, seg_rand=batch["seg_rand"][order])

    ident = torch.arange(32)
    out = run(ident)
    aux = out["_aux"]
    dur = aux["durations"].cpu()
    xl, ml = batch["x_lengths"], batch["mel_lengths"]
    # monotonic alignment search: every frame of every utterance is assigned to exactly one real token, none to padding
    assert torch.equal(dur.sum(1).long(), ml)
    assert float((dur * (torch.arange(192)[None] >= xl[:, None])).abs().max()) == 0.0
    assert float(dur.min()) >= 0.0 and torch.equal(dur, dur.round())
    # attention log-probabilities: finite inside (frames x tokens) of each utterance, -inf at padded tokens
    lp = aux["log_p_attn"]
    for b in (0, 5, 31):
        blk = lp[b, : int(ml[b]), : int(xl[b])]
        assert torch.isfinite(blk).all()
        if int(xl[b]) < 192:
            assert torch.isinf(lp[b, : int(ml[b]), int(xl[b]):]).all()
    for key in ("loss", "align_loss", "duration_loss", "pitch_loss", "energy_loss"):
        assert math.isfinite(float(out[key])), key
    assert out["wav_hat"].shape == (32, spec.segment_size * spec.hop_length) and float(out["wav_hat"].abs().max()) <= 1.0
    # segment starts: floor(rand * max(len - 4 - 64, 0))
    exp_start = (batch["seg_rand"] * (ml.float() - 4 - 64).clamp(min=0)).long()
    assert torch.equal(out["start_idx"].cpu(), exp_start)

    perm = torch.randperm(32, generator=torch.Generator().manual_seed(2))
    out_p = run(perm)
    assert torch.equal(out_p["_aux"]["durations"].cpu(), dur[perm])
    assert float((out_p["_aux"]["pitch_avg"].cpu() - aux["pitch_avg"].cpu()[perm]).abs().max()) <= 1e-5
    # the segment vocoder runs on single fp16 operands and (2048 rows = 32 tiles) splits the intermediate dimension over CTAs
    # whose partial sums meet in L2 in arrival order: an fp32 last-bit difference can flip an fp16 rounding (2^-11 = 4.9e-4)
    assert float((out_p["wav_hat"].cpu() - out["wav_hat"].cpu()[perm]).abs().max()) <= 2e-3
    assert abs(float(out_p["align_loss"]) - float(out["align_loss"])) <= 1e-4 * abs(float(out["align_loss"]))


def test_attention_long_ragged_sequence(cuda_device):
    """11 key blocks, lengths that end inside a block / on a block boundary / at one key."""
    from optispeech_b200 import ops

    dev = cuda_device
    B, T, H, D = 3, 1300, 2, 256
    g = torch.Generator().manual_seed(9)
    qkv = torch.randn(B, T, 3 * D, generator=g).to(dev).half()
    lens = torch.tensor([1300, 1024, 1], device=dev)
    ctx, _, _ = ops.mha_fwd(qkv, H, lens)
    x = qkv.float()
    q, k, v = (x[..., i * D:(i + 1) * D].view(B, T, H, 128).transpose(1, 2) for i in range(3))
    s = q @ k.transpose(-1, -2) / math.sqrt(128)
    masked = ~(torch.arange(T, device=dev)[None] < lens[:, None])[:, None, None, :]
    p = torch.softmax(s.masked_fill(masked, torch.finfo(s.dtype).min), dim=-1).masked_fill(masked, 0.0)
    ref = (p @ v).transpose(1, 2).reshape(B, T, D)
    err = float((ctx.float() - ref).abs().max())
    print(f"  T=1300 ragged: max-abs err {err:.3e}")
    assert err <= 4e-3
    assert float((ctx[2].float() - x[2, :1, 2 * D:].expand(T, D)).abs().max()) <= 2e-3   # one key: the context is that key's value
