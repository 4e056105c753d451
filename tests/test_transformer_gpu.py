"""GPU parity of the Transformer backbone path (BASELINE config 4): the fused attention kernels against plain fp32 torch, the
backbone forward / backward against golden vectors produced by the REAL reference (tests/golden/transformer.npz), and the
generator with Transformer encoder + decoder against the oracle.  All calls go through the C ABI."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import model as O
from oracle.spec import ModelSpec, deterministic_state_dict, generator_shapes

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))


def _ref_attention(qkv, lens, heads):
    """fp32 restatement of attention.py:84-125 on the (already fp16-rounded) q|k|v."""
    B, T, D3 = qkv.shape
    D = D3 // 3
    dk = D // heads
    q, k, v = (qkv[..., i * D:(i + 1) * D].view(B, T, heads, dk).transpose(1, 2) for i in range(3))
    s = q @ k.transpose(-1, -2) / math.sqrt(dk)
    masked = ~(torch.arange(T, device=qkv.device)[None] < lens[:, None])[:, None, None, :]
    s = s.masked_fill(masked, torch.finfo(s.dtype).min)
    p = torch.softmax(s, dim=-1).masked_fill(masked, 0.0)
    return (p @ v).transpose(1, 2).reshape(B, T, D)


@pytest.mark.parametrize("B,T,lens", [(2, 50, [50, 17]), (3, 128, [128, 1, 0]), (2, 300, [300, 129]), (2, 864, [864, 500])])
def test_mha_forward_backward_match_torch(cuda_device, B, T, lens):
    from optispeech_b200 import ops

    dev = cuda_device
    H, D = 2, 256
    g = torch.Generator().manual_seed(T)
    qkv = (torch.randn(B, T, 3 * D, generator=g) * 1.5).to(dev).half()
    lens_t = torch.tensor(lens, device=dev, dtype=torch.int64)
    ctx, rmax, rinv = ops.mha_fwd(qkv, H, lens_t, save_stats=True)
    ctx_split, _, _ = ops.mha_fwd(qkv, H, lens_t, split_out=True)
    x = qkv.float().requires_grad_(True)
    ref = _ref_attention(x, lens_t, H)
    err = float((ctx.float() - ref).abs().max())
    print(f"  fwd max-abs err {err:.3e} (ref max {float(ref.abs().max()):.2f})")
    assert err <= 4e-3
    both = ctx_split[..., :D].float() + ctx_split[..., D:].float()
    assert torch.equal(ctx_split[..., :D], ctx)
    assert float((both - ref).abs().max()) <= err + 1e-6          # hi + lo carries the fp32 result
    assert float(ctx[2].abs().max() if B > 2 else 0.0) == 0.0 or lens[2] != 0   # a sample without keys gives exactly 0
    w = torch.randn(B, T, D, generator=g).to(dev)
    (ref * w).sum().backward()
    dqkv = ops.mha_bwd(qkv, H, lens_t, ctx, w.half(), rmax, rinv)
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        r = rel(dqkv[..., sl], x.grad[..., sl])
        print(f"  {name}: rel err {r:.3e}")
        assert r <= 6e-3, name


def test_mha_dropout_mask_is_regenerated_in_backward(cuda_device):
    """With attention dropout the context is linear in V for a fixed mask: <d_ctx, ctx(V + dV) - ctx(V)> must equal <dV_grad, dV>."""
    from optispeech_b200 import ops

    dev = cuda_device
    B, T, H, D, p, seed = 2, 200, 2, 256, 0.3, 4242
    g = torch.Generator().manual_seed(1)
    qkv = torch.randn(B, T, 3 * D, generator=g).to(dev).half()
    lens = torch.tensor([200, 150], device=dev)
    ctx, rmax, rinv = ops.mha_fwd(qkv, H, lens, save_stats=True, dropout_p=p, dropout_seed=seed)
    ctx0, _, _ = ops.mha_fwd(qkv, H, lens)
    assert rel(ctx, ctx0) > 0.1, "dropout had no effect"
    assert torch.equal(ctx, ops.mha_fwd(qkv, H, lens, dropout_p=p, dropout_seed=seed)[0])
    w = torch.randn(B, T, D, generator=g).to(dev).half()
    dqkv = ops.mha_bwd(qkv, H, lens, ctx, w, rmax, rinv, dropout_p=p, dropout_seed=seed)
    # perturb V along the sign of the analytic gradient: the directional derivative is then large against the fp16 rounding
    # noise of the two contexts (a random direction gives |<g, dV>| ~ 2 with ~0.05 of noise, whatever the mask)
    dv = (0.25 * torch.sign(dqkv[..., 2 * D:].float())).half()
    qkv2 = qkv.clone()
    qkv2[..., 2 * D:] += dv
    dv_eff = (qkv2[..., 2 * D:].float() - qkv[..., 2 * D:].float())
    ctx2, _, _ = ops.mha_fwd(qkv2, H, lens, dropout_p=p, dropout_seed=seed)
    fd = float(((ctx2.float() - ctx.float()) * w.float()).sum())
    an = float((dqkv[..., 2 * D:].float() * dv_eff).sum())
    print(f"  dV directional: fd {fd:.4f} analytic {an:.4f}")
    assert abs(fd - an) <= 2e-2 * max(1.0, abs(an))


@pytest.fixture(scope="module")
def backbone(cuda_device):
    from optispeech_b200.factory import TRANSFORMER_BACKBONE
    from optispeech_b200.model.generator.modules import Transformer

    spec = ModelSpec(backbone="transformer")
    sd = deterministic_state_dict(generator_shapes(spec), seed=0, frames_per_token=3.0)
    bb = Transformer(dim=spec.dim, **TRANSFORMER_BACKBONE)
    bb.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=True)
    return bb.to(cuda_device).eval()


@pytest.mark.parametrize("split,tol", [(True, 2e-4), (False, 6e-3)])
def test_backbone_forward_matches_reference_golden(backbone, cuda_device, split, tol):
    fx = np.load(os.path.join(GOLD, "transformer.npz"))
    x = torch.from_numpy(fx["bb_x"]).to(cuda_device)
    lens = torch.from_numpy(fx["bb_lens"]).to(cuda_device)
    pad = ~(torch.arange(x.shape[1], device=cuda_device)[None] < lens[:, None])
    with torch.no_grad():
        out = backbone(x, pad, split=split)
    ref = torch.from_numpy(fx["bb_out"]).to(cuda_device)
    err = float((out - ref).abs().max())
    print(f"  split={split}: max-abs err {err:.3e} (ref max {float(ref.abs().max()):.2f})")
    assert err <= tol * max(1.0, float(ref.abs().max()))


def test_backbone_backward_matches_reference_golden(backbone, cuda_device):
    fx = np.load(os.path.join(GOLD, "transformer.npz"))
    x = torch.from_numpy(fx["bb_x"]).to(cuda_device).requires_grad_(True)
    w = torch.from_numpy(fx["bb_w"]).to(cuda_device)
    lens = torch.from_numpy(fx["bb_lens"]).to(cuda_device)
    pad = ~(torch.arange(x.shape[1], device=cuda_device)[None] < lens[:, None])
    backbone.zero_grad()
    out = backbone(x, pad)
    ref = torch.from_numpy(fx["bb_out"]).to(cuda_device)
    assert float((out - ref).abs().max()) <= 6e-3 * max(1.0, float(ref.abs().max()))
    (out * w).sum().backward()
    r = rel(x.grad, torch.from_numpy(fx["bb_dx"]).to(cuda_device))
    print(f"  dx rel err {r:.3e}")
    assert r <= 2e-2   # fp16 single-pass operands through four layers (the ConvNeXt training path shows the same 1e-2 level)
    params = dict(backbone.named_parameters())
    r = rel(params["transformer.encoders.0.self_attn.linear_q.weight"].grad, torch.from_numpy(fx["bb_grad_q0"]).to(cuda_device))
    print(f"  d linear_q.weight (layer 0) rel err {r:.3e}")
    assert r <= 2e-2
    worst = 0.0
    for k, ref_norm in zip(fx["bb_grad_keys"], fx["bb_grad_norms"]):
        gpar = params[str(k)].grad
        assert gpar is not None, k
        if ref_norm < 1e-4:     # linear_k.bias: mathematically zero gradient
            assert float(gpar.norm()) <= 5e-2, k
            continue
        if str(k).endswith("embed.0.alpha"):
            continue            # a scalar: checked below with its own (cancellation-aware) tolerance
        e = abs(float(gpar.norm()) - ref_norm) / ref_norm
        worst = max(worst, e)
        assert e <= 2e-2, (k, float(gpar.norm()), ref_norm)
    # d alpha = <dx, pe> is a sum of 38 400 signed terms that nearly cancel (|<dx,pe>| ~ 2e-3 |dx||pe|): the 1.2e-2 relative
    # rounding noise of dx shows up as ~ 1.2e-2 * |dx||pe| / sqrt(n) absolute, i.e. several percent of the small remainder
    dx_ref = torch.from_numpy(fx["bb_dx"])
    noise = 1.2e-2 * float(dx_ref.norm()) * math.sqrt(0.5)   # |pe element| rms = sqrt(1/2)
    got_alpha, ref_alpha = float(params["transformer.embed.0.alpha"].grad), float(fx["bb_grad_alpha"])
    print(f"  d alpha {got_alpha:.4f} (reference {ref_alpha:.4f}, rounding-noise scale {noise:.3f})")
    assert abs(got_alpha - ref_alpha) <= 4 * noise
    print(f"  worst parameter-gradient norm error {worst:.3e}")


@pytest.fixture(scope="module")
def tf_generator(cuda_device):
    from optispeech_b200.factory import build_generator, model_config_from_spec

    spec = ModelSpec(backbone="transformer")
    sd = deterministic_state_dict(generator_shapes(spec), seed=0, frames_per_token=3.0)
    gen = build_generator(model_config_from_spec(spec))
    gen.load_state_dict(sd, strict=True)
    return spec, sd, gen.to(cuda_device).eval()


def test_transformer_synthesise_matches_reference_golden(tf_generator, cuda_device):
    spec, sd, gen = tf_generator
    fx = np.load(os.path.join(GOLD, "transformer.npz"))
    x, xl = torch.from_numpy(fx["synth_x"]), torch.from_numpy(fx["synth_x_lengths"])
    out = gen.synthesise(x.to(cuda_device), xl, d_factor=1.1, p_factor=1.6, e_factor=1.2, durations=torch.from_numpy(fx["synth_durations"]))
    assert np.array_equal(out["wav_lengths"].numpy(), fx["synth_wav_lengths"])
    err = 0.0
    for b in range(x.shape[0]):
        n = int(fx["synth_wav_lengths"][b])
        err = max(err, float(np.abs(out["wav"][b, :n].numpy() - fx["synth_wav"][b, :n]).max()))
    print(f"  waveform max-abs diff vs the reference {err:.3e}")
    assert err <= 1e-3
    free = gen.synthesise(x.to(cuda_device), xl, d_factor=1.1, p_factor=1.6, e_factor=1.2)
    diff = np.abs(free["durations"].numpy() - fx["synth_durations"])
    assert diff.max() <= 1 and (diff != 0).mean() <= 0.02   # ceil() may flip by one frame at fp rounding distance
    assert float(np.abs(out["pitch"].numpy() - fx["synth_pitch"]).max()) <= 2e-3


def test_transformer_training_forward_backward_matches_reference_golden(tf_generator, cuda_device):
    spec, sd, gen = tf_generator
    fx = np.load(os.path.join(GOLD, "transformer.npz"))
    dev = cuda_device
    t = lambda k: torch.from_numpy(fx[k]).to(dev)  # noqa: E731
    from optispeech_b200.model.generator.training import generator_training_forward

    gen.zero_grad(set_to_none=True)
    out = generator_training_forward(gen, t("train_x"), t("train_x_lengths"), t("train_mel"), t("train_mel_lengths"), t("train_pitches"),
                                     t("train_energies"), None, None, seg_rand=torch.from_numpy(fx["train_seg_rand"]))
    assert np.array_equal(out["start_idx"].cpu().numpy(), fx["train_start_idx"])
    wav_err = float(np.abs(out["wav_hat"].detach().cpu().numpy() - fx["train_wav_hat"]).max())
    print(f"  wav_hat max-abs diff (fp16 single-pass operands) {wav_err:.3e}")
    assert wav_err <= 1e-2
    for key in ("loss", "align_loss", "duration_loss", "pitch_loss", "energy_loss"):
        ref = float(fx[f"train_{key}"])
        got = float(out[key])
        print(f"  {key}: {got:.5f} (reference {ref:.5f})")
        assert abs(got - ref) <= 3e-3 * max(1.0, abs(ref)), key
    (out["loss"] * 1024.0).backward()   # the training step's static loss scale
    params = dict(gen.named_parameters())
    errs = []
    for k, ref_norm in zip(fx["train_grad_keys"], fx["train_grad_norms"]):
        gpar = params[str(k)].grad
        if ref_norm < 0:
            assert gpar is None or float(gpar.abs().max()) == 0.0, k
            continue
        assert gpar is not None, k
        if ref_norm < 1e-4 or str(k).endswith("embed.0.alpha"):   # zero-gradient bias / near-cancelling scalar (see the backbone test)
            continue
        errs.append((abs(float(gpar.norm()) / 1024.0 - ref_norm) / ref_norm, str(k)))
    errs.sort()
    print(f"  parameter-gradient norms: median err {errs[len(errs) // 2][0]:.3e}, worst {errs[-1][0]:.3e} ({errs[-1][1]})")
    assert errs[-1][0] <= 6e-2


def test_transformer_training_step_eager_and_graphed(cuda_device):
    """OptiSpeech.training_step with the Transformer configuration: train mode (all dropouts on), pre-training phase, eager and
    CUDA-graph replay.  The encoder moves, the decoder (no gradient at the reference commit) does not."""
    from optispeech_b200.factory import build_model, model_config_from_spec

    spec = ModelSpec(backbone="transformer")
    torch.manual_seed(1234)
    model = build_model(model_config_from_spec(spec), train_args=dict(pretraining_steps=1000))
    model.generator.load_state_dict(deterministic_state_dict(generator_shapes(spec), seed=0, frames_per_token=3.0))
    model = model.to(cuda_device).train()
    g = torch.Generator().manual_seed(3)
    B, Tx, Tm = 2, 40, 170
    xl = torch.tensor([40, 29])
    x = torch.randint(1, 159, (B, Tx), generator=g) * (torch.arange(Tx)[None] < xl[:, None])
    ml = torch.tensor([170, 123])
    mm = torch.arange(Tm)[None] < ml[:, None]
    batch = dict(x=x, x_lengths=xl, mel=torch.randn(B, spec.n_feats, Tm, generator=g) * mm[:, None, :], mel_lengths=ml,
                 pitches=torch.randn(B, Tm, generator=g) * mm, energies=torch.randn(B, Tm, generator=g) * mm,
                 wav=(torch.rand(B, Tm * spec.hop_length, generator=g) * 2 - 1).numpy().astype(np.float32), sids=None, lids=None)
    enc = model.generator.encoder.transformer.encoders[0].self_attn.linear_q.weight
    dec = model.generator.decoder.transformer.encoders[0].self_attn.linear_q.weight
    enc0, dec0 = enc.detach().clone(), dec.detach().clone()
    losses = []
    for i in range(3):
        model.training_step(batch, i)
        losses.append(float(model.logged["total_loss/generator"]))
    model.cuda_graph = True
    for i in range(3, 9):   # 3 eager warm-up steps of the graphed wrapper, capture, 2 replays
        model.training_step(batch, i)
        losses.append(float(model.logged["total_loss/generator"]))
    print("  losses:", [round(v, 3) for v in losses])
    assert all(np.isfinite(v) for v in losses)
    assert model._graphed is not None and model._graphed.replays >= 2
    assert not torch.equal(enc0, enc.detach()) and torch.equal(dec0, dec.detach())
    # (no monotonic-loss check: nine steps into a 1000-step warm-up the learning rate is ~1e-6 and dropout noise dominates)
    model._graphed.release()
