"""Period discriminators on the native kernels (disc/native.py) against the same modules on stock PyTorch (fp32, TF32 off):
scores, every feature map, the generator-turn gradient d loss / d wav_hat and the discriminator-turn weight gradients.
Reference: vocoder/wavenext/disc/_discriminators.py:41-97, disc/loss.py:11-85."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _disc(period, dev, seed=0):
    from optispeech_b200.model.vocoder.wavenext.disc._discriminators import DiscriminatorP

    torch.manual_seed(seed)
    d = DiscriminatorP(period=period)
    with torch.no_grad():   # non-trivial weight_g so that the weight-norm backward is exercised
        for conv in list(d.convs) + [d.conv_post]:
            conv.weight_g.mul_(1.0 + 0.3 * torch.rand_like(conv.weight_g))
            conv.bias.add_(0.05 * torch.randn_like(conv.bias))
    return d.to(dev)


def _torch_path(fn):
    from optispeech_b200.model.vocoder.wavenext.disc import _discriminators as D

    old, old_r, old_tf32 = D.NATIVE_MPD, D.NATIVE_MRD, torch.backends.cudnn.allow_tf32
    D.NATIVE_MPD, D.NATIVE_MRD, torch.backends.cudnn.allow_tf32 = False, False, False
    try:
        return fn()
    finally:
        D.NATIVE_MPD, D.NATIVE_MRD, torch.backends.cudnn.allow_tf32 = old, old_r, old_tf32


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize("period,B,T", [(2, 2, 16384), (3, 2, 16384), (5, 3, 8000), (7, 2, 16384), (11, 2, 16384), (11, 1, 1201)])
def test_period_forward_matches_torch(cuda_device, period, B, T):
    from optispeech_b200.model.vocoder.wavenext.disc.native import FlatMap

    d = _disc(period, cuda_device)
    g = torch.Generator().manual_seed(period * 1000 + T)
    wav = (torch.rand(B, T, generator=g) * 2 - 1).to(cuda_device)
    with torch.no_grad():
        score, fmap = d(wav)
        score_ref, fmap_ref = _torch_path(lambda: d(wav))
    assert score.shape == score_ref.shape, (score.shape, score_ref.shape)
    assert len(fmap) == len(fmap_ref) == 5
    worst = _rel(score, score_ref)
    for m, r in zip(fmap, fmap_ref):
        dense = m.dense() if isinstance(m, FlatMap) else m
        assert dense.shape == r.shape, (dense.shape, r.shape)
        worst = max(worst, _rel(dense, r))
        if isinstance(m, FlatMap):   # the gap rows are exact zeros (they are the next layer's padding)
            rows, Cc = m.data.shape
            gap = m.data.view(rows // m.P, m.P, Cc)[:, m.L:]
            assert float(gap.abs().max()) == 0.0
    print(f"  period {period} T {T}: worst relative L2 difference over scores and feature maps {worst:.3e}")
    assert worst <= 3e-3     # fp16 operands and fp16 activations, fp32 accumulation


def _param_rel(a, b):
    """Relative L2 difference; gradients that cancel to (almost) nothing in the reference (the conv_post bias under a hinge
    loss with every term active: -1 + 1) are compared on an absolute scale."""
    nb = float(b.double().norm())
    if nb < 1e-6:
        return float((a.double() - b.double()).norm()) / 1e-3
    return _rel(a, b)


@pytest.mark.parametrize("period,T", [(2, 16384), (5, 8000), (11, 16384)])
def test_period_generator_turn_gradient(cuda_device, period, T):
    """Generator turn (weights frozen, loss scale 1024 as in training): d loss / d wav_hat for the hinge term plus a SMOOTH
    feature distance (mean squared difference of every feature map).  The L1 feature-matching gradient itself is a sign
    pattern: a rounding-level change of a feature flips entries, so it is compared separately and loosely below, and its
    kernels exactly in test_l1_pair_kernels."""
    from optispeech_b200.model.vocoder.wavenext.disc.loss import FeatureMatchingLoss, GeneratorLoss
    from optispeech_b200.model.vocoder.wavenext.disc.native import FlatMap, period_forward_pair

    d = _disc(period, cuda_device).requires_grad_(False)
    g = torch.Generator().manual_seed(period + T)
    wav = (torch.rand(2, T, generator=g) * 2 - 1).to(cuda_device)
    wav_hat0 = (0.7 * wav.cpu() + 0.3 * (torch.rand(2, T, generator=g) * 2 - 1)).to(cuda_device)

    def dense(m):
        return m.dense() if isinstance(m, FlatMap) else m

    def loss_of(native, smooth):
        wh = wav_hat0.clone().requires_grad_(True)
        if native:
            _, sg, fr, fg = period_forward_pair(d, wav, wh)
        else:
            (_, fr), (sg, fg) = d(wav), d(wh)
        if smooth:
            fm = sum(((dense(a).detach() - dense(b)) ** 2).mean() for a, b in zip(fr, fg)) * 10.0
        else:
            fm = FeatureMatchingLoss()([fr], [fg])
        loss = GeneratorLoss()([sg])[0] + fm
        (loss * 1024.0).backward()
        return float(loss), wh.grad / 1024.0

    for smooth, tol in ((True, 1e-2), (False, 0.15)):
        ln, gn = loss_of(True, smooth)
        lr, gr = _torch_path(lambda: loss_of(False, smooth))
        rel = _rel(gn, gr)
        cos = float((gn.double() * gr.double()).sum() / (gn.double().norm() * gr.double().norm()))
        print(f"  period {period} {'smooth' if smooth else 'L1'} feature term: loss {ln:.6f} vs {lr:.6f}; d loss / d wav_hat relative L2 "
              f"difference {rel:.3e}, cosine {cos:.6f}")
        assert abs(ln - lr) <= 2e-3 * abs(lr)
        assert rel <= tol


def test_l1_pair_kernels(cuda_device):
    """mean |a - b| and coef * sign(b - a) / n on fp16 maps (disc/loss.py:67-85) against torch on the same fp16 values."""
    from optispeech_b200 import ops

    g = torch.Generator().manual_seed(3)
    a = torch.randn(1000, 128, generator=g).half().to(cuda_device)
    b = (a.cpu().float() + 0.1 * torch.randn(1000, 128, generator=g)).half().to(cuda_device)
    b[::7] = a[::7]                                    # ties: sign(0) = 0
    s = ops.l1_pair_fwd(a, b)
    ref = (a.float() - b.float()).abs().sum()
    assert abs(float(s[0]) - float(ref)) <= 1e-5 * float(ref)
    coef = torch.tensor([3.0], device=cuda_device)
    db = ops.l1_pair_bwd(a, b, coef, 0.25)
    want = (0.75 * torch.sign(b.float() - a.float())).half()
    assert torch.equal(db, want)


@pytest.mark.parametrize("period,T", [(3, 16384), (7, 8000)])
def test_period_discriminator_turn_weight_gradients(cuda_device, period, T):
    """Hinge discriminator loss over (real, generated): gradients of every weight_g / weight_v / bias."""
    from optispeech_b200.model.vocoder.wavenext.disc.loss import DiscriminatorLoss
    from optispeech_b200.model.vocoder.wavenext.disc.native import period_forward_pair

    d = _disc(period, cuda_device)
    g = torch.Generator().manual_seed(period + T)
    wav = (torch.rand(3, T, generator=g) * 2 - 1).to(cuda_device)
    wav_hat = (0.5 * wav.cpu() + 0.5 * (torch.rand(3, T, generator=g) * 2 - 1)).to(cuda_device)

    def grads_of(native):
        d.zero_grad(set_to_none=True)
        if native:
            sr, sg, _, _ = period_forward_pair(d, wav, wav_hat)
        else:
            (sr, _), (sg, _) = d(wav), d(wav_hat)
        # hinge loss, with the real term weighted 1.75x: the plain sum is -1/n on the real rows and +1/n on the generated ones,
        # and with white-noise signals every parameter gradient would be a difference of two nearly equal sums
        loss = DiscriminatorLoss()([sr], [sg])[0] + 0.75 * torch.clamp(1 - sr, min=0).mean()
        (loss * 1024.0).backward()
        return float(loss), {n: p.grad.detach().clone() / 1024.0 for n, p in d.named_parameters()}

    ln, gn = grads_of(True)
    lr, gr = _torch_path(lambda: grads_of(False))
    assert abs(ln - lr) <= 2e-3 * abs(lr)
    # fp16 gradient rows (5e-4 per element) through five layers, sums of ~1e5 terms of mixed sign: 3e-2
    report = []
    for n in gr:
        r = _param_rel(gn[n], gr[n])
        report.append(f"{n} {r:.2e}")
        assert r <= 3e-2, (n, r, report)
    print(f"  period {period}: loss {ln:.6f} vs {lr:.6f}; parameter-gradient relative L2 differences: " + ", ".join(report))


# --------------------------------------------------------------------------------------------------
# resolution discriminators (reference _discriminators.py:139-216)
# --------------------------------------------------------------------------------------------------
def _disc_r(resolution, dev, seed=0):
    from optispeech_b200.model.vocoder.wavenext.disc._discriminators import DiscriminatorR

    torch.manual_seed(seed)
    d = DiscriminatorR(resolution=resolution)
    with torch.no_grad():
        for conv in list(d.convs) + [d.conv_post]:
            conv.weight_g.mul_(1.0 + 0.3 * torch.rand_like(conv.weight_g))
            conv.bias.add_(0.05 * torch.randn_like(conv.bias))
    return d.to(dev)


@pytest.mark.parametrize("resolution,B,T", [((1024, 256, 1024), 2, 16384), ((2048, 512, 2048), 2, 16384), ((512, 128, 512), 3, 16384),
                                            ((512, 128, 512), 1, 5000)])
def test_resolution_forward_matches_torch(cuda_device, resolution, B, T):
    from optispeech_b200.model.vocoder.wavenext.disc.native import FlatMap

    d = _disc_r(resolution, cuda_device)
    g = torch.Generator().manual_seed(resolution[0] + T)
    wav = (torch.rand(B, T, generator=g) * 2 - 1).to(cuda_device)
    with torch.no_grad():
        score, fmap = d(wav)
        score_ref, fmap_ref = _torch_path(lambda: d(wav))
    assert score.shape == score_ref.shape, (score.shape, score_ref.shape)
    assert len(fmap) == len(fmap_ref) == 6
    worst = _rel(score, score_ref)
    for m, r in zip(fmap, fmap_ref):
        dense = m.dense() if isinstance(m, FlatMap) else m
        assert dense.shape == r.shape, (dense.shape, r.shape)
        worst = max(worst, _rel(dense, r))
        if isinstance(m, FlatMap):
            rows, Cc = m.data.shape
            assert float(m.data.view(rows // m.P, m.P, Cc)[:, m.L:].abs().max()) == 0.0
    print(f"  resolution {resolution} T {T}: worst relative L2 difference over scores and feature maps {worst:.3e}")
    assert worst <= 3e-3


@pytest.mark.parametrize("resolution", [(1024, 256, 1024), (512, 128, 512)])
def test_resolution_gradients(cuda_device, resolution):
    """Generator turn: d loss / d wav_hat through the stack and torch.stft (hinge + smooth feature distance); discriminator
    turn: every weight_g / weight_v / bias gradient of the hinge loss."""
    from optispeech_b200.model.vocoder.wavenext.disc.loss import DiscriminatorLoss, GeneratorLoss
    from optispeech_b200.model.vocoder.wavenext.disc.native import FlatMap, resolution_forward_pair

    d = _disc_r(resolution, cuda_device)
    g = torch.Generator().manual_seed(resolution[1])
    T = 16384
    wav = (torch.rand(2, T, generator=g) * 2 - 1).to(cuda_device)
    wav_hat0 = (0.6 * wav.cpu() + 0.4 * (torch.rand(2, T, generator=g) * 2 - 1)).to(cuda_device)

    def dense(m):
        return m.dense() if isinstance(m, FlatMap) else m

    def gen_turn(native):
        d.requires_grad_(False)
        wh = wav_hat0.clone().requires_grad_(True)
        if native:
            _, sg, fr, fg = resolution_forward_pair(d, wav, wh)
        else:
            (_, fr), (sg, fg) = d(wav), d(wh)
        loss = GeneratorLoss()([sg])[0] + sum(((dense(a).detach() - dense(b)) ** 2).mean() for a, b in zip(fr, fg)) * 10.0
        (loss * 1024.0).backward()
        d.requires_grad_(True)
        return float(loss), wh.grad / 1024.0

    ln, gn = gen_turn(True)
    lr, gr = _torch_path(lambda: gen_turn(False))
    rel = _rel(gn, gr)
    print(f"  resolution {resolution}: generator turn loss {ln:.6f} vs {lr:.6f}; d loss / d wav_hat relative L2 difference {rel:.3e}")
    assert abs(ln - lr) <= 2e-3 * abs(lr) and rel <= 2e-2

    def disc_turn(native):
        d.zero_grad(set_to_none=True)
        if native:
            sr, sg, _, _ = resolution_forward_pair(d, wav, wav_hat0)
        else:
            (sr, _), (sg, _) = d(wav), d(wav_hat0)
        loss = DiscriminatorLoss()([sr], [sg])[0] + 0.75 * torch.clamp(1 - sr, min=0).mean()   # see the period test
        (loss * 1024.0).backward()
        return float(loss), {n: p.grad.detach().clone() / 1024.0 for n, p in d.named_parameters()}

    ln, gn = disc_turn(True)
    lr, gr = _torch_path(lambda: disc_turn(False))
    assert abs(ln - lr) <= 2e-3 * abs(lr)
    report = []
    for n in gr:
        r = _param_rel(gn[n], gr[n])
        report.append(f"{n} {r:.2e}")
        assert r <= 3e-2, (n, r, report)
    print(f"  resolution {resolution}: discriminator turn loss {ln:.6f} vs {lr:.6f}; parameter-gradient relative L2 differences: " + ", ".join(report))


@pytest.mark.parametrize("kind", ["period", "resolution"])
def test_discriminator_turn_reuses_generator_turn_activations(cuda_device, kind):
    """With `cache_generator_outputs` the discriminator turn sees the waveforms and weights of the generator turn; the layer
    outputs computed there are reused (native._pair).  Same loss and same weight gradients as a recomputed forward; a changed
    input or weight must NOT hit the cache."""
    from optispeech_b200.model.vocoder.wavenext.disc import native
    from optispeech_b200.model.vocoder.wavenext.disc.loss import DiscriminatorLoss, FeatureMatchingLoss, GeneratorLoss

    d = _disc(5, cuda_device) if kind == "period" else _disc_r((1024, 256, 1024), cuda_device)
    pair = native.period_forward_pair if kind == "period" else native.resolution_forward_pair
    g = torch.Generator().manual_seed(11)
    wav = (torch.rand(2, 16384, generator=g) * 2 - 1).to(cuda_device)
    wav_hat = (0.6 * wav.cpu() + 0.4 * (torch.rand(2, 16384, generator=g) * 2 - 1)).to(cuda_device).requires_grad_(True)

    def step(reuse):
        native.REUSE_GENERATOR_TURN = reuse
        try:
            d.requires_grad_(False)                          # generator turn: weights frozen (toggle_optimizer)
            _, sg, fr, fg = pair(d, wav, wav_hat)
            (GeneratorLoss()([sg])[0] + FeatureMatchingLoss()([fr], [fg])).backward()
            d.requires_grad_(True)
            d.zero_grad(set_to_none=True)
            sr, sg2, _, _ = pair(d, wav, wav_hat.detach())   # discriminator turn on the cached waveforms
            hit = native.LAST_PAIR_REUSED
            loss = DiscriminatorLoss()([sr], [sg2])[0] + 0.75 * torch.clamp(1 - sr, min=0).mean()
            (loss * 1024.0).backward()
            return float(loss), {n: p.grad.detach().clone() for n, p in d.named_parameters()}, hit
        finally:
            native.REUSE_GENERATOR_TURN = True

    l0, g0, hit0 = step(False)
    l1, g1, hit1 = step(True)
    assert not hit0 and hit1
    assert l0 == l1                                           # the very same forward values
    worst = max(_param_rel(g1[n], g0[n]) for n in g0)
    print(f"  {kind}: discriminator-turn loss {l1:.6f}; worst weight-gradient difference reuse vs recompute {worst:.3e}")
    assert worst <= 1e-3                                      # atomics in the weight-gradient kernels reorder sums
    # a different generated waveform (new storage) or a changed weight: no reuse
    d.requires_grad_(False)
    pair(d, wav, wav_hat)
    d.requires_grad_(True)
    pair(d, wav, wav_hat.detach().clone())
    assert not native.LAST_PAIR_REUSED
    d.requires_grad_(False)
    pair(d, wav, wav_hat)
    d.requires_grad_(True)
    with torch.no_grad():
        next(d.parameters()).mul_(1.0)                        # in-place write bumps the version counter
    pair(d, wav, wav_hat.detach())
    assert not native.LAST_PAIR_REUSED
