"""Fused STFT-domain loss kernels (forward partial sums + hand-written gradient) against the oracle (torch.stft based
restatement of the reference, itself pinned to the reference's MelSpecReconstructionLoss / MultiResolutionSTFTLoss by
tests/golden)."""
import numpy as np
import pytest
import torch

from oracle import losses as OL
from oracle.spec import ModelSpec

pytestmark = pytest.mark.gpu


def _disc(dev):
    from types import SimpleNamespace as NS

    from optispeech_b200.model.vocoder.wavenext.disc.loss import MelSpecReconstructionLoss, MultiResolutionSTFTLoss

    spec = ModelSpec()
    mel = MelSpecReconstructionLoss(spec.sample_rate, spec.n_fft, spec.hop_length, spec.win_length, spec.n_feats, spec.f_min, spec.f_max)
    return spec, mel.to(dev), MultiResolutionSTFTLoss().to(dev)


@pytest.mark.parametrize("B,L", [(2, 16384), (3, 5000)])
def test_mr_stft_and_mel_losses_match_oracle(cuda_device, B, L):
    spec, mel_mod, stft_mod = _disc(cuda_device)
    g = torch.Generator().manual_seed(21)
    y = (torch.rand(B, L, generator=g) * 2 - 1) * 0.8
    x = (y + 0.3 * torch.randn(B, L, generator=g)).clamp(-1, 1)
    x[0, :100] = 0.0   # a silent stretch exercises the magnitude clamp
    xr = x.clone().requires_grad_(True)
    fb = OL.mel_filterbank(spec.sample_rate, spec.n_fft, spec.n_feats, spec.f_min, spec.f_max)
    ref_mel = OL.mel_loss(xr, y, fb, spec.n_fft, spec.hop_length, spec.win_length)
    ref_sc, ref_mag = OL.mr_stft_loss(xr, y)
    ref_total = 45.0 * ref_mel + 2.5 * (ref_sc + ref_mag)
    (rg,) = torch.autograd.grad(ref_total, xr)

    xc = x.to(cuda_device).requires_grad_(True)
    yc = y.to(cuda_device)
    mel = mel_mod(xc, yc)
    sc, mag = stft_mod(xc, yc)
    total = 45.0 * mel + 2.5 * (sc + mag)
    (gg,) = torch.autograd.grad(total, xc)
    print(f"  mel {float(mel):.6f}/{float(ref_mel):.6f} sc {float(sc):.6f}/{float(ref_sc):.6f} mag {float(mag):.6f}/{float(ref_mag):.6f}")
    assert abs(float(mel) - float(ref_mel)) <= 2e-4 * abs(float(ref_mel))
    assert abs(float(sc) - float(ref_sc)) <= 2e-4 * abs(float(ref_sc))
    assert abs(float(mag) - float(ref_mag)) <= 2e-4 * abs(float(ref_mag))
    rel = float((gg.cpu() - rg).norm() / rg.norm())
    print(f"  d(loss)/d(x_hat) rel err {rel:.3e}")
    assert rel <= 2e-3


def test_forward_val_through_discriminator_object(cuda_device):
    from types import SimpleNamespace as NS

    from optispeech_b200.model.vocoder.wavenext.disc import VocosDiscriminator

    spec = ModelSpec()
    fe = NS(sample_rate=spec.sample_rate, n_fft=spec.n_fft, hop_length=spec.hop_length, win_length=spec.win_length, n_feats=spec.n_feats,
            f_min=spec.f_min, f_max=spec.f_max)
    disc = VocosDiscriminator(fe, NS(lambda_mrd=1.0, lambda_mel=45.0, lambda_mr_stft=2.5)).to(cuda_device)
    g = torch.Generator().manual_seed(22)
    y = torch.rand(2, 16384, generator=g) * 2 - 1
    x = torch.rand(2, 16384, generator=g) * 2 - 1
    loss, log = disc.forward_val(y.to(cuda_device), x.to(cuda_device))
    ml, stl, _, _ = OL.forward_val_losses(y, x, spec)
    assert abs(float(loss) - float(ml + stl)) <= 2e-4 * float(ml + stl)
    assert set(log) == {"mel_loss", "mr_stft_loss"}
