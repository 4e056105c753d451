"""GPU feature extraction (osb_mel_energy) against the CPU restatement of the reference's per-utterance pipeline
(feature_extractors/__init__.py:114-200): a ragged batch in one launch equals the per-utterance loop."""
import numpy as np
import pytest
import torch

from oracle import data as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sr,n_fft,hop,win,n_mels,fmax", [(22050, 1024, 256, 1024, 100, 11025), (22050, 1024, 256, 1024, 80, 8000),
                                                           (24000, 2048, 300, 1200, 100, 12000), (16000, 512, 128, 512, 40, 8000)])
def test_mel_and_energy_of_a_ragged_batch(cuda_device, sr, n_fft, hop, win, n_mels, fmax):
    from optispeech_b200.dataset.feature_extractors import CommonFeatureExtractor, slaney_mel_basis

    g = torch.Generator().manual_seed(n_fft + n_mels)
    lengths = torch.tensor([hop * 37, hop * 12 + 5, hop * 3, hop * 21 + hop - 1])
    Lmax = int(lengths.max())
    t = torch.arange(Lmax) / sr
    wav = torch.zeros(4, Lmax)
    for b, L in enumerate(lengths.tolist()):
        sig = 0.5 * torch.sin(2 * np.pi * (110.0 * (b + 1)) * t[:L]) + 0.1 * torch.randn(L, generator=g)
        wav[b, :L] = sig
    fe = CommonFeatureExtractor(sample_rate=sr, n_feats=n_mels, n_fft=n_fft, hop_length=hop, win_length=win, f_min=0, f_max=fmax, center=False)
    mel, energy, frames = fe.mel_and_energy(wav.to(cuda_device), lengths.to(cuda_device))
    basis = torch.from_numpy(slaney_mel_basis(sr, n_fft, n_mels, 0, fmax))
    worst_m = worst_e = 0.0
    for b, L in enumerate(lengths.tolist()):
        rm, re = O.mel_and_energy(wav[b, :L], basis, n_fft, hop, win)
        f = int(frames[b])
        assert f == rm.shape[-1] == L // hop
        worst_m = max(worst_m, float((mel[b, :, :f].cpu() - rm).abs().max()))
        worst_e = max(worst_e, float(((energy[b, :f].cpu() - re).abs() / re.abs().clamp(min=1e-3)).max()))
        assert float(mel[b, :, f:].abs().max() if f < mel.shape[-1] else 0.0) == 0.0     # frames past the utterance are zero
        assert float(energy[b, f:].abs().max() if f < energy.shape[-1] else 0.0) == 0.0
    print(f"  n_fft {n_fft} hop {hop} win {win}: log-mel max-abs diff {worst_m:.3e}, energy max relative diff {worst_e:.3e}")
    assert worst_m <= 2e-3 and worst_e <= 1e-4
    # single-utterance spelling of the reference
    m1 = fe.get_mel(wav[1, : int(lengths[1])].numpy())
    assert torch.allclose(m1.cpu(), mel[1, :, : int(frames[1])].cpu(), atol=1e-6)
    e1 = fe.get_energy(wav[1, : int(lengths[1])], mel_length=int(frames[1]) + 2)
    assert e1.shape[0] == int(frames[1]) + 2 and float(e1[-2:].abs().max()) == 0.0
