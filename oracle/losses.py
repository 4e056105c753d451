"""Spectral reconstruction losses of the discriminator object, restated in plain fp32 PyTorch.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows optispeech/model/vocoder/wavenext/disc/loss.py
(MelSpecReconstructionLoss :88-120, stft :123-142, STFTLoss/SpectralConvergence/LogSTFTMagnitude/
MultiResolutionSTFTLoss :145-270) and disc/__init__.py:98-111 (loss weights, argument order).
torchaudio's MelSpectrogram is restated from its published definition (HTK mel scale, norm=None,
power=1, centre/reflect STFT with a periodic Hann window) so that the oracle does not need torchaudio.
"""
from __future__ import annotations

import math

import torch

MR_STFT_RESOLUTIONS = ((1024, 120, 600), (2048, 240, 1200), (512, 50, 240))  # (n_fft, hop, win)  loss.py:151-153


def stft_magnitude(x: torch.Tensor, n_fft: int, hop: int, win: int, clamp: float | None = 1e-7) -> torch.Tensor:
    """(B, T) -> (B, frames, n_fft/2+1).  Hann(win) periodic, zero-padded (centred) to n_fft by torch.stft."""
    window = torch.hann_window(win, dtype=x.dtype, device=x.device)
    spec = torch.stft(x, n_fft, hop, win, window, center=True, pad_mode="reflect", return_complex=True)
    power = spec.real ** 2 + spec.imag ** 2
    if clamp is not None:
        power = torch.clamp(power, min=clamp)
    return torch.sqrt(power).transpose(2, 1)


def mr_stft_loss(x_hat: torch.Tensor, y: torch.Tensor):
    """-> (spectral convergence, log-magnitude L1), each averaged over the three resolutions.  x_hat = prediction,
    y = ground truth (the denominator of SC is the ground-truth norm, disc/__init__.py:110)."""
    sc_total, mag_total = 0.0, 0.0
    for n_fft, hop, win in MR_STFT_RESOLUTIONS:
        xm = stft_magnitude(x_hat, n_fft, hop, win)
        ym = stft_magnitude(y, n_fft, hop, win)
        sc_total = sc_total + torch.linalg.norm((ym - xm).reshape(-1)) / torch.linalg.norm(ym.reshape(-1))
        mag_total = mag_total + (torch.log(ym) - torch.log(xm)).abs().mean()
    n = len(MR_STFT_RESOLUTIONS)
    return sc_total / n, mag_total / n


def mel_filterbank(sample_rate: int, n_fft: int, n_mels: int, f_min: float, f_max: float) -> torch.Tensor:
    """(n_fft/2+1, n_mels) triangular HTK filters without area normalisation (torchaudio melscale_fbanks)."""
    n_freqs = n_fft // 2 + 1
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def log_mel(x: torch.Tensor, fb: torch.Tensor, n_fft: int, hop: int, win: int, clip_val: float = 1e-7) -> torch.Tensor:
    """(B, T) -> (B, n_mels, frames) log-mel magnitudes (power=1, no clamp before the filterbank)."""
    mag = stft_magnitude(x, n_fft, hop, win, clamp=None)  # (B, frames, bins)
    mel = mag @ fb
    return torch.log(torch.clip(mel, min=clip_val)).transpose(1, 2)


def mel_loss(x_hat: torch.Tensor, y: torch.Tensor, fb: torch.Tensor, n_fft: int, hop: int, win: int) -> torch.Tensor:
    return (log_mel(y, fb, n_fft, hop, win) - log_mel(x_hat, fb, n_fft, hop, win)).abs().mean()


def forward_val_losses(wav: torch.Tensor, wav_hat: torch.Tensor, spec, fb: torch.Tensor | None = None):
    """VocosDiscriminator.forward_val (disc/__init__.py:98-103): (45 * mel L1, 2.5 * (SC + MAG))."""
    if fb is None:
        fb = mel_filterbank(spec.sample_rate, spec.n_fft, spec.n_feats, spec.f_min, spec.f_max)
    ml = mel_loss(wav_hat, wav, fb, spec.n_fft, spec.hop_length, spec.win_length) * spec.lambda_mel
    sc, mag = mr_stft_loss(wav_hat, wav)
    return ml, (sc + mag) * spec.lambda_mr_stft, sc, mag
