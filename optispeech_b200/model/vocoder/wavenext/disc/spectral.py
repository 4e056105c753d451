"""Spectral reconstruction losses: framing (centre / reflect), Hann window, rFFT, magnitude, reductions.

Semantics follow reference disc/loss.py:123-142 (`stft`: sqrt(clamp(re^2+im^2, 1e-7))), :231-270 (spectral
convergence = ||Y - X||_F / ||Y||_F over the whole batch tensor; log-magnitude L1) and :109-120 (log-mel L1,
magnitudes NOT clamped before the filterbank, log(clip(., 1e-7)) after).
"""
from __future__ import annotations

import torch


def _stft_power(x: torch.Tensor, n_fft: int, hop: int, win: int, window: torch.Tensor) -> torch.Tensor:
    spec = torch.stft(x, n_fft, hop, win, window, center=True, pad_mode="reflect", return_complex=True)
    return spec.real ** 2 + spec.imag ** 2  # (B, bins, frames)


def stft_sc_mag_loss(x_hat, y, window, n_fft: int, hop: int, win: int):
    xm = torch.sqrt(torch.clamp(_stft_power(x_hat, n_fft, hop, win, window), min=1e-7))
    ym = torch.sqrt(torch.clamp(_stft_power(y, n_fft, hop, win, window), min=1e-7))
    sc = torch.linalg.norm((ym - xm).reshape(-1)) / torch.linalg.norm(ym.reshape(-1))
    mag = (torch.log(ym) - torch.log(xm)).abs().mean()
    return sc, mag


def mel_l1_loss(x_hat, y, window, fb, n_fft: int, hop: int, win: int, clip_val: float = 1e-7):
    def log_mel(sig):
        mag = torch.sqrt(_stft_power(sig, n_fft, hop, win, window))       # (B, bins, frames), power=1, no clamp
        mel = torch.matmul(mag.transpose(1, 2), fb)                        # (B, frames, n_mels)
        return torch.log(torch.clip(mel, min=clip_val))

    return (log_mel(y) - log_mel(x_hat)).abs().mean()
