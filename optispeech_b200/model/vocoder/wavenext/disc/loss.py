"""Losses of the Vocos-style discriminator object (reference: disc/loss.py:11-270).

GAN hinge / feature-matching terms are thin reductions over discriminator outputs.  The two spectral
reconstruction losses (mel L1 and multi-resolution STFT) keep the reference's module tree — including the
buffers that appear in its state_dict (`mel_spec.spectrogram.window`, `mel_spec.mel_scale.fb`,
`stft_losses.{i}.window`) — and evaluate through `spectral.py`.
"""
from __future__ import annotations

import math
from typing import List

import torch
from torch import Tensor, nn

from . import spectral


class GeneratorLoss(nn.Module):
    """Hinge generator loss: sum over sub-discriminators of mean(relu(1 - D(G)))  (reference :11-31)."""

    def forward(self, disc_outputs: List[Tensor]):
        parts = [torch.clamp(1 - dg, min=0).mean() for dg in disc_outputs]
        return sum(parts), parts


class DiscriminatorLoss(nn.Module):
    """Hinge discriminator loss (reference :34-64).  Per-sub-discriminator terms are returned as device tensors
    (the reference calls .item() on each, forcing a host sync per term)."""

    def forward(self, disc_real_outputs: List[Tensor], disc_generated_outputs: List[Tensor]):
        r_losses = [torch.clamp(1 - dr, min=0).mean() for dr in disc_real_outputs]
        g_losses = [torch.clamp(1 + dg, min=0).mean() for dg in disc_generated_outputs]
        return sum(r_losses) + sum(g_losses), r_losses, g_losses


class FeatureMatchingLoss(nn.Module):
    """L1 between real / generated feature maps, summed over layers and sub-discriminators (reference :67-85)."""

    def forward(self, fmap_r, fmap_g) -> Tensor:
        from .native import feature_l1  # flat fp16 maps of the native period discriminators, or plain tensors

        return sum(feature_l1(rl, gl) for dr, dg in zip(fmap_r, fmap_g) for rl, gl in zip(dr, dg))


class _Spectrogram(nn.Module):
    def __init__(self, win_length: int):
        super().__init__()
        self.register_buffer("window", torch.hann_window(win_length))


class _MelScale(nn.Module):
    def __init__(self, n_stft: int, n_mels: int, sample_rate: int, f_min: float, f_max: float):
        super().__init__()
        self.register_buffer("fb", htk_mel_filterbank(n_stft, n_mels, sample_rate, f_min, f_max))


class _MelSpec(nn.Module):
    def __init__(self, sample_rate, n_fft, win_length, n_mels, f_min, f_max):
        super().__init__()
        self.spectrogram = _Spectrogram(win_length)
        self.mel_scale = _MelScale(n_fft // 2 + 1, n_mels, sample_rate, f_min, f_max)
        self._ranges = None

    def ranges(self):
        """Non-zero extents of the triangular filters (derived from `fb`, cached; recomputed if `fb` was reloaded)."""
        fb = self.mel_scale.fb
        key = (fb.data_ptr(), fb._version, str(fb.device))
        if self._ranges is None or self._ranges[0] != key:
            self._ranges = (key, spectral.filterbank_ranges(fb).to(fb.device))
        return self._ranges[1]


def htk_mel_filterbank(n_freqs: int, n_mels: int, sample_rate: int, f_min: float, f_max: float) -> Tensor:
    """(n_freqs, n_mels) triangular filters on the HTK mel scale, no area normalisation — what
    torchaudio.transforms.MelSpectrogram(mel_scale='htk', norm=None) builds (reference :94-107)."""
    freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    to_mel = lambda f: 2595.0 * math.log10(1.0 + f / 700.0)  # noqa: E731
    m_pts = torch.linspace(to_mel(f_min), to_mel(f_max), n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - freqs.unsqueeze(1)
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.min(down, up), min=0.0)


class MelSpecReconstructionLoss(nn.Module):
    """L1 distance of log-mel magnitudes (reference :88-120)."""

    def __init__(self, sample_rate, n_fft, hop_length, win_length, n_mels, f_min, f_max, clip_val=1e-7):
        super().__init__()
        self.clip_val = clip_val
        self.n_fft, self.hop_length, self.win_length = n_fft, hop_length, win_length
        self.mel_spec = _MelSpec(sample_rate, n_fft, win_length, n_mels, f_min, f_max)

    def forward(self, y_hat: Tensor, y: Tensor) -> Tensor:
        return spectral.mel_l1_loss(y_hat, y, self.mel_spec.spectrogram.window, self.mel_spec.mel_scale.fb, self.mel_spec.ranges(),
                                    self.n_fft, self.hop_length, self.win_length, self.clip_val)


class STFTLoss(nn.Module):
    """One resolution of the MR-STFT loss (reference :197-228): (spectral convergence, log-magnitude L1)."""

    def __init__(self, fft_size=1024, shift_size=120, win_length=600, window="hann_window"):
        super().__init__()
        self.fft_size, self.shift_size, self.win_length = fft_size, shift_size, win_length
        self.register_buffer("window", getattr(torch, window)(win_length))

    def forward(self, x: Tensor, y: Tensor):
        return spectral.stft_sc_mag_loss(x, y, self.window, self.fft_size, self.shift_size, self.win_length)


class MultiResolutionSTFTLoss(nn.Module):
    """Average of STFTLoss over three resolutions (reference :145-194)."""

    def __init__(self, fft_sizes=(1024, 2048, 512), hop_sizes=(120, 240, 50), win_lengths=(600, 1200, 240), window="hann_window"):
        super().__init__()
        assert len(fft_sizes) == len(hop_sizes) == len(win_lengths)
        self.stft_losses = nn.ModuleList([STFTLoss(fs, ss, wl, window) for fs, ss, wl in zip(fft_sizes, hop_sizes, win_lengths)])

    def forward(self, x: Tensor, y: Tensor):
        if x.dim() == 3:
            x, y = x.reshape(-1, x.size(2)), y.reshape(-1, y.size(2))
        sc, mag = 0.0, 0.0
        for f in self.stft_losses:
            s, m = f(x, y)
            sc, mag = sc + s, mag + m
        n = len(self.stft_losses)
        return sc / n, mag / n
