"""Multi-period and multi-resolution discriminators (reference: disc/_discriminators.py:10-216).

The module tree is the reference's, so that `state_dict` keys (weight-norm `weight_g` / `weight_v` included) are
interchangeable.  On CUDA tensors the period discriminators run on this package's kernels (disc/native.py: tcgen05 implicit
GEMMs over a flat fp16 sequence layout; feature maps come back as `native.FlatMap`, `.dense()` gives the reference's
(B, C, L, period) tensor), and so do the resolution discriminators (rectangular-window |STFT| by torch.stft / cuFFT, then the
Conv2d stack as gathers + tcgen05 implicit GEMMs).  CPU tensors take the stock PyTorch path (module-tree / state_dict tests).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.utils import weight_norm


NATIVE_MPD = True   # False: the period discriminators run on cuDNN as well (A/B comparisons in tests and bench.py)
NATIVE_MRD = True   # likewise for the resolution discriminators


def _wn_conv(cin, cout, kernel, stride, padding):
    return weight_norm(nn.Conv2d(cin, cout, kernel, stride, padding=padding))


class DiscriminatorP(nn.Module):
    """Period discriminator: (B,T) -> reflect tail-pad to a multiple of `period` -> (B,1,T/p,p) -> (5,1) strided convs."""

    def __init__(self, period: int, kernel_size: int = 5, stride: int = 3, lrelu_slope: float = 0.1):
        super().__init__()
        self.period = period
        pad = (kernel_size // 2, 0)
        chans = [1, 32, 128, 512, 1024]
        layers = [_wn_conv(chans[i], chans[i + 1], (kernel_size, 1), (stride, 1), pad) for i in range(4)]
        layers.append(_wn_conv(1024, 1024, (kernel_size, 1), (1, 1), pad))
        self.convs = nn.ModuleList(layers)
        self.conv_post = _wn_conv(1024, 1, (3, 1), 1, (1, 0))
        self.lrelu_slope = lrelu_slope

    def forward(self, x: torch.Tensor):
        if x.is_cuda and NATIVE_MPD:
            from . import native
            return native.period_forward(self, x)
        x = x.unsqueeze(1)
        b, c, t = x.shape
        rem = t % self.period
        if rem != 0:
            x = F.pad(x, (0, self.period - rem), "reflect")
            t = x.shape[-1]
        x = x.view(b, c, t // self.period, self.period)
        fmap = []
        for i, conv in enumerate(self.convs):
            x = F.leaky_relu(conv(x), self.lrelu_slope)
            if i > 0:  # the first layer's map is not part of the feature-matching set (reference :81-85)
                fmap.append(x)
        x = self.conv_post(x)
        fmap.append(x)
        return torch.flatten(x, 1, -1), fmap


class DiscriminatorR(nn.Module):
    """Resolution discriminator on the rectangular-window magnitude STFT (reference :139-216)."""

    def __init__(self, resolution: Tuple[int, int, int], channels: int = 64, lrelu_slope: float = 0.1):
        super().__init__()
        self.resolution = resolution
        self.lrelu_slope = lrelu_slope
        self.convs = nn.ModuleList(
            [
                _wn_conv(1, channels, (7, 5), (2, 2), (3, 2)),
                _wn_conv(channels, channels, (5, 3), (2, 1), (2, 1)),
                _wn_conv(channels, channels, (5, 3), (2, 2), (2, 1)),
                _wn_conv(channels, channels, 3, (2, 1), 1),
                _wn_conv(channels, channels, 3, (2, 2), 1),
            ]
        )
        self.conv_post = _wn_conv(channels, 1, (3, 3), 1, (1, 1))

    def spectrogram(self, x: torch.Tensor) -> torch.Tensor:
        n_fft, hop, win = self.resolution
        return torch.stft(x, n_fft=n_fft, hop_length=hop, win_length=win, window=torch.ones(n_fft, device=x.device), center=True,
                          return_complex=True).abs()

    def forward(self, x: torch.Tensor):
        if x.is_cuda and NATIVE_MRD:
            from . import native
            return native.resolution_forward(self, x)
        x = self.spectrogram(x).unsqueeze(1)
        fmap = []
        for conv in self.convs:
            x = F.leaky_relu(conv(x), self.lrelu_slope)
            fmap.append(x)
        x = self.conv_post(x)
        fmap.append(x)
        return torch.flatten(x, 1, -1), fmap


class _Multi(nn.Module):
    def forward(self, y: torch.Tensor, y_hat: torch.Tensor):
        real, fake, fr, ff = [], [], [], []
        for d in self.discriminators:
            a, fa = d(y)
            b, fb = d(y_hat)
            real.append(a); fr.append(fa); fake.append(b); ff.append(fb)
        return real, fake, fr, ff


class MultiPeriodDiscriminator(_Multi):
    def __init__(self, periods: Tuple[int, ...] = (2, 3, 5, 7, 11)):
        super().__init__()
        self.discriminators = nn.ModuleList([DiscriminatorP(period=p) for p in periods])

    def forward(self, y: torch.Tensor, y_hat: torch.Tensor):
        if not (y.is_cuda and NATIVE_MPD):
            return super().forward(y, y_hat)
        from . import native
        real, fake, fr, ff = [], [], [], []
        for a, b, fa, fb in native.fan_out(self.discriminators, y, y_hat, native.period_forward_pair, native.DISC_STREAM_SLOT0):
            real.append(a); fr.append(fa); fake.append(b); ff.append(fb)
        return real, fake, fr, ff


class MultiResolutionDiscriminator(_Multi):
    def __init__(self, resolutions=((1024, 256, 1024), (2048, 512, 2048), (512, 128, 512))):
        super().__init__()
        self.discriminators = nn.ModuleList([DiscriminatorR(resolution=r) for r in resolutions])

    def forward(self, y: torch.Tensor, y_hat: torch.Tensor):
        if not (y.is_cuda and NATIVE_MRD):
            return super().forward(y, y_hat)
        from . import native
        real, fake, fr, ff = [], [], [], []
        for a, b, fa, fb in native.fan_out(self.discriminators, y, y_hat, native.resolution_forward_pair, native.DISC_STREAM_SLOT0 + 8):
            real.append(a); fr.append(fa); fake.append(b); ff.append(fb)
        return real, fake, fr, ff
