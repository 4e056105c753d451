"""Functional fp32 restatement of the GAN-phase discriminators and their losses.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows optispeech/model/vocoder/wavenext/disc/_discriminators.py
(MultiPeriodDiscriminator :10-38, DiscriminatorP :41-97, MultiResolutionDiscriminator :100-135, DiscriminatorR :138-216),
disc/loss.py (GeneratorLoss :11-31, DiscriminatorLoss :34-64, FeatureMatchingLoss :67-85) and the loss assembly of
disc/__init__.py:44-96 (`forward_disc`, `forward_gen`).  Takes a flat state dict with the reference's keys
(`multiperioddisc.discriminators.{i}.convs.{j}.weight_g|weight_v|bias`, `...conv_post.*`, `multiresddisc....`):
`torch.nn.utils.weight_norm` stores w = g * v / ||v|| with the norm over every dimension but the first.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from . import losses as L

SD = Dict[str, torch.Tensor]
PERIODS = (2, 3, 5, 7, 11)
RESOLUTIONS = ((1024, 256, 1024), (2048, 512, 2048), (512, 128, 512))
LRELU = 0.1


def wn_weight(sd: SD, prefix: str) -> torch.Tensor:
    """weight_norm(dim=0): w[o] = g[o] * v[o] / ||v[o]||_2."""
    g, v = sd[f"{prefix}.weight_g"], sd[f"{prefix}.weight_v"]
    return g * v / v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))


def discriminator_p(sd: SD, prefix: str, x: torch.Tensor, period: int) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """(B, T) -> reflect tail-pad to a multiple of `period` -> (B,1,T/p,p) -> four (5,1)/(3,1) convs, one (5,1)/(1,1), post (3,1).
    Feature maps: outputs of convs 1..4 (after LeakyReLU) and of conv_post (the first conv's output is NOT collected)."""
    x = x.unsqueeze(1)
    b, c, t = x.shape
    if t % period != 0:
        n_pad = period - (t % period)
        x = F.pad(x, (0, n_pad), "reflect")
        t = t + n_pad
    x = x.view(b, c, t // period, period)
    fmap = []
    for i in range(5):
        stride = (3, 1) if i < 4 else (1, 1)
        x = F.conv2d(x, wn_weight(sd, f"{prefix}.convs.{i}"), sd[f"{prefix}.convs.{i}.bias"], stride=stride, padding=(2, 0))
        x = F.leaky_relu(x, LRELU)
        if i > 0:
            fmap.append(x)
    x = F.conv2d(x, wn_weight(sd, f"{prefix}.conv_post"), sd[f"{prefix}.conv_post.bias"], stride=1, padding=(1, 0))
    fmap.append(x)
    return torch.flatten(x, 1, -1), fmap


def discriminator_r(sd: SD, prefix: str, x: torch.Tensor, resolution) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """(B, T) -> |STFT| with a RECTANGULAR window (no clamp) -> (B,1,freq,frames) -> five Conv2d(64) + LeakyReLU -> post."""
    n_fft, hop, win = resolution
    spec = torch.stft(x, n_fft=n_fft, hop_length=hop, win_length=win, window=torch.ones(n_fft, dtype=x.dtype), center=True,
                      return_complex=True).abs()
    x = spec.unsqueeze(1)
    cfg = (((7, 5), (2, 2), (3, 2)), ((5, 3), (2, 1), (2, 1)), ((5, 3), (2, 2), (2, 1)), ((3, 3), (2, 1), (1, 1)), ((3, 3), (2, 2), (1, 1)))
    fmap = []
    for i, (_, stride, pad) in enumerate(cfg):
        x = F.conv2d(x, wn_weight(sd, f"{prefix}.convs.{i}"), sd[f"{prefix}.convs.{i}.bias"], stride=stride, padding=pad)
        x = F.leaky_relu(x, LRELU)
        fmap.append(x)
    x = F.conv2d(x, wn_weight(sd, f"{prefix}.conv_post"), sd[f"{prefix}.conv_post.bias"], padding=(1, 1))
    fmap.append(x)
    return torch.flatten(x, 1, -1), fmap


def _multi(sd: SD, y, y_hat, which: str):
    outs_r, outs_g, fr, fg = [], [], [], []
    if which == "mpd":
        items = [(f"multiperioddisc.discriminators.{i}", lambda s, p, x, per=per: discriminator_p(s, p, x, per)) for i, per in enumerate(PERIODS)]
    else:
        items = [(f"multiresddisc.discriminators.{i}", lambda s, p, x, res=res: discriminator_r(s, p, x, res)) for i, res in enumerate(RESOLUTIONS)]
    for prefix, fn in items:
        r, fmr = fn(sd, prefix, y)
        g, fmg = fn(sd, prefix, y_hat)
        outs_r.append(r); outs_g.append(g); fr.append(fmr); fg.append(fmg)
    return outs_r, outs_g, fr, fg


def generator_loss(outs):
    parts = [torch.mean(torch.clamp(1 - dg, min=0)) for dg in outs]
    return sum(parts), parts


def discriminator_loss(outs_r, outs_g):
    parts = [torch.mean(torch.clamp(1 - dr, min=0)) + torch.mean(torch.clamp(1 + dg, min=0)) for dr, dg in zip(outs_r, outs_g)]
    return sum(parts), parts


def feature_matching_loss(fr, fg):
    return sum(torch.mean(torch.abs(rl - gl)) for dr, dg in zip(fr, fg) for rl, gl in zip(dr, dg))


def forward_disc(sd: SD, wav, wav_hat, lambda_mrd: float = 1.0):
    """disc/__init__.py:44-62: hinge losses of both discriminator families, each averaged over its sub-discriminators."""
    r_mp, g_mp, _, _ = _multi(sd, wav, wav_hat, "mpd")
    r_mr, g_mr, _, _ = _multi(sd, wav, wav_hat, "mrd")
    l_mp, p_mp = discriminator_loss(r_mp, g_mp)
    l_mr, p_mr = discriminator_loss(r_mr, g_mr)
    l_mp, l_mr = l_mp / len(p_mp), l_mr / len(p_mr)
    return l_mp + l_mr * lambda_mrd, dict(loss_mp=l_mp, loss_mrd=l_mr)


def forward_gen(sd: SD, wav, wav_hat, spec, fb=None):
    """disc/__init__.py:64-96: adversarial + feature-matching terms of both families plus 45 * mel and 2.5 * MR-STFT."""
    _, g_mp, fr_mp, fg_mp = _multi(sd, wav, wav_hat, "mpd")
    _, g_mr, fr_mr, fg_mr = _multi(sd, wav, wav_hat, "mrd")
    l_g_mp, p_mp = generator_loss(g_mp)
    l_g_mr, p_mr = generator_loss(g_mr)
    l_g_mp, l_g_mr = l_g_mp / len(p_mp), l_g_mr / len(p_mr)
    l_fm_mp = feature_matching_loss(fr_mp, fg_mp) / len(fr_mp)
    l_fm_mr = feature_matching_loss(fr_mr, fg_mr) / len(fr_mr)
    mel, mrstft, _, _ = L.forward_val_losses(wav, wav_hat, spec, fb)
    lam = spec.lambda_mrd
    loss = l_g_mp + l_g_mr * lam + l_fm_mp + l_fm_mr * lam + mel + mrstft
    return loss, dict(loss_gen_mp=l_g_mp, loss_gen_mrd=l_g_mr, loss_fm_mp=l_fm_mp, loss_fm_mrd=l_fm_mr, mel_loss=mel, mr_stft_loss=mrstft)


def discriminator_shapes() -> Dict[str, Tuple[int, ...]]:
    """state_dict keys -> shapes of the trainable part of VocosDiscriminator (the loss modules' buffers are not listed)."""
    s: Dict[str, Tuple[int, ...]] = {}

    def conv(prefix, cout, cin, kh, kw):
        s[f"{prefix}.bias"] = (cout,)
        s[f"{prefix}.weight_g"] = (cout, 1, 1, 1)
        s[f"{prefix}.weight_v"] = (cout, cin, kh, kw)

    for i in range(len(PERIODS)):
        p = f"multiperioddisc.discriminators.{i}"
        chans = [1, 32, 128, 512, 1024, 1024]
        for j in range(5):
            conv(f"{p}.convs.{j}", chans[j + 1], chans[j], 5, 1)
        conv(f"{p}.conv_post", 1, 1024, 3, 1)
    for i in range(len(RESOLUTIONS)):
        p = f"multiresddisc.discriminators.{i}"
        ks = [(7, 5), (5, 3), (5, 3), (3, 3), (3, 3)]
        for j, (kh, kw) in enumerate(ks):
            conv(f"{p}.convs.{j}", 64, 1 if j == 0 else 64, kh, kw)
        conv(f"{p}.conv_post", 1, 64, 3, 3)
    return s
