"""Period discriminators on the B200 kernels (reference: disc/_discriminators.py:41-97).

A `DiscriminatorP` views the waveform as `period` interleaved sequences and runs six (k,1) convolutions along them.  Here the
sequences of a layer are ONE flat fp16 matrix (a sequence = P consecutive rows: L valid rows, then zeros; see
include/osb200.h, "Multi-period discriminator"), so that

  layer 1 (C_in = 1)          osb_mpd_first_fwd      period split + reflect padding + FIR + LeakyReLU
  layers 2-4 (stride 3)       osb_gemm               tcgen05 implicit GEMM, TMA traversal stride 3, LeakyReLU + keep-mask epilogue
  layer 5 (stride 1)          osb_gemm
  conv_post (C_out = 1)       osb_mpd_post_fwd

and, backward: osb_lrelu_bwd_h16 (gate), osb_gemm_wgrad_strided (weight gradient), one GEMM against the tap-reversed weight
pack + osb_col2im_h16 (data gradient of the strided layers), the MN-major forward pack (data gradient of layer 5).

Weight normalisation stays in autograd (`torch._weight_norm`), so `weight_g` / `weight_v` receive the reference's gradients.
fp16 gradients inside the stack carry an extra factor GRAD_SCALE on top of the caller's loss scale (the 1/numel of the
feature-matching means would otherwise sit in fp16's subnormal range); it is removed wherever an fp32 gradient leaves.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F

from ..... import ops

GRAD_SCALE = 256.0
KSIZE, PAD = 5, 2
C1_PAD = 64          # layer 1 has 32 channels; its output is stored 64 wide (one 64-element k-block of the next GEMM)


@dataclass(frozen=True)
class Geometry:
    """Row bookkeeping of one period discriminator for signals of T samples."""
    period: int
    T: int
    L: Tuple[int, ...]   # valid rows per sequence: L[0] input, L[1..5] layer outputs
    P: Tuple[int, ...]   # rows a sequence owns in the flat matrix of layer i (P[0] unused)

    @staticmethod
    def make(T: int, period: int, stride: int = 3) -> "Geometry":
        L = [(T + period - 1) // period]
        for _ in range(4):
            L.append((L[-1] + 2 * PAD - KSIZE) // stride + 1)
        L.append(L[-1])
        p4 = L[4] + 2
        P = (0, p4 * stride ** 3, p4 * stride ** 2, p4 * stride, p4, p4)
        for i in range(1, 6):
            assert P[i] >= L[i] + 2 or i >= 4, (P, L)
        return Geometry(period, T, tuple(L), P)


class FlatMap:
    """A feature map in the flat layout: data (NSEQ*P, C) fp16.  `dense()` gives the reference's (N, C, L, period) tensor."""

    def __init__(self, data: torch.Tensor, period: int, L: int, P: int):
        self.data, self.period, self.L, self.P = data, period, L, P

    @property
    def n_valid(self) -> int:
        return (self.data.shape[0] // self.P) * self.L * self.data.shape[1]

    def half(self, which: int) -> "FlatMap":
        rows = self.data.shape[0] // 2
        return FlatMap(self.data[which * rows:(which + 1) * rows], self.period, self.L, self.P)

    def dense(self) -> torch.Tensor:
        """fp32 (N, C, L, period), differentiable: gradients of a loss on the dense map re-enter the stack scaled."""
        return _DenseFn.apply(self.data, self.period, self.L, self.P)


class _DenseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, period: int, L: int, P: int):
        rows, Cc = data.shape
        ctx.geo = (rows, Cc, period, L, P)
        return data.view(rows // (P * period), period, P, Cc)[:, :, :L].permute(0, 3, 2, 1).float()

    @staticmethod
    def backward(ctx, g):
        rows, Cc, period, L, P = ctx.geo
        flat = torch.zeros((rows, Cc), device=g.device, dtype=torch.float16)
        flat.view(rows // (P * period), period, P, Cc)[:, :, :L] = (g.permute(0, 3, 2, 1) * GRAD_SCALE).to(torch.float16)
        return flat, None, None, None


class _FirstFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, wav, w, bias, geom: Geometry, stride: int, slope: float):
        w2 = w.reshape(32, KSIZE).contiguous()
        y = ops.mpd_first_fwd(wav, w2, bias, geom.period, geom.L[1], geom.P[1], C1_PAD, stride, slope)
        ctx.save_for_backward(wav, w2, y)
        ctx.geom, ctx.stride, ctx.slope, ctx.w_shape = geom, stride, slope, w.shape
        return y

    @staticmethod
    def backward(ctx, gy):
        wav, w2, y = ctx.saved_tensors
        geom = ctx.geom
        g = ops.lrelu_bwd_h16(gy.contiguous(), y, geom.P[1], geom.L[1], ctx.slope)
        want_dwav, want_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dwav, dw, db = ops.mpd_first_bwd(g, wav, w2, geom.period, geom.L[1], geom.P[1], ctx.stride, 1.0 / GRAD_SCALE, want_dwav, want_dw)
        return dwav, (dw.reshape(ctx.w_shape) if dw is not None else None), db, None, None, None


class _ConvFn(torch.autograd.Function):
    """One (5,1) convolution + LeakyReLU on the flat layout: x (rows_in, Cin_p) fp16 -> y (rows_out, Cout) fp16."""

    @staticmethod
    def forward(ctx, x, w, bias, wp, stride: int, P_out: int, L_out: int, slope: float):
        rows_in, cin_p = x.shape
        rows_out = rows_in // stride
        _, y, _ = ops.gemm(x.view(1, rows_in, cin_p), wp, epi=ops.EPI_BIAS, flags=ops.FLAG_OUT_H16 | ops.FLAG_NO_F32 | ops.FLAG_KEEPMASK,
                           pad=PAD, bias=bias, seq_rows=(P_out, L_out), row_stride=stride, lrelu=slope)
        y = y.view(rows_out, -1)
        ctx.save_for_backward(x, w, wp, y)
        ctx.stride, ctx.P_out, ctx.L_out, ctx.slope = stride, P_out, L_out, slope
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, wp, y = ctx.saved_tensors
        stride = ctx.stride
        rows_in, cin_p = x.shape
        rows_out, cout = y.shape
        cin = w.shape[1]
        g = ops.lrelu_bwd_h16(gy.contiguous(), y, ctx.P_out, ctx.L_out, ctx.slope)
        dx = dw = db = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dwp = torch.zeros((KSIZE, cout, cin_p), device=x.device, dtype=torch.float32)
            ops.gemm_wgrad_strided(g, x, dwp, taps=KSIZE, pad=PAD, stride=stride)
            dw = (dwp[:, :, :cin].permute(1, 2, 0) * (1.0 / GRAD_SCALE)).contiguous().view(w.shape)
            db = ops.colsum_h16(g) * (1.0 / GRAD_SCALE)
        if ctx.needs_input_grad[0]:
            if stride == 1:   # the forward pack read as an MN-major operand, taps reversed
                _, dx, _ = ops.gemm(g.view(1, rows_out, cout), wp, epi=ops.EPI_BIAS, flags=ops.FLAG_OUT_H16 | ops.FLAG_NO_F32,
                                    pad=KSIZE - 1 - PAD, w_mn=True, tap_reverse=True, N=cin_p)
                dx = dx.view(rows_in, cin_p)
            else:
                # per-tap products in one GEMM (N = 5 * Cin_p), gathered by col2im: row r of g meets input row 3r + tap - 2
                w3 = w.reshape(cout, cin, KSIZE)
                if cin_p != cin:
                    w3 = F.pad(w3, (0, 0, 0, cin_p - cin))
                wt = ops.pack_conv_h16(w3, transpose_reverse=True).view(1, KSIZE * cin_p, cout)
                _, col, _ = ops.gemm(g.view(1, rows_out, cout), wt, epi=ops.EPI_BIAS, flags=ops.FLAG_OUT_H16 | ops.FLAG_NO_F32)
                dx = ops.col2im_h16(col.view(rows_out, KSIZE * cin_p), rows_in, cin_p, KSIZE, PAD, stride, reversed_taps=True)
        return dx, dw, db, None, None, None, None, None


class _PostFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, period: int, L: int, P: int):
        w2 = w.reshape(-1, 3).contiguous()
        score = ops.mpd_post_fwd(x, w2, bias, period, L, P)
        ctx.save_for_backward(x, w2)
        ctx.geo, ctx.w_shape = (period, L, P), w.shape
        return score

    @staticmethod
    def backward(ctx, dscore):
        x, w2 = ctx.saved_tensors
        period, L, P = ctx.geo
        want_dw = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dx, dw, db = ops.mpd_post_bwd(dscore, x, w2, period, L, P, GRAD_SCALE, ctx.needs_input_grad[0], want_dw)
        return dx, (dw.reshape(ctx.w_shape) if dw is not None else None), db, None, None, None


class L1PairFn(torch.autograd.Function):
    """mean |real - fake| over the valid elements of two flat feature maps; gradient with respect to `fake` only (the real
    map's graph is not needed: in the generator turn the discriminator weights are frozen)."""

    @staticmethod
    def forward(ctx, real, fake, n_valid: int):
        s = ops.l1_pair_fwd(real, fake)
        ctx.save_for_backward(real, fake)
        ctx.n_valid = n_valid
        return s[0] / n_valid

    @staticmethod
    def backward(ctx, gout):
        real, fake = ctx.saved_tensors
        return None, ops.l1_pair_bwd(real, fake, gout.reshape(1).float().contiguous(), GRAD_SCALE / ctx.n_valid), None


def feature_l1(real, fake) -> torch.Tensor:
    """One term of the feature-matching loss for a pair of maps (FlatMap or plain tensors)."""
    if isinstance(real, FlatMap):
        return L1PairFn.apply(real.data.detach(), fake.data, real.n_valid)
    return (real - fake).abs().mean()


def effective_weights(disc):
    """Weight-normalised fp32 weights (autograd tracks them back to weight_g / weight_v) and their fp16 forward packs."""
    layers = []
    for i, conv in enumerate(list(disc.convs) + [disc.conv_post]):
        w = torch._weight_norm(conv.weight_v, conv.weight_g, 0)
        wp = None
        if 1 <= i <= 4:
            cout, cin = w.shape[0], w.shape[1]
            wp = ops.pack_conv_h16(w.detach().reshape(cout, cin, KSIZE), k_pad=max(cin, C1_PAD))
        layers.append((w, conv.bias, wp))
    return layers


def period_forward(disc, wav: torch.Tensor, layers=None):
    """Native forward of one DiscriminatorP on (NS, T) fp32 signals: (score (NS, L5*period) fp32, [FlatMap x 4, score map])."""
    if not wav.is_cuda:
        raise RuntimeError("the period discriminators run on the CUDA kernels only (no CPU path)")
    wav = wav.float().contiguous()
    NS, T = wav.shape
    geom = Geometry.make(T, disc.period)
    layers = layers if layers is not None else effective_weights(disc)
    slope = float(disc.lrelu_slope)
    (w1, b1, _), rest = layers[0], layers[1:5]
    x = _FirstFn.apply(wav, w1, b1, geom, 3, slope)
    fmap: List[object] = []
    for i, (w, b, wp) in enumerate(rest, start=2):
        stride = 3 if i <= 4 else 1
        x = _ConvFn.apply(x, w, b, wp, stride, geom.P[i], geom.L[i], slope)
        fmap.append(FlatMap(x, disc.period, geom.L[i], geom.P[i]))
    wpost, bpost, _ = layers[5]
    score = _PostFn.apply(x, wpost, bpost, disc.period, geom.L[5], geom.P[5])
    fmap.append(score.view(NS, 1, geom.L[5], disc.period))
    return score, fmap


def period_forward_pair(disc, y: torch.Tensor, y_hat: torch.Tensor):
    """(score_real, score_fake, fmap_real, fmap_fake) of one period discriminator.  Generator turn (y_hat carries a graph):
    the real signals run without a graph and the generated ones alone are differentiated; otherwise both halves share one
    pass (one GEMM per layer over 2B signals)."""
    layers = effective_weights(disc)
    if torch.is_grad_enabled() and y_hat.requires_grad:
        with torch.no_grad():
            sr, fr = period_forward(disc, y, [(w.detach(), b.detach(), wp) for w, b, wp in layers])
        sg, fg = period_forward(disc, y_hat, layers)
        return sr, sg, fr, fg
    B = y.shape[0]
    s, f = period_forward(disc, torch.cat((y.float(), y_hat.float()), dim=0), layers)
    fr = [m.half(0) if isinstance(m, FlatMap) else m[:B] for m in f]
    fg = [m.half(1) if isinstance(m, FlatMap) else m[B:] for m in f]
    return s[:B], s[B:], fr, fg
