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

import itertools
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F

from ..... import ops

GRAD_SCALE = 256.0
KSIZE, PAD = 5, 2
C1_PAD = 64          # layer 1 has 32 channels; its output is stored 64 wide (one 64-element k-block of the next GEMM)

# fp16 operand packs of the discriminator weights, shared by the generator turn and the discriminator turn of one step (the
# weights do not change in between).  Keyed by the layer and its parameters' versions (torch in-place updates bump `_version`,
# FlatAdamW bumps `_osb_epoch`); emptied at the start of every generator turn (VocosDiscriminator.forward_gen), so an entry never
# outlives a step — inside a captured step the pack kernels of the generator turn are part of the graph and replayed with it.
_PACK_MEMO = {}
MEMO_IS_FRESH = False   # set by VocosDiscriminator.prefetch_real: the memo was emptied for this step already


def reset_pack_memo() -> None:
    _PACK_MEMO.clear()


def _memo(key, make):
    if key is None:
        return make()
    hit = _PACK_MEMO.get(key)
    if hit is None:
        if len(_PACK_MEMO) > 512:
            _PACK_MEMO.clear()
        hit = make()
        _PACK_MEMO[key] = hit
    return hit


_UIDS = itertools.count(1)


def _layer_key(conv):
    uid = conv.__dict__.get("_osb_uid")
    if uid is None:          # id() of a collected module can come back for a new one; this token cannot
        uid = conv.__dict__["_osb_uid"] = next(_UIDS)
    return (uid, conv.weight_v._version, conv.weight_g._version, getattr(conv.weight_v, "_osb_epoch", 0),
            getattr(conv.weight_g, "_osb_epoch", 0))


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
    def forward(ctx, wav, w, bias, geom: Geometry, stride: int, slope: float, y_pre=None):
        w2 = w.reshape(32, KSIZE).contiguous()
        y = y_pre if y_pre is not None else ops.mpd_first_fwd(wav, w2, bias, geom.period, geom.L[1], geom.P[1], C1_PAD, stride, slope)
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
        return dwav, (dw.reshape(ctx.w_shape) if dw is not None else None), db, None, None, None, None


def _conv_dgrad_pack(w, cin_p: int, stride: int, pkey):
    """Phase-decomposed data-gradient pack of a strided (5,1) layer (memoised per step).  It is first requested in the FORWARD
    pass of a turn that will need it: the pack kernels then sit before the point where the discriminator turn forks off
    (BaseModule: the two turns overlap), so both turns read a finished pack."""
    def make():
        cout, cin = w.shape[0], w.shape[1]
        w4 = w.detach().reshape(cout, cin, KSIZE, 1)
        if cin_p != cin:
            w4 = F.pad(w4, (0, 0, 0, 0, 0, cin_p - cin))
        return _phase_dgrad_pack(w4, PAD, stride)

    return _memo(("dgrad",) + pkey if pkey else None, make)


class _ConvFn(torch.autograd.Function):
    """One (5,1) convolution + LeakyReLU on the flat layout: x (rows_in, Cin_p) fp16 -> y (rows_out, Cout) fp16."""

    @staticmethod
    def forward(ctx, x, w, bias, wp, stride: int, P_out: int, L_out: int, slope: float, y_pre=None, pkey=None):
        ctx.pkey = pkey
        rows_in, cin_p = x.shape
        rows_out = rows_in // stride
        if stride != 1 and pkey is not None and ctx.needs_input_grad[0]:
            _conv_dgrad_pack(w, cin_p, stride, pkey)
        if y_pre is not None:
            y = y_pre
        else:
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
        want_w = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dx = dw = db = None
        if want_w:    # the bias gradient (column sums of g) comes out of the gate kernel
            g, db = ops.lrelu_bwd_h16(gy.contiguous(), y, ctx.P_out, ctx.L_out, ctx.slope, colsum_scale=1.0 / GRAD_SCALE)
            dwp = torch.zeros((KSIZE, cout, cin_p), device=x.device, dtype=torch.float32)
            ops.gemm_wgrad_strided(g, x, dwp, taps=KSIZE, pad=PAD, stride=stride)
            dw = (dwp[:, :, :cin].permute(1, 2, 0) * (1.0 / GRAD_SCALE)).contiguous().view(w.shape)
        else:
            g = ops.lrelu_bwd_h16(gy.contiguous(), y, ctx.P_out, ctx.L_out, ctx.slope)
        if ctx.needs_input_grad[0]:
            if stride == 1:   # the forward pack read as an MN-major operand, taps reversed
                _, dx, _ = ops.gemm(g.view(1, rows_out, cout), wp, epi=ops.EPI_BIAS, flags=ops.FLAG_OUT_H16 | ops.FLAG_NO_F32,
                                    pad=KSIZE - 1 - PAD, w_mn=True, tap_reverse=True, N=cin_p)
                dx = dx.view(rows_in, cin_p)
            else:
                # phase decomposition: one stride-1 GEMM over g writes the three interleaved input-row phases side by side
                wd, dpad = _conv_dgrad_pack(w, cin_p, stride, ctx.pkey)
                _, dxp, _ = ops.gemm(g.view(1, rows_out, cout), wd, epi=ops.EPI_BIAS, flags=ops.FLAG_OUT_H16 | ops.FLAG_NO_F32, pad=dpad)
                dx = dxp.view(rows_in, cin_p)
        return dx, dw, db, None, None, None, None, None, None, None


class _PostFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, period: int, L: int, P: int, s_pre=None):
        w2 = w.reshape(-1, 3).contiguous()
        score = s_pre if s_pre is not None else ops.mpd_post_fwd(x, w2, bias, period, L, P)
        ctx.save_for_backward(x, w2)
        ctx.geo, ctx.w_shape = (period, L, P), w.shape
        return score

    @staticmethod
    def backward(ctx, dscore):
        x, w2 = ctx.saved_tensors
        period, L, P = ctx.geo
        want_dw = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dx, dw, db = ops.mpd_post_bwd(dscore, x, w2, period, L, P, GRAD_SCALE, ctx.needs_input_grad[0], want_dw)
        return dx, (dw.reshape(ctx.w_shape) if dw is not None else None), db, None, None, None, None


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
        wp, key = None, _layer_key(conv)
        if 1 <= i <= 4:
            cout, cin = w.shape[0], w.shape[1]
            wp = _memo(("fwd",) + key, lambda: ops.pack_conv_h16(w.detach().reshape(cout, cin, KSIZE), k_pad=max(cin, C1_PAD)))
        layers.append((w, conv.bias, wp, key))
    return layers


def period_forward(disc, wav: torch.Tensor, layers=None, pre=None, record=None):
    """Native forward of one DiscriminatorP on (NS, T) fp32 signals: (score (NS, L5*period) fp32, [FlatMap x 4, score map]).
    `pre` = the six layer outputs computed earlier for the same signals and weights (the kernels are skipped, the autograd
    graph is built around them); `record` = a list that receives the six layer outputs."""
    if not wav.is_cuda:
        raise RuntimeError("the period discriminators run on the CUDA kernels only (no CPU path)")
    wav = wav.float().contiguous()
    NS, T = wav.shape
    geom = Geometry.make(T, disc.period)
    layers = layers if layers is not None else effective_weights(disc)
    slope = float(disc.lrelu_slope)
    (w1, b1, _, _), rest = layers[0], layers[1:5]
    x = _FirstFn.apply(wav, w1, b1, geom, 3, slope, pre[0] if pre else None)
    outs = [x]
    fmap: List[object] = []
    for i, (w, b, wp, key) in enumerate(rest, start=2):
        stride = 3 if i <= 4 else 1
        x = _ConvFn.apply(x, w, b, wp, stride, geom.P[i], geom.L[i], slope, pre[i - 1] if pre else None, key)
        outs.append(x)
        fmap.append(FlatMap(x, disc.period, geom.L[i], geom.P[i]))
    wpost, bpost = layers[5][0], layers[5][1]
    score = _PostFn.apply(x, wpost, bpost, disc.period, geom.L[5], geom.P[5], pre[5] if pre else None)
    outs.append(score)
    if record is not None:
        record.extend(t.detach() for t in outs)
    fmap.append(score.view(NS, 1, geom.L[5], disc.period))
    return score, fmap


REUSE_GENERATOR_TURN = True   # the discriminator turn reuses the layer outputs of the generator turn (see _pair below)
LAST_PAIR_REUSED = False      # whether the latest discriminator-turn call hit that cache (tests)


def _reuse_key(disc, y, y_hat):
    return (y.data_ptr(), y._version, tuple(y.shape), y_hat.data_ptr(), y_hat._version,
            tuple((p._version, getattr(p, "_osb_epoch", 0)) for p in disc.parameters()))


def _real_key(disc, y):
    return (y.data_ptr(), y._version, tuple(y.shape), tuple((p._version, getattr(p, "_osb_epoch", 0)) for p in disc.parameters()))


def prefetch_real(discs, y: torch.Tensor, forward, weights_of, slot0: int) -> None:
    """Generator turn, early: the real signals' forward pass (no graph, frozen weights) of every discriminator, queued on the
    discriminator's stream behind whatever produced `y` and NOT joined — the generator-turn call of `_pair` for the same
    `y` and the same weights, which runs on the same stream, takes the results instead of recomputing them."""
    dev = y.device
    cur = torch.cuda.current_stream(dev)
    for i, d in enumerate(discs):
        s = ops.side_stream(dev, slot0 + i) if PARALLEL_DISCRIMINATORS else cur
        if s != cur:
            s.wait_stream(cur)
        with torch.cuda.stream(s), torch.no_grad():
            rec = []
            sr, fr = forward(d, y, weights_of(d), record=rec)
            d.__dict__["_real_prefetch"] = (_real_key(d, y), sr, fr, rec)


def _pair(disc, y, y_hat, forward, weights, detach):
    """(score_real, score_fake, fmap_real, fmap_fake) of one discriminator.

    Generator turn (y_hat carries a graph, the discriminator weights are frozen): the real signals run without a graph, the
    generated ones alone are differentiated.  Discriminator turn: both halves share one pass over 2B signals.  With
    `cache_generator_outputs` (the reference default) the discriminator turn sees the SAME waveforms and the SAME weights as the
    generator turn of the step (the discriminator optimizer steps afterwards), so its forward values are the ones the
    generator turn already produced: they are kept (keyed by the tensors' storage, versions and every parameter version) and the
    discriminator turn only builds its autograd graph around them — one of the step's three discriminator forwards is not
    recomputed.  The reference recomputes it; the values are identical."""
    gen_turn = torch.is_grad_enabled() and y_hat.requires_grad
    ahead = disc.__dict__.pop("_real_prefetch", None)
    if gen_turn:
        rec_r, rec_g = [], []
        if ahead is not None and ahead[0] == _real_key(disc, y):
            _, sr, fr, rec_r = ahead
        else:
            with torch.no_grad():
                sr, fr = forward(disc, y, detach(weights), record=rec_r)
        sg, fg = forward(disc, y_hat, weights, record=rec_g)
        if REUSE_GENERATOR_TURN:
            disc.__dict__["_turn_cache"] = (_reuse_key(disc, y, y_hat), rec_r, rec_g)
        return sr, sg, fr, fg
    global LAST_PAIR_REUSED
    B = y.shape[0]
    pre = None
    cached = disc.__dict__.pop("_turn_cache", None)
    if cached is not None and REUSE_GENERATOR_TURN and cached[0] == _reuse_key(disc, y, y_hat):
        pre = [tuple(torch.cat((a, b)) for a, b in zip(r, g)) if isinstance(r, tuple) else torch.cat((r, g)) for r, g in zip(cached[1], cached[2])]
    LAST_PAIR_REUSED = pre is not None
    s, f = forward(disc, torch.cat((y.float(), y_hat.float()), dim=0), weights, pre=pre)
    fr = [m.half(0) if isinstance(m, FlatMap) else m[:B] for m in f]
    fg = [m.half(1) if isinstance(m, FlatMap) else m[B:] for m in f]
    return s[:B], s[B:], fr, fg


DISC_STREAM_SLOT0 = 20       # side-stream slots 20.. : one per discriminator (default priority)
PARALLEL_DISCRIMINATORS = True
STREAM_SET = 0               # 1 while the discriminator turn of an overlapped step runs (BaseModule): its own set of streams
_DEFERRED_JOINS: list = []   # side streams of fan-outs whose join was left to the caller (deferred_join)


class deferred_join:
    """Context manager: fan-outs inside it do not join their streams on return; the join happens when the context ends.
    VocosDiscriminator uses it so that the resolution discriminators start next to the period discriminators instead of behind
    their join."""

    def __enter__(self):
        self.depth0 = len(_DEFERRED_JOINS)
        _DEFERRED_JOINS.append(None)   # marker: a context is open
        return self

    def __exit__(self, *exc):
        pending = _DEFERRED_JOINS[self.depth0 + 1:]
        del _DEFERRED_JOINS[self.depth0:]
        for s in pending:
            torch.cuda.current_stream(s.device).wait_stream(s)
        return False


def fan_out(discs, y: torch.Tensor, y_hat: torch.Tensor, pair_fn, slot0: int):
    """Runs `pair_fn(d, y, y_hat)` for every discriminator on the discriminator's OWN side stream and joins them.

    The five period and three resolution discriminators are independent chains of ~40 launches each whose small kernels
    (layer 1, conv_post, gates, gathers, weight-norm, packs) and GEMM wave tails leave most SMs idle when they run one after
    the other; on separate streams they fill each other's gaps.  Autograd replays a node's backward on the stream its forward
    ran on, so the backward chains overlap the same way, and inside a captured step the forks / joins are graph edges.  A
    discriminator keeps its slot, so the layer outputs it leaves for the discriminator turn (`_turn_cache`) and its memoised
    weight packs are produced and consumed in stream order."""
    dev = y.device
    cur = torch.cuda.current_stream(dev)
    outs, streams = [], []
    for i, d in enumerate(discs):
        s = ops.side_stream(dev, slot0 + 20 * STREAM_SET + i) if PARALLEL_DISCRIMINATORS else cur
        if s != cur:
            s.wait_stream(cur)
            streams.append(s)
        with torch.cuda.stream(s):
            outs.append(pair_fn(d, y, y_hat))
    if _DEFERRED_JOINS:
        _DEFERRED_JOINS.extend(streams)
    else:
        for s in streams:
            cur.wait_stream(s)
    return outs


def period_forward_pair(disc, y: torch.Tensor, y_hat: torch.Tensor):
    return _pair(disc, y, y_hat, period_forward, effective_weights(disc), lambda ws: [(w.detach(), b.detach(), wp, k) for w, b, wp, k in ws])


# --------------------------------------------------------------------------------------------------
# Resolution discriminators (reference _discriminators.py:139-216) on the same flat layout: sequences are (signal, frame)
# pairs, rows run along frequency.  `FlatMap(data, period=W, L=H, P)` and its `dense()` -> (N, 64, H, W) apply unchanged.
# --------------------------------------------------------------------------------------------------
MRD_LAYERS = ((7, 5, 2, 2, 3, 2), (5, 3, 2, 1, 2, 1), (5, 3, 2, 2, 2, 1), (3, 3, 2, 1, 1, 1), (3, 3, 2, 2, 1, 1))  # kh, kw, sh, sw, ph, pw


@dataclass(frozen=True)
class GeometryR:
    H: Tuple[int, ...]   # valid frequency rows: H[0] = bins of the spectrogram, H[1..5] layer outputs
    W: Tuple[int, ...]   # frames
    P: Tuple[int, ...]   # rows a (signal, frame) sequence owns in layer i's matrix (P[0] unused)

    @staticmethod
    def make(F_bins: int, frames: int) -> "GeometryR":
        H, W = [F_bins], [frames]
        for kh, kw, sh, sw, ph, pw in MRD_LAYERS:
            H.append((H[-1] + 2 * ph - kh) // sh + 1)
            W.append((W[-1] + 2 * pw - kw) // sw + 1)
        p5 = H[5] + 2
        P = (0, p5 * 16, p5 * 8, p5 * 4, p5 * 2, p5)
        for i in range(1, 5):
            assert P[i] >= H[i] + 2, (P, H)
        return GeometryR(tuple(H), tuple(W), P)


def _rfirst_dgrad_pack(w, pkey):
    return _memo(("dgrad",) + pkey if pkey else None,
                 lambda: F.pad(w.detach().reshape(64, 35).t(), (0, 0, 0, 29)).to(torch.float16).contiguous().view(1, 64, 64))   # [tap][cout]


class _RFirstFn(torch.autograd.Function):
    """Layer 1 (Conv2d(1, 64, (7,5), (2,2), (3,2)) + LeakyReLU) as a GEMM: the 35 taps of every output position are gathered
    into a 64-wide fp16 row (osb_spec_im2col_h16), which is also the operand of the weight gradient."""

    @staticmethod
    def forward(ctx, spec, xcol, w, bias, geom: GeometryR, slope: float, y_pre=None, pkey=None):
        """`spec` only routes the gradient (it may be None when nothing upstream needs one); `xcol` is its tap gather."""
        ctx.pkey = pkey
        rows = xcol.shape[0]
        if pkey is not None and spec is not None and ctx.needs_input_grad[0]:
            _rfirst_dgrad_pack(w, pkey)
        if y_pre is not None:
            y = y_pre
        else:
            wp = _memo(("fwd",) + pkey if pkey else None, lambda: F.pad(w.detach().reshape(64, 35), (0, 29)).to(torch.float16).view(1, 64, 64))
            _, y, _ = ops.gemm(xcol.view(1, rows, 64), wp, epi=ops.EPI_BIAS, flags=ops.FLAG_OUT_H16 | ops.FLAG_NO_F32 | ops.FLAG_KEEPMASK,
                               bias=bias, seq_rows=(geom.P[1], geom.H[1]), lrelu=slope)
            y = y.view(rows, 64)
        ctx.save_for_backward(xcol, w, y)
        ctx.geom, ctx.slope, ctx.spec_shape = geom, slope, (tuple(spec.shape) if spec is not None else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        xcol, w, y = ctx.saved_tensors
        geom = ctx.geom
        rows = y.shape[0]
        dspec = dw = db = None
        if ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:
            g, db = ops.lrelu_bwd_h16(gy.contiguous(), y, geom.P[1], geom.H[1], ctx.slope, colsum_scale=1.0 / GRAD_SCALE)
            dwp = torch.zeros((1, 64, 64), device=g.device, dtype=torch.float32)
            ops.gemm_wgrad(g.view(1, rows, 64), xcol.view(1, rows, 64), dwp)
            dw = (dwp[0, :, :35] * (1.0 / GRAD_SCALE)).reshape(w.shape)
        else:
            g = ops.lrelu_bwd_h16(gy.contiguous(), y, geom.P[1], geom.H[1], ctx.slope)
        if ctx.needs_input_grad[0] and ctx.spec_shape is not None:
            wd = _rfirst_dgrad_pack(w, ctx.pkey)
            _, col, _ = ops.gemm(g.view(1, rows, 64), wd, epi=ops.EPI_BIAS, flags=ops.FLAG_OUT_H16 | ops.FLAG_NO_F32)
            dspec = ops.spec_col2im(col.view(rows, 64), ctx.spec_shape, geom.H[1], geom.W[1], geom.P[1], 1.0 / GRAD_SCALE)
        return dspec, None, dw, db, None, None, None, None


def _phase_dgrad_pack(w: torch.Tensor, ph: int, stride: int = 2):
    """Data gradient of a strided convolution along the rows as ONE stride-1 GEMM over the output-gradient rows: input row
    stride*m + phi collects g[m + d] . W[kh] for the taps with phi + ph - kh = stride*d.  All phases share the tap offsets d
    and sit side by side in the N dimension, so the (rows_out, stride*K) result IS the (stride*rows_out, K) input-gradient
    matrix — no per-tap product buffer, no scatter.   w (cout, cin, kh, kw) -> (fp16 (taps, stride*kw*cin, cout), pad)."""
    cout, cin, kh, kw = w.shape
    terms = [(phi, k, (phi + ph - k) // stride) for phi in range(stride) for k in range(kh) if (phi + ph - k) % stride == 0]
    d_min, d_max = min(t[2] for t in terms), max(t[2] for t in terms)
    wd = torch.zeros((d_max - d_min + 1, stride, kw, cin, cout), device=w.device, dtype=torch.float32)
    for phi, k, d in terms:
        wd[d - d_min, phi] = w[:, :, k, :].permute(2, 1, 0)
    return wd.reshape(d_max - d_min + 1, stride * kw * cin, cout).to(torch.float16).contiguous(), -d_min


class _RConvFn(torch.autograd.Function):
    """One Conv2d(64, 64, (kh, kw), (2, sw)) + LeakyReLU: x (NS*W_in*P_in, 64) -> y (NS*W_out*P_out, 64)."""

    @staticmethod
    def forward(ctx, x, w, bias, NS: int, W_in: int, W_out: int, P_in: int, P_out: int, H_out: int, layer, slope: float, y_pre=None,
                pkey=None):
        kh, kw, _sh, sw, ph, pw = layer
        ctx.pkey = pkey
        if pkey is not None and ctx.needs_input_grad[0]:
            _memo(("dgrad",) + pkey, lambda: _phase_dgrad_pack(w.detach(), ph))
        if y_pre is not None:
            y = y_pre
        else:
            xcol = ops.wim2col_h16(x, NS, W_in, W_out, P_in, kw, pw, sw)
            cout, cin = w.shape[0], w.shape[1]
            wp = _memo(("fwd",) + pkey if pkey else None,
                       lambda: w.detach().permute(2, 0, 3, 1).reshape(kh, cout, kw * cin).to(torch.float16).contiguous())   # [kh][cout][(kw, cin)]
            rows_in = xcol.shape[0]
            _, y, _ = ops.gemm(xcol.view(1, rows_in, kw * cin), wp, epi=ops.EPI_BIAS, flags=ops.FLAG_OUT_H16 | ops.FLAG_NO_F32 | ops.FLAG_KEEPMASK,
                               pad=ph, bias=bias, seq_rows=(P_out, H_out), row_stride=2, lrelu=slope)
            y = y.view(rows_in // 2, cout)
        ctx.save_for_backward(x, w, y)
        ctx.geo = (NS, W_in, W_out, P_in, P_out, H_out, layer, slope)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        NS, W_in, W_out, P_in, P_out, H_out, layer, slope = ctx.geo
        kh, kw, _sh, sw, ph, pw = layer
        cout, cin = w.shape[0], w.shape[1]
        rows_out = y.shape[0]
        dx = dw = db = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            g, db = ops.lrelu_bwd_h16(gy.contiguous(), y, P_out, H_out, slope, colsum_scale=1.0 / GRAD_SCALE)
            xcol = ops.wim2col_h16(x, NS, W_in, W_out, P_in, kw, pw, sw)      # recomputed: cheaper than keeping 3x the activation
            dwp = torch.zeros((kh, cout, kw * cin), device=x.device, dtype=torch.float32)
            ops.gemm_wgrad_strided(g, xcol, dwp, taps=kh, pad=ph, stride=2)
            dw = (dwp.view(kh, cout, kw, cin).permute(1, 3, 0, 2) * (1.0 / GRAD_SCALE)).contiguous()
        else:
            g = ops.lrelu_bwd_h16(gy.contiguous(), y, P_out, H_out, slope)
        if ctx.needs_input_grad[0]:
            wd, pad = _memo(("dgrad",) + ctx.pkey if ctx.pkey else None, lambda: _phase_dgrad_pack(w.detach(), ph))
            _, dxcol, _ = ops.gemm(g.view(1, rows_out, cout), wd, epi=ops.EPI_BIAS, flags=ops.FLAG_OUT_H16 | ops.FLAG_NO_F32, pad=pad)
            dx = ops.wcol2im_h16(dxcol.view(2 * rows_out, kw * cin), NS, W_in, W_out, P_in, cin, kw, pw, sw)
        return (dx, dw, db) + (None,) * 10


class _RPostFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, NS: int, W: int, H: int, P: int, s_pre=None):
        w2 = w.reshape(64, 9).contiguous()
        score = s_pre if s_pre is not None else ops.mrd_post_fwd(x, w2, bias, NS, W, H, P)
        ctx.save_for_backward(x, w2)
        ctx.geo, ctx.w_shape = (NS, W, H, P), w.shape
        return score

    @staticmethod
    def backward(ctx, dscore):
        x, w2 = ctx.saved_tensors
        NS, W, H, P = ctx.geo
        want_dw = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dx, dw, db = ops.mrd_post_bwd(dscore, x, w2, NS, W, H, P, GRAD_SCALE, ctx.needs_input_grad[0], want_dw)
        return dx, (dw.reshape(ctx.w_shape) if dw is not None else None), db, None, None, None, None, None


def _resolution_weights(disc):
    return [(torch._weight_norm(c.weight_v, c.weight_g, 0), c.bias, _layer_key(c)) for c in list(disc.convs) + [disc.conv_post]]


def _stft_frames(T: int, hop: int) -> int:
    return T // hop + 1      # torch.stft(center=True)


def resolution_forward(disc, wav: torch.Tensor, weights=None, pre=None, record=None):
    """Native forward of one DiscriminatorR on (NS, T) signals: (score (NS, H5*W5) fp32, [FlatMap x 5, score map]).
    `pre` / `record` as in period_forward (the first entry is the pair (tap gather, layer-1 output))."""
    if not wav.is_cuda:
        raise RuntimeError("the resolution discriminators run on the CUDA kernels only (no CPU path)")
    NS = wav.shape[0]
    if pre is None:
        spec = disc.spectrogram(wav.float()).contiguous()          # rectangular-window |STFT| (torch.stft / cuFFT), (NS, F, W)
        F_bins, frames = spec.shape[1], spec.shape[2]
    else:
        spec = None
        n_fft, hop, _win = disc.resolution
        F_bins, frames = n_fft // 2 + 1, _stft_frames(wav.shape[1], hop)
    geom = GeometryR.make(F_bins, frames)
    if weights is None:
        weights = _resolution_weights(disc)
    slope = float(disc.lrelu_slope)
    xcol1 = pre[0][0] if pre else ops.spec_im2col_h16(spec, geom.H[1], geom.W[1], geom.P[1])
    x = _RFirstFn.apply(spec, xcol1, weights[0][0], weights[0][1], geom, slope, pre[0][1] if pre else None, weights[0][2])
    outs = [x]
    fmap: List[object] = [FlatMap(x, geom.W[1], geom.H[1], geom.P[1])]
    for i in range(2, 6):
        w, b, key = weights[i - 1]
        x = _RConvFn.apply(x, w, b, NS, geom.W[i - 1], geom.W[i], geom.P[i - 1], geom.P[i], geom.H[i], MRD_LAYERS[i - 1], slope,
                           pre[i - 1] if pre else None, key)
        outs.append(x)
        fmap.append(FlatMap(x, geom.W[i], geom.H[i], geom.P[i]))
    wpost, bpost = weights[5][0], weights[5][1]
    score = _RPostFn.apply(x, wpost, bpost, NS, geom.W[5], geom.H[5], geom.P[5], pre[5] if pre else None)
    outs.append(score)
    if record is not None:
        record.append((xcol1.detach(), outs[0].detach()))
        record.extend(t.detach() for t in outs[1:])
    fmap.append(score.view(NS, 1, geom.H[5], geom.W[5]))
    return score, fmap


def resolution_forward_pair(disc, y: torch.Tensor, y_hat: torch.Tensor):
    """As period_forward_pair, for one resolution discriminator."""
    return _pair(disc, y, y_hat, resolution_forward, _resolution_weights(disc), lambda ws: [(w.detach(), b.detach(), k) for w, b, k in ws])
