"""Tensor-level wrappers over the C ABI (include/osb200.h).

Each function takes CUDA tensors, allocates outputs with torch (device memory and streams are
PyTorch's job here), and enqueues the library's kernels on the current stream.  There is no
fallback: non-CUDA tensors or a missing library raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import (EPI_ATTN_LOGP, EPI_AXPY, EPI_BIAS, EPI_BIAS_LN, EPI_GELU, EPI_GELU_BWD, EPI_LN_BWD, EPI_RELU, EPI_RELU_BWD, EPI_RELU_LN,
                   EPI_RELU_LN_BWD, EPI_RESID, FLAG_CLIP, FLAG_DOT, FLAG_KEEPMASK, FLAG_OUT_H16, FLAG_RELU, FLAG_SAVE_PRE, FLAG_SPLIT_IN,
                   FLAG_SPLIT_OUT)

__all__ = [
    "gemm", "embed_text", "dwconv_ln", "layernorm", "variance_embed", "durations", "centres", "gaussian_upsample",
    "expand_gather", "pack_h16", "EPI_BIAS", "EPI_GELU", "EPI_RESID", "EPI_RELU_LN", "EPI_BIAS_LN", "EPI_RELU",
    "FLAG_CLIP", "FLAG_KEEPMASK", "FLAG_OUT_H16", "FLAG_SAVE_PRE", "FLAG_DOT", "FLAG_SPLIT_IN", "FLAG_SPLIT_OUT",
]


_STEP_COUNTERS = {}


def step_counter(device) -> torch.Tensor:
    """Per-device int64 scalar that the training step increments once per step (on the device).  Counter-based dropout adds it
    to its seed, so a captured CUDA graph draws fresh masks on every replay without any host-side value baked into the graph."""
    key = torch.device(device).index or 0
    t = _STEP_COUNTERS.get(key)
    if t is None:
        t = torch.zeros((), device=device, dtype=torch.int64)
        _STEP_COUNTERS[key] = t
    return t


_SIDE_STREAMS = {}
ATTN_SLOT = 10   # text-side alignment convs + attention: own high-priority stream (its backward then overlaps the predictors')
SIDE_STREAMS_ENABLED = True   # False: every branch stays on the current stream (per-kernel timing passes want serial execution)


def side_stream(device, slot: int = 0) -> "torch.cuda.Stream":
    """A cached auxiliary stream per device: long single-wave kernels (the forward-sum recursion: 32 CTAs for 0.4 ms) run
    there concurrently with the main stream's work.  Callers fork with wait_stream(current) and join before using results."""
    if not SIDE_STREAMS_ENABLED:
        return torch.cuda.current_stream(device)
    key = (torch.device(device).index or 0, slot)
    s = _SIDE_STREAMS.get(key)
    if s is None:
        # Branches on the step's critical path (alignment / predictor branches) get a high-priority stream; the gradient-free
        # decoder -> vocoder branch (slot 5) and the weight-gradient streams (slots >= 6) keep the default (lowest) priority:
        # their full-machine kernels (216 CTAs x 224 KB of shared memory) otherwise hold every SM while the small kernels of
        # the critical path wait for a free one (timeline of the captured step, round 2).  A captured kernel node inherits
        # the priority of the stream it was captured on.
        prio = -1 if (slot <= 4 or slot == ATTN_SLOT) else 0
        s = torch.cuda.Stream(device=device, priority=prio)
        _SIDE_STREAMS[key] = s
    return s


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None or isinstance(t, int):
        return t
    if not t.is_cuda:
        raise _lib.OsbError("optispeech_b200 ops need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise _lib.OsbError("optispeech_b200 ops need contiguous tensors")
    return t.data_ptr()


def _f32(t: torch.Tensor) -> torch.Tensor:
    assert t.dtype == torch.float32, t.dtype
    return t


def gemm(a: torch.Tensor, w: torch.Tensor, *, epi: int, flags: int = 0, pad: int = 0, K: Optional[int] = None,
         bias=None, out=None, aux=None, resid=None, gamma=None, row_scale=None, pad_mask=None, ln_w=None, ln_b=None,
         ln_eps: float = 0.0, dot_w=None, dot_b=None, out_dot=None, aux_in=None, row_stat=None, dropout_p: float = 0.0, dropout_seed: int = 0, w_batched: bool = False, col_len=None, N: Optional[int] = None, colsum=None, w_mn: bool = False, tap_reverse: bool = False,
         row_stride: int = 1, lrelu: Optional[float] = None, seq_rows: Optional[Tuple[int, int]] = None):
    """acc[b,t,n] = sum_tap sum_k a[b,t+tap-pad,k] w[tap,n,k]; then the fused epilogue `epi`.

    a: fp16 (B,T,lda); w: fp16 (taps,N,ldw) — or, with FLAG_SPLIT_IN, a = (B,T,[hi K|lo K]) and
    w = (2,taps,N,ldw).  fp16 outputs are (B,T,N), or (B,T,2N) = [hi|lo] with FLAG_SPLIT_OUT.
    Outputs are allocated when not given.  Returns (out, aux, out_dot)."""
    assert a.dtype == torch.float16 and w.dtype == torch.float16 and a.dim() == 3
    B, T, lda = a.shape
    if w_mn:
        # dgrad on the forward pack: w (taps, K, ldw) with the output index contiguous (a layer's forward pack (taps, N_fwd, K_fwd))
        if w.dim() == 4:
            w = w[0]
        taps, Kw, ldw = w.shape
        K = K if K is not None else min(lda, Kw)
        N = N if N is not None else ldw
        flags |= _lib.FLAG_W_MN | (_lib.FLAG_TAP_REVERSE if tap_reverse else 0)
    elif w_batched:
        assert w.dim() == 3 and w.shape[0] == B
        taps, ldw = 1, w.shape[2]
        N = N if N is not None else w.shape[1]
        K = K if K is not None else (min(lda, ldw) // 2 if (flags & FLAG_SPLIT_IN) else min(lda, ldw))
    elif flags & FLAG_SPLIT_IN:
        assert w.dim() == 4 and w.shape[0] == 2
        _, taps, N, ldw = w.shape
        K = K if K is not None else min(lda // 2, ldw)
    else:
        if w.dim() == 4:  # packed (2, taps, N, K): single-pass mode uses the hi parts only
            w = w[0]
        taps, N, ldw = w.shape
        K = K if K is not None else min(lda, ldw)
    T_in = T
    if row_stride > 1:   # strided convolution: T becomes the number of OUTPUT rows
        T = (T_in + 2 * pad - taps) // row_stride + 1
    if lrelu is not None:
        flags |= _lib.FLAG_LRELU
    f32_out = epi in (EPI_BIAS, EPI_RESID, EPI_BIAS_LN, EPI_LN_BWD, EPI_ATTN_LOGP, EPI_AXPY)
    hN = 2 * N if (flags & FLAG_SPLIT_OUT) else N
    if out is None and not (epi == EPI_RELU_LN and (flags & FLAG_DOT) and not (flags & FLAG_OUT_H16)) and not (flags & _lib.FLAG_NO_F32):
        out = torch.empty((B, T, N if f32_out else hN), device=a.device, dtype=torch.float32 if f32_out else torch.float16)
    if (flags & FLAG_OUT_H16) and aux is None:
        aux = torch.empty((B, T, hN), device=a.device, dtype=torch.float16)
    if (flags & FLAG_SAVE_PRE) and aux is None:
        aux = torch.empty((B, T, N), device=a.device, dtype=torch.float16)
    if ((flags & FLAG_DOT) or epi == EPI_ATTN_LOGP) and out_dot is None:
        out_dot = torch.empty((B, T), device=a.device, dtype=torch.float32)
    d = _lib.GemmDesc()
    d.a, d.w, d.lda, d.ldw = _ptr(a), _ptr(w), lda, ldw
    d.B, d.T, d.N, d.K, d.taps, d.pad = B, T, N, K, taps, pad
    d.epi, d.flags = epi, flags
    d.out, d.aux_h16, d.ldo = _ptr(out), _ptr(aux), N
    d.bias, d.resid, d.gamma, d.row_scale = _ptr(bias), _ptr(resid), _ptr(gamma), _ptr(row_scale)
    d.pad_mask = _ptr(pad_mask)
    d.ln_w, d.ln_b, d.ln_eps = _ptr(ln_w), _ptr(ln_b), ln_eps
    d.dot_w, d.dot_b, d.out_dot = _ptr(dot_w), _ptr(dot_b), _ptr(out_dot)
    d.aux_in_h16, d.row_stat = _ptr(aux_in), _ptr(row_stat)
    d.dropout_p, d.dropout_seed = float(dropout_p), int(dropout_seed)
    d.dropout_seed_dev = _ptr(step_counter(a.device)) if dropout_p > 0.0 else None
    d.w_batched, d.col_len = int(w_batched), _ptr(col_len)
    d.row_stride, d.T_in, d.lrelu_slope = int(row_stride), int(T_in), float(lrelu or 0.0)
    if seq_rows is not None:   # KEEPMASK by arithmetic: (pitch, valid rows) of the flat sequence layout
        d.seq_pitch, d.seq_valid = int(seq_rows[0]), int(seq_rows[1])
    if colsum is not None:  # bias gradient of the layer whose dgrad this is: column sums of the fp16 output, fused (pre-zeroed fp32 (N,))
        assert epi in (EPI_GELU_BWD, EPI_RELU_BWD, EPI_RELU_LN_BWD) and colsum.dtype == torch.float32 and colsum.numel() == N
        d.flags |= _lib.FLAG_COLSUM
        d.out_colsum = _ptr(colsum)
    _lib.check(_lib.load().osb_gemm(C.byref(d), _stream()), "osb_gemm")
    return out, aux, out_dot


def gemm_wgrad(dy: torch.Tensor, a: torch.Tensor, dw: torch.Tensor, *, taps: int = 1, pad: int = 0, K: Optional[int] = None):
    """dw[tap,n,k] += sum_{b,t} dy[b,t,n] a[b,t+tap-pad,k]; dy, a fp16 channels-last; dw fp32 (taps,N,K)."""
    assert dy.dtype == torch.float16 and a.dtype == torch.float16 and dw.dtype == torch.float32
    B, T, N = dy.shape
    K = K if K is not None else a.shape[2]
    _lib.check(_lib.load().osb_gemm_wgrad(_ptr(dy), dy.shape[2], _ptr(a), a.shape[2], _ptr(dw), B, T, N, K, taps, pad, _stream()),
               "osb_gemm_wgrad")
    return dw


def embed_text(ids, table, inv_freq, scale):
    B, T = ids.shape
    dim = table.shape[1]
    out = torch.empty((B, T, dim), device=ids.device, dtype=torch.float32)
    _lib.check(_lib.load().osb_embed_text(_ptr(ids), _ptr(_f32(table)), _ptr(_f32(inv_freq)), _ptr(_f32(scale)), _ptr(out), B, T, dim,
                                          table.shape[0], _stream()), "osb_embed_text")
    return out


def dwconv_ln(x, w, bias, eps: float, want_rstd: bool = False, split: bool = False):
    B, T, Cc = x.shape
    xhat = torch.empty((B, T, 2 * Cc if split else Cc), device=x.device, dtype=torch.float16)
    rstd = torch.empty((B, T), device=x.device, dtype=torch.float32) if want_rstd else None
    _lib.check(_lib.load().osb_dwconv_ln(_ptr(_f32(x)), _ptr(_f32(w)), _ptr(_f32(bias)), _ptr(xhat), _ptr(rstd), B, T, Cc, eps,
                                         int(split), _stream()), "osb_dwconv_ln")
    return xhat, rstd


def _h16_like(x, split: bool):
    return torch.empty((*x.shape[:-1], (2 if split else 1) * x.shape[-1]), device=x.device, dtype=torch.float16)


def layernorm(x, w, b, eps: float, f32: bool = True, h16: bool = False, split: bool = False):
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    o32 = torch.empty_like(x) if f32 else None
    o16 = _h16_like(x, split) if h16 else None
    _lib.check(_lib.load().osb_layernorm(_ptr(_f32(x)), _ptr(_f32(w)), _ptr(_f32(b)), _ptr(o32), _ptr(o16), rows, Cc, eps, int(split),
                                         _stream()), "osb_layernorm")
    return o32, o16


def relu_layernorm(x, w, b, eps: float, h16: bool = True, split: bool = False, dot_w=None, dot_b=None, pad_mask=None):
    """LayerNorm(relu(x)) * w + b per row -> (fp16 (…, C) or split (…, 2C) or None, fp32 (rows-shaped) masked dot or None)."""
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    o16 = _h16_like(x, split) if h16 else None
    od = torch.empty(x.shape[:-1], device=x.device, dtype=torch.float32) if dot_w is not None else None
    _lib.check(_lib.load().osb_relu_layernorm(_ptr(_f32(x)), _ptr(_f32(w)), _ptr(_f32(b)), _ptr(o16), rows, Cc, eps, int(split),
                                              _ptr(dot_w), _ptr(dot_b), _ptr(pad_mask), _ptr(od), _stream()), "osb_relu_layernorm")
    return o16, od


def variance_embed(x, val, w, bias, pad_mask, f32: bool = True, h16: bool = False, split: bool = False, emb_scale=None):
    B, T, Cc = x.shape
    k = w.shape[-1]
    o32 = torch.empty_like(x) if f32 else None
    o16 = _h16_like(x, split) if h16 else None
    _lib.check(_lib.load().osb_variance_embed(_ptr(_f32(x)), _ptr(_f32(val)), _ptr(_f32(w)), _ptr(_f32(bias)), _ptr(pad_mask), _ptr(emb_scale),
                                              _ptr(o32), _ptr(o16), B, T, Cc, k, int(split), _stream()), "osb_variance_embed")
    return o32, o16


def durations(log_d, pad_mask, factor: float, clip_val: float) -> Tuple[torch.Tensor, torch.Tensor]:
    B, T = log_d.shape
    dur = torch.empty((B, T), device=log_d.device, dtype=torch.int64)
    lengths = torch.empty((B,), device=log_d.device, dtype=torch.int64)
    _lib.check(_lib.load().osb_durations(_ptr(_f32(log_d)), _ptr(pad_mask), _ptr(dur), _ptr(lengths), B, T, float(factor), float(clip_val),
                                         _stream()), "osb_durations")
    return dur, lengths


def centres(dur, want_csum: bool = True):
    B, T = dur.shape
    assert dur.dtype in (torch.int64, torch.float32)
    c = torch.empty((B, T), device=dur.device, dtype=torch.float32)
    cs = torch.empty((B, T), device=dur.device, dtype=torch.int64) if want_csum else None
    _lib.check(_lib.load().osb_centres(_ptr(dur), 1 if dur.dtype == torch.int64 else 0, _ptr(c), _ptr(cs), B, T, _stream()), "osb_centres")
    return c, cs


def gaussian_upsample(hs, centres_, x_len, y_len, Tm: int, delta: float = 0.1, f32: bool = True, h16: bool = False):
    B, Tx, Cc = hs.shape
    o32 = torch.empty((B, Tm, Cc), device=hs.device, dtype=torch.float32) if f32 else None
    o16 = torch.empty((B, Tm, Cc), device=hs.device, dtype=torch.float16) if h16 else None
    _lib.check(_lib.load().osb_gaussian_upsample(_ptr(_f32(hs)), _ptr(centres_), _ptr(x_len), _ptr(y_len), _ptr(o32), _ptr(o16), B, Tx, Tm,
                                                 Cc, delta, _stream()), "osb_gaussian_upsample")
    return o32, o16


def gaussian_upsample_window(hs, centres_, x_len, y_len, win_start, Tm: int, W: int, halo: int, delta: float = 0.1):
    """Rows win_start[b] - halo + j (j < W) of gaussian_upsample's (B, Tm, C) result, zero outside [0, Tm): (B, W, C) fp32."""
    B, Tx, Cc = hs.shape
    out = torch.empty((B, W, Cc), device=hs.device, dtype=torch.float32)
    _lib.check(_lib.load().osb_gaussian_upsample_window(_ptr(_f32(hs)), _ptr(centres_), _ptr(x_len), _ptr(y_len),
                                                        _ptr(win_start.to(torch.int64).contiguous()), _ptr(out), None, B, Tx, int(Tm),
                                                        int(W), int(halo), Cc, delta, _stream()), "osb_gaussian_upsample_window")
    return out


def expand_gather(x, csum, Tm: int):
    B, Tx, Cc = x.shape
    out = torch.empty((B, Tm, Cc), device=x.device, dtype=torch.float32)
    idx = torch.empty((B, Tm), device=x.device, dtype=torch.int32)
    _lib.check(_lib.load().osb_expand_gather(_ptr(_f32(x)), _ptr(csum), _ptr(out), _ptr(idx), B, Tx, Tm, Cc, _stream()), "osb_expand_gather")
    return out, idx


def pack_h16(src: torch.Tensor, *, rows: int, cols: int, src_ld: int, src_cs: int = 1, dst_cols: Optional[int] = None, col_scale=None,
             out: Optional[torch.Tensor] = None, out_lo: Optional[torch.Tensor] = None, dst_rs: Optional[int] = None):
    """fp32 -> fp16 (hi) and optionally the rounding residual (lo), with source strides (elements), optional
    per-column scale and zero column padding up to dst_cols.  Destination row stride dst_rs (default dst_cols)."""
    dst_cols = dst_cols or cols
    dst_rs = dst_rs or dst_cols
    if out is None:
        out = torch.empty((rows, dst_cols), device=src.device, dtype=torch.float16)
    _lib.check(_lib.load().osb_pack_h16(_ptr(_f32(src)), src_ld, src_cs, _ptr(col_scale), _ptr(out), _ptr(out_lo), dst_rs, dst_cols, rows,
                                        cols, _stream()), "osb_pack_h16")
    return out


def to_h16(x: torch.Tensor, pad_to: Optional[int] = None, split: bool = False) -> torch.Tensor:
    """Channels-last fp32 activation -> fp16 operand copy: plain (.., Cp) or split (.., [hi Cp | lo Cp])."""
    Cc = x.shape[-1]
    Cp = pad_to or Cc
    rows = x.numel() // Cc
    if not split:
        return pack_h16(x, rows=rows, cols=Cc, src_ld=Cc, dst_cols=Cp).view(*x.shape[:-1], Cp)
    out = torch.empty((rows, 2 * Cp), device=x.device, dtype=torch.float16)
    pack_h16(x, rows=rows, cols=Cc, src_ld=Cc, dst_cols=Cp, dst_rs=2 * Cp, out=out, out_lo=out.data_ptr() + 2 * Cp)
    return out.view(*x.shape[:-1], 2 * Cp)


# ------------------------------------------------------------------------------------------------
# backward kernels
# ------------------------------------------------------------------------------------------------
class _ZeroPool:
    """Zero-initialised fp32 scratch for one backward call: the many small accumulators the backward kernels add into
    (bias / LayerNorm / weight gradients) are carved from one pre-zeroed block instead of one fill launch each."""

    def __init__(self, device, block: int):
        self.device, self.block = device, block
        self.buf, self.pos = None, 0

    def take(self, shape):
        n = 1
        for s in shape:
            n *= int(s)
        n4 = (n + 3) // 4 * 4  # 16-byte aligned carve-outs
        if self.buf is None or self.pos + n4 > self.buf.numel():
            self.buf = torch.zeros(max(n4, self.block), device=self.device, dtype=torch.float32)
            self.pos = 0
        out = self.buf[self.pos:self.pos + n].view(shape)
        self.pos += n4
        return out


_POOL: Optional[_ZeroPool] = None


def pooled(fn):
    """Decorator for autograd backward functions: zeros() inside the call share pre-zeroed blocks."""
    import functools

    @functools.wraps(fn)
    def wrapper(ctx, *grads):
        global _POOL
        dev = next((g.device for g in grads if g is not None), None)
        if dev is None or dev.type != "cuda":
            return fn(ctx, *grads)
        prev, _POOL = _POOL, _ZeroPool(dev, 1 << 20)
        try:
            return fn(ctx, *grads)
        finally:
            _POOL = prev

    return wrapper


def zeros(shape, like):
    shape = tuple(shape) if not isinstance(shape, int) else (shape,)
    if _POOL is not None and _POOL.device == like.device:
        return _POOL.take(shape)
    return torch.zeros(shape, device=like.device, dtype=torch.float32)


_zeros = zeros


def resid_bwd_prep(dout, z_h16, gamma, pad_mask, row_scale, T: int):
    Cc = dout.shape[-1]
    rows = dout.numel() // Cc
    dyg = torch.empty(dout.shape, device=dout.device, dtype=torch.float16)
    dgamma, db2 = _zeros((Cc,), dout), _zeros((Cc,), dout)
    _lib.check(_lib.load().osb_resid_bwd_prep(_ptr(_f32(dout)), _ptr(z_h16), _ptr(gamma), _ptr(pad_mask), _ptr(row_scale), _ptr(dyg),
                                              _ptr(dgamma), _ptr(db2), rows, T, Cc, _stream()), "osb_resid_bwd_prep")
    return dyg, dgamma, db2


def colsum_h16(x_h16):
    N = x_h16.shape[-1]
    rows = x_h16.numel() // N
    out = _zeros((N,), x_h16)
    _lib.check(_lib.load().osb_colsum_h16(_ptr(x_h16), _ptr(out), rows, N, _stream()), "osb_colsum_h16")
    return out


def ln_fold_bwd(dw1f, w1, ln_w, ln_b, db1):
    I, Cc = w1.shape
    dln_w, dln_b = _zeros((Cc,), w1), _zeros((Cc,), w1)
    _lib.check(_lib.load().osb_ln_fold_bwd(_ptr(dw1f), _ptr(_f32(w1)), _ptr(ln_w), _ptr(ln_b), _ptr(db1), _ptr(dln_w), _ptr(dln_b), I, Cc,
                                           _stream()), "osb_ln_fold_bwd")
    return dw1f, dln_w, dln_b


def dwconv_bwd(dd, dout, x, w, pad_mask):
    B, T, Cc = x.shape
    dx = torch.empty_like(x)
    ddw, ddb = _zeros((Cc, 7), x), _zeros((Cc,), x)
    _lib.check(_lib.load().osb_dwconv_bwd(_ptr(_f32(dd)), _ptr(_f32(dout)), _ptr(_f32(x)), _ptr(_f32(w)), _ptr(pad_mask), _ptr(dx),
                                          _ptr(ddw), _ptr(ddb), B, T, Cc, _stream()), "osb_dwconv_bwd")
    return dx, ddw, ddb


def layernorm_bwd(dy, x, w, eps: float):
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    dx = torch.empty_like(x)
    dw, db = _zeros((Cc,), x), _zeros((Cc,), x)
    _lib.check(_lib.load().osb_layernorm_bwd(_ptr(_f32(dy)), _ptr(_f32(x)), _ptr(_f32(w)), _ptr(dx), _ptr(dw), _ptr(db), rows, Cc, eps,
                                             _stream()), "osb_layernorm_bwd")
    return dx, dw, db


def predictor_tail_bwd(d_out, pad_mask, r_h16, ln_w, ln_b, lin_w, eps: float, dropout_p: float = 0.0, dropout_seed: int = 0):
    Cc = r_h16.shape[-1]
    rows = r_h16.numel() // Cc
    g = torch.empty(r_h16.shape, device=r_h16.device, dtype=torch.float16)
    dlin_w, dlin_b, dln_w, dln_b = _zeros((Cc,), r_h16), _zeros((1,), r_h16), _zeros((Cc,), r_h16), _zeros((Cc,), r_h16)
    _lib.check(_lib.load().osb_predictor_tail_bwd(_ptr(_f32(d_out)), _ptr(pad_mask), _ptr(r_h16), _ptr(ln_w), _ptr(ln_b), _ptr(lin_w),
                                                  _ptr(g), _ptr(dlin_w), _ptr(dlin_b), _ptr(dln_w), _ptr(dln_b), rows, Cc, eps,
                                                  float(dropout_p), int(dropout_seed),
                                                  _ptr(step_counter(r_h16.device)) if dropout_p > 0.0 else None, _stream()),
               "osb_predictor_tail_bwd")
    return g, dlin_w, dlin_b, dln_w, dln_b


def ln_param_grad(gy_h16, r_h16, ln_w, eps: float):
    Cc = r_h16.shape[-1]
    rows = r_h16.numel() // Cc
    dln_w, dln_b = _zeros((Cc,), r_h16), _zeros((Cc,), r_h16)
    _lib.check(_lib.load().osb_ln_param_grad(_ptr(gy_h16), _ptr(r_h16), _ptr(ln_w), _ptr(dln_w), _ptr(dln_b), rows, Cc, eps, _stream()),
               "osb_ln_param_grad")
    return dln_w, dln_b


def variance_embed_bwd(dout, val, pad_mask, ksize: int, want_dx: bool = True, emb_scale=None):
    B, T, Cc = dout.shape
    dx = torch.empty_like(dout) if want_dx else None
    dw, db = _zeros((Cc, ksize), dout), _zeros((Cc,), dout)
    _lib.check(_lib.load().osb_variance_embed_bwd(_ptr(_f32(dout)), _ptr(_f32(val)), _ptr(pad_mask), _ptr(emb_scale), _ptr(dx), _ptr(dw), _ptr(db),
                                                  B, T, Cc, ksize, _stream()), "osb_variance_embed_bwd")
    return dx, dw, db


def embed_text_bwd(dout, ids, inv_freq, n_vocab: int, padding_idx: int):
    B, T, dim = dout.shape
    dtable, dscale = _zeros((n_vocab, dim), dout), _zeros((1,), dout)
    _lib.check(_lib.load().osb_embed_text_bwd(_ptr(_f32(dout)), _ptr(ids), _ptr(inv_freq), _ptr(dtable), _ptr(dscale), B, T, dim, n_vocab,
                                              padding_idx, _stream()), "osb_embed_text_bwd")
    return dtable, dscale


def mas(log_p_attn, x_len, m_len):
    """-> (path (B,Tm) int32, durations (B,Tx) fp32); bit-exact monotonic alignment search on the device."""
    B, Tm, Tx = log_p_attn.shape
    path = torch.empty((B, Tm), device=log_p_attn.device, dtype=torch.int32)
    dur = torch.empty((B, Tx), device=log_p_attn.device, dtype=torch.float32)
    _lib.check(_lib.load().osb_mas(_ptr(_f32(log_p_attn)), _ptr(x_len), _ptr(m_len), _ptr(path), _ptr(dur), B, Tm, Tx, _stream()), "osb_mas")
    return path, dur


def average_by_duration(ds, xs, x_len, m_len):
    B, Tx = ds.shape
    Tm = xs.shape[1]
    out = torch.empty((B, Tx), device=ds.device, dtype=torch.float32)
    _lib.check(_lib.load().osb_average_by_duration(_ptr(_f32(ds)), _ptr(_f32(xs)), _ptr(x_len), _ptr(m_len), _ptr(out), B, Tm, Tx, _stream()),
               "osb_average_by_duration")
    return out


def gemm_wgrad_batched(dy: torch.Tensor, a: torch.Tensor, dw: torch.Tensor, *, N: int, K: int):
    """dw[b,n,k] += sum_t dy[b,t,n] a[b,t,k]; dy fp16 (B,T,ldy), a fp16 (B,T,lda), dw fp32 (B,N,K)."""
    B, T, ldy = dy.shape
    _lib.check(_lib.load().osb_gemm_wgrad_batched(_ptr(dy), ldy, _ptr(a), a.shape[2], _ptr(dw), B, T, N, K, _stream()), "osb_gemm_wgrad_batched")
    return dw


def rownorm_sq(x):
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    out = torch.empty(x.shape[:-1], device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().osb_rownorm_sq(_ptr(_f32(x)), _ptr(out), rows, Cc, _stream()), "osb_rownorm_sq")
    return out


def attn_bwd_prep(G, lp, prior, lse, x_len, m_len):
    B, Tm, Tx = lp.shape
    ldw = (Tx + 7) // 8 * 8
    wn = torch.empty((B, Tm, ldw), device=lp.device, dtype=torch.float16)
    neg_rsn = torch.empty((B, Tm), device=lp.device, dtype=torch.float32)
    csn = torch.zeros((B, Tx), device=lp.device, dtype=torch.float32)
    _lib.check(_lib.load().osb_attn_bwd_prep(_ptr(_f32(G)), _ptr(lp), _ptr(prior), _ptr(lse), _ptr(x_len), _ptr(m_len), _ptr(wn),
                                             _ptr(neg_rsn), _ptr(csn), B, Tm, Tx, ldw, _stream()), "osb_attn_bwd_prep")
    return wn, neg_rsn, csn


def transpose_pack_h16(x, Tp: Optional[int] = None):
    B, T, Cc = x.shape
    Tp = Tp or (T + 7) // 8 * 8
    out = torch.empty((B, Cc, Tp), device=x.device, dtype=torch.float16)
    _lib.check(_lib.load().osb_transpose_pack_h16(_ptr(_f32(x)), _ptr(out), B, T, Cc, Tp, _stream()), "osb_transpose_pack_h16")
    return out


def scale_rows(x, row_scale, sign: float = 1.0):
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    out = torch.empty_like(x)
    _lib.check(_lib.load().osb_scale_rows(_ptr(_f32(x)), _ptr(row_scale), _ptr(out), rows, Cc, float(sign), _stream()), "osb_scale_rows")
    return out


def beta_binomial_prior(log_factorial, x_len, m_len, Tm: int, Tx: int):
    """(B,Tm,Tx) fp32 log-prior from the float64 log-factorial table (device-side, no host knowledge of the lengths)."""
    B = x_len.shape[0]
    assert log_factorial.dtype == torch.float64
    out = torch.empty((B, Tm, Tx), device=x_len.device, dtype=torch.float32)
    _lib.check(_lib.load().osb_beta_binomial_prior(_ptr(log_factorial), log_factorial.numel(), _ptr(x_len), _ptr(m_len), _ptr(out), B, Tm,
                                                   Tx, _stream()), "osb_beta_binomial_prior")
    return out


def forward_sum(log_p_attn, x_len, m_len, blank_logit: float = -1.0):
    """-> (per-sample loss (B,), grad (B,Tm,Tx) of sum(loss)/B w.r.t. log_p_attn)."""
    B, Tm, Tx = log_p_attn.shape
    ws = torch.empty((2 * B * Tm * Tx + B * Tm + B,), device=log_p_attn.device, dtype=torch.float32)
    loss = torch.empty((B,), device=log_p_attn.device, dtype=torch.float32)
    grad = torch.empty((B, Tm, Tx), device=log_p_attn.device, dtype=torch.float32)
    _lib.check(_lib.load().osb_forward_sum(_ptr(_f32(log_p_attn)), _ptr(x_len), _ptr(m_len), float(blank_logit), _ptr(ws), _ptr(loss),
                                           _ptr(grad), B, Tm, Tx, _stream()), "osb_forward_sum")
    return loss, grad


def pack_conv_h16(w: torch.Tensor, k_pad: Optional[int] = None, transpose_reverse: bool = False, split: bool = False):
    """Conv1d weight (N, Cin, k) fp32 -> fp16 operand in one launch: (k, N, Kp) forward form or (k, Cin, N) tap-reversed
    dgrad form; with split -> (2, ...) hi / lo."""
    N, Cin, k = w.shape
    Kp = k_pad or Cin
    shape = (k, Cin, N) if transpose_reverse else (k, N, Kp)
    out = torch.empty(((2,) + shape) if split else shape, device=w.device, dtype=torch.float16)
    lo = out[1] if split else None
    dst = out[0] if split else out
    _lib.check(_lib.load().osb_pack_conv_h16(_ptr(_f32(w.contiguous())), _ptr(dst), _ptr(lo), N, Cin, k, Kp, int(transpose_reverse),
                                             _stream()), "osb_pack_conv_h16")
    return out


FUSED_BLOCK_SHAPES = ((256, 1024), (384, 1152))


def convnext_block_fwd(x, dw_w, dw_b, w1f_h16, b1f, w2_h16, b2, gamma, row_scale=None, pad_mask=None, eps: float = 1e-6):
    """Fused ConvNeXt block forward (one launch).  x fp32 (B,T,C); w1f_h16 fp16 (I,C), w2_h16 fp16 (C,I)."""
    B, T, Cc = x.shape
    I = w1f_h16.shape[-2]
    out = torch.empty_like(x)
    _lib.check(_lib.load().osb_convnext_block_fwd(_ptr(_f32(x)), _ptr(_f32(dw_w)), _ptr(_f32(dw_b)), _ptr(w1f_h16), _ptr(_f32(b1f)),
                                                  _ptr(w2_h16), _ptr(_f32(b2)), _ptr(_f32(gamma)), _ptr(row_scale), _ptr(pad_mask),
                                                  _ptr(out), B, T, Cc, I, float(eps), _stream()), "osb_convnext_block_fwd")
    return out


def convnext_block_fwd_train(x, dw_w, dw_b, w1f_h16, b1f, w2_h16, b2, gamma, row_scale=None, pad_mask=None, eps: float = 1e-6):
    """Fused ConvNeXt block forward that also emits the activations the backward needs (one launch).
    -> (out fp32 (B,T,C), xhat fp16 (B,T,C), rstd fp32 (B,T), pre fp16 (B,T,I), h fp16 (B,T,I))."""
    B, T, Cc = x.shape
    I = w1f_h16.shape[-2]
    out = torch.empty_like(x)
    xhat = torch.empty((B, T, Cc), device=x.device, dtype=torch.float16)
    rstd = torch.empty((B, T), device=x.device, dtype=torch.float32)
    pre = torch.empty((B, T, I), device=x.device, dtype=torch.float16)
    h = torch.empty((B, T, I), device=x.device, dtype=torch.float16)
    _lib.check(_lib.load().osb_convnext_block_fwd_train(_ptr(_f32(x)), _ptr(_f32(dw_w)), _ptr(_f32(dw_b)), _ptr(w1f_h16), _ptr(_f32(b1f)),
                                                        _ptr(w2_h16), _ptr(_f32(b2)), _ptr(_f32(gamma)), _ptr(row_scale), _ptr(pad_mask),
                                                        _ptr(out), _ptr(xhat), _ptr(rstd), _ptr(pre), _ptr(h), B, T, Cc, I, float(eps),
                                                        _stream()), "osb_convnext_block_fwd_train")
    return out, xhat, rstd, pre, h


def convnext_block_bwd(dout, gamma, row_scale, pad_mask, pre, w2_h16, w1f_h16):
    """Both data-gradient contractions of a ConvNeXt block in one launch.
    -> (dyg fp16 (B,T,C), dh fp16 (B,T,I), dxhat fp32 (B,T,C))."""
    B, T, Cc = dout.shape
    I = pre.shape[-1]
    dyg = torch.empty((B, T, Cc), device=dout.device, dtype=torch.float16)
    dh = torch.empty((B, T, I), device=dout.device, dtype=torch.float16)
    parts = int(_lib.load().osb_convnext_block_bwd_parts(B, T, I))
    dxh = torch.empty((parts, B, T, Cc), device=dout.device, dtype=torch.float32)   # one partial sum per split of I
    _lib.check(_lib.load().osb_convnext_block_bwd(_ptr(_f32(dout)), _ptr(_f32(gamma)), _ptr(row_scale), _ptr(pad_mask), _ptr(pre), _ptr(w2_h16),
                                                  _ptr(w1f_h16), _ptr(dyg), _ptr(dh), _ptr(dxh), B, T, Cc, I, _stream()),
               "osb_convnext_block_bwd")
    return dyg, dh, dxh


def ln_dwconv_bwd(dxh, xhat, rstd, dout, x, dw_w, pad_mask):
    """LayerNorm backward + depthwise conv backward + residual path -> (dx, ddw (C,7), ddb (C))."""
    B, T, Cc = x.shape
    dx = torch.empty_like(x)
    dparam = _zeros((8, Cc), x)      # rows 0-6: the 7 taps (tap-major), row 7: the bias
    nparts = dxh.shape[0] if dxh.dim() == 4 else 1
    _lib.check(_lib.load().osb_ln_dwconv_bwd(_ptr(_f32(dxh)), nparts, _ptr(xhat), _ptr(_f32(rstd)), _ptr(_f32(dout)), _ptr(_f32(x)), _ptr(_f32(dw_w)),
                                             _ptr(pad_mask), _ptr(dx), _ptr(dparam), B, T, Cc, _stream()), "osb_ln_dwconv_bwd")
    return dx, dparam[:7].t(), dparam[7]


def resid_param_grad(dout, out, x, gamma, pad_mask, row_scale):
    """-> (dgamma (C), db2 (C)) of the residual epilogue from the block's input and output (no saved pwconv2 output)."""
    B, T, Cc = x.shape
    dgamma, db2 = _zeros((Cc,), x), _zeros((Cc,), x)
    _lib.check(_lib.load().osb_resid_param_grad(_ptr(_f32(dout)), _ptr(_f32(out)), _ptr(_f32(x)), _ptr(_f32(gamma)), _ptr(pad_mask),
                                                _ptr(row_scale), _ptr(dgamma), _ptr(db2), B * T, T, Cc, _stream()), "osb_resid_param_grad")
    return dgamma, db2


# ------------------------------------------------------------------------------------------------
# weight-gradient side streams: parameter gradients are only consumed by the optimizer at the end of the step, so the
# contractions / reductions that produce them leave the critical path (data gradients) of the backward pass
# ------------------------------------------------------------------------------------------------
_GRAD_PENDING = {"streams": [], "keep": [], "hooked": False, "rr": 0}
GRAD_SIDE_SLOTS = (6, 7, 8, 9)


class grad_side:
    """Context manager: run the enclosed weight-gradient work on a side stream forked from the current one.  The side stream
    is joined (and the tensors it reads released) when the running backward pass finishes — a callback queued on the autograd
    engine — or explicitly by join_grad_streams().  Inside a CUDA-graph capture the fork / join become graph dependencies.

    Pass the INPUT tensors of the enclosed work (allocated on the forking stream) to be kept alive until the join.  Do NOT
    hold extra references to the gradients it produces: autograd's AccumulateGrad then takes them over as they are; with
    another reference alive it would clone them — on the forking stream, unordered with the side stream that writes them."""

    def __init__(self, like: torch.Tensor, *keep):
        self.dev = like.device
        self.keep = keep
        self.enabled = SIDE_STREAMS_ENABLED and like.is_cuda

    def __enter__(self):
        if not self.enabled:
            return self
        st = _GRAD_PENDING
        slot = GRAD_SIDE_SLOTS[st["rr"] % len(GRAD_SIDE_SLOTS)]
        st["rr"] += 1
        self.side = side_stream(self.dev, slot)
        self.side.wait_stream(torch.cuda.current_stream(self.dev))
        self.ctx = torch.cuda.stream(self.side)
        self.ctx.__enter__()
        # zero-initialised accumulators requested inside the context come from a block that is allocated AND filled on the
        # side stream (a block filled on the forking stream after the fork would not be ordered before the side stream's use)
        global _POOL
        self._prev_pool, _POOL = _POOL, _ZeroPool(self.dev, 1 << 18)
        return self

    def keepalive(self, *tensors):
        """Tensors allocated on the main stream that the side stream reads / writes: held until the join."""
        if self.enabled:
            _GRAD_PENDING["keep"].extend(t for t in tensors if t is not None)

    def __exit__(self, *exc):
        if not self.enabled:
            return False
        global _POOL
        _POOL = self._prev_pool
        self.ctx.__exit__(*exc)
        st = _GRAD_PENDING
        if self.side not in st["streams"]:
            st["streams"].append(self.side)
        st["keep"].extend(t for t in self.keep if t is not None)
        if not st["hooked"]:
            try:
                torch.autograd.Variable._execution_engine.queue_callback(join_grad_streams)
                st["hooked"] = True
            except RuntimeError:     # not inside a backward pass: the caller joins
                pass
        return False


def join_grad_streams():
    st = _GRAD_PENDING
    if st["streams"]:
        cur = torch.cuda.current_stream(st["streams"][0].device)
        for s_ in st["streams"]:
            cur.wait_stream(s_)
    st["streams"], st["keep"], st["hooked"] = [], [], False


# ------------------------------------------------------------------------------------------------
# Transformer backbone: fused multi-head attention, positional encoding, dropout-pack
# ------------------------------------------------------------------------------------------------
FLAG_NO_F32 = _lib.FLAG_NO_F32
MHA_DK = 128


def _seed_dev(t, p):
    return _ptr(step_counter(t.device)) if p > 0.0 else None


def mha_fwd(qkv: torch.Tensor, heads: int, kv_len: Optional[torch.Tensor], *, split_out: bool = False, save_stats: bool = False,
            dropout_p: float = 0.0, dropout_seed: int = 0):
    """qkv fp16 (B, T, 3*D) = [q | k | v], D = heads*128 -> ctx fp16 (B, T, D) (or (B, T, [hi D | lo D]) with split_out),
    and optionally (row_max, row_inv_l) fp32 (B, heads, T) for the backward pass."""
    assert qkv.dtype == torch.float16 and qkv.dim() == 3
    B, T, ld = qkv.shape
    D = heads * MHA_DK
    assert ld == 3 * D, (ld, D)
    ctx = torch.empty((B, T, 2 * D if split_out else D), device=qkv.device, dtype=torch.float16)
    rmax = rinv = None
    if save_stats:
        rmax = torch.empty((B, heads, T), device=qkv.device, dtype=torch.float32)
        rinv = torch.empty((B, heads, T), device=qkv.device, dtype=torch.float32)
    base = _ptr(qkv)
    _lib.check(_lib.load().osb_mha_fwd(base, base + 2 * D, base + 4 * D, ld, _ptr(kv_len), _ptr(ctx), ctx.shape[2], D if split_out else 0,
                                       _ptr(rmax), _ptr(rinv), B, T, heads, MHA_DK, MHA_DK ** -0.5, float(dropout_p), int(dropout_seed),
                                       _seed_dev(qkv, dropout_p), _stream()), "osb_mha_fwd")
    return ctx, rmax, rinv


def mha_bwd(qkv: torch.Tensor, heads: int, kv_len, ctx: torch.Tensor, d_ctx: torch.Tensor, rmax, rinv, *, dropout_p: float = 0.0,
            dropout_seed: int = 0) -> torch.Tensor:
    """-> dqkv fp16 (B, T, 3*D): gradient of the loss w.r.t. [q | k | v] given d_ctx fp16 (B, T, D)."""
    B, T, ld = qkv.shape
    D = heads * MHA_DK
    Tp = (T + 7) // 8 * 8
    dqkv = torch.empty((B, T, 3 * D), device=qkv.device, dtype=torch.float16)
    ds = torch.empty((B, T, heads * Tp), device=qkv.device, dtype=torch.float16)
    pd = torch.empty((B, T, heads * Tp), device=qkv.device, dtype=torch.float16)
    base = _ptr(qkv)
    lib = _lib.load()
    _lib.check(lib.osb_mha_bwd(base, base + 2 * D, base + 4 * D, ld, _ptr(kv_len), _ptr(ctx), ctx.shape[2], _ptr(d_ctx), d_ctx.shape[2],
                               _ptr(rmax), _ptr(rinv), _ptr(dqkv), 3 * D, _ptr(ds), _ptr(pd), heads * Tp, Tp, B, T, heads, MHA_DK,
                               MHA_DK ** -0.5, float(dropout_p), int(dropout_seed), _seed_dev(qkv, dropout_p), _stream()), "osb_mha_bwd")
    # dK_h = dS_h^T Q_h, dV_h = (P o D)_h^T dO_h : contractions over the query rows, both operands MN-major in place
    dkv = _zeros((2, heads, B, T, MHA_DK), qkv)
    for h in range(heads):
        _lib.check(lib.osb_gemm_wgrad_batched(_ptr(ds) + 2 * h * Tp, heads * Tp, base + 2 * h * MHA_DK, ld, _ptr(dkv[0, h]), B, T, T,
                                              MHA_DK, _stream()), "osb_gemm_wgrad_batched")
        _lib.check(lib.osb_gemm_wgrad_batched(_ptr(pd) + 2 * h * Tp, heads * Tp, _ptr(d_ctx) + 2 * h * MHA_DK, d_ctx.shape[2],
                                              _ptr(dkv[1, h]), B, T, T, MHA_DK, _stream()), "osb_gemm_wgrad_batched")
    for i in range(2):
        _lib.check(lib.osb_mha_pack_heads(_ptr(dkv[i]), _ptr(dqkv) + 2 * (i + 1) * D, 3 * D, B, T, heads, _stream()), "osb_mha_pack_heads")
    return dqkv


def dropout_pack_h16(x: torch.Tensor, dropout_p: float = 0.0, dropout_seed: int = 0) -> torch.Tensor:
    N = x.shape[-1]
    out = torch.empty(x.shape, device=x.device, dtype=torch.float16)
    _lib.check(_lib.load().osb_dropout_pack_h16(_ptr(_f32(x)), _ptr(out), x.numel() // N, N, float(dropout_p), int(dropout_seed),
                                                _seed_dev(x, dropout_p), _stream()), "osb_dropout_pack_h16")
    return out


def add_posenc(x: torch.Tensor, pe: Optional[torch.Tensor], alpha: Optional[torch.Tensor], dropout_p: float = 0.0, dropout_seed: int = 0):
    B, T, Cc = x.shape
    out = torch.empty_like(x)
    _lib.check(_lib.load().osb_add_posenc(_ptr(_f32(x)), _ptr(pe), _ptr(alpha), _ptr(out), B, T, Cc, float(dropout_p), int(dropout_seed),
                                          _seed_dev(x, dropout_p), _stream()), "osb_add_posenc")
    return out


_ALIGN_WS = {}


def fs2_losses(d_hat, p_hat, e_hat, ds, p_tgt, e_tgt, x_len):
    """-> (losses (3,) fp32 = [duration, pitch, energy], g_d, g_p, g_e (B,Tx) = gradients of the respective loss)."""
    B, Tx = d_hat.shape
    losses = torch.empty(3, device=d_hat.device, dtype=torch.float32)
    g = [torch.empty((B, Tx), device=d_hat.device, dtype=torch.float32) for _ in range(3)]
    _lib.check(_lib.load().osb_fs2_losses(_ptr(_f32(d_hat)), _ptr(_f32(p_hat)), _ptr(_f32(e_hat)), _ptr(_f32(ds)), _ptr(_f32(p_tgt)),
                                          _ptr(_f32(e_tgt)), _ptr(x_len), _ptr(losses), _ptr(g[0]), _ptr(g[1]), _ptr(g[2]), B, Tx,
                                          _stream()), "osb_fs2_losses")
    return losses, g[0], g[1], g[2]


def align_loss_fold(log_p_attn, path, m_len, per_sample_fs, fs_grad):
    """-> (3,) fp32 = [align_loss, forward-sum part, bin part]; adds the bin-loss gradient into `fs_grad` in place."""
    B, Tm, Tx = log_p_attn.shape
    out = torch.empty(3, device=log_p_attn.device, dtype=torch.float32)
    key = (log_p_attn.device.index or 0, B)
    ws = _ALIGN_WS.get(key)
    if ws is None:   # per-sample partial sums + a counter the kernel leaves at zero: allocated (and zeroed) once
        ws = torch.zeros(B + 1, device=log_p_attn.device, dtype=torch.float32)
        _ALIGN_WS[key] = ws
    _lib.check(_lib.load().osb_align_loss_fold(_ptr(_f32(log_p_attn)), _ptr(path), _ptr(m_len), _ptr(per_sample_fs), _ptr(fs_grad), _ptr(out),
                                               _ptr(ws), B, Tm, Tx, _stream()), "osb_align_loss_fold")
    return out


# ------------------------------------------------------------------------------------------------
# index / mask glue (osb_glue.cu)
# ------------------------------------------------------------------------------------------------
def sequence_masks(lengths: torch.Tensor, T: int, valid: bool = True, pad: bool = True):
    """-> (valid (B,T) bool | None, pad (B,T) bool | None) from one launch (utils/model.py:12-21)."""
    B = lengths.shape[0]
    lengths = lengths.to(torch.int64).contiguous()
    v = torch.empty((B, T), device=lengths.device, dtype=torch.bool) if valid else None
    p = torch.empty((B, T), device=lengths.device, dtype=torch.bool) if pad else None
    _lib.check(_lib.load().osb_sequence_mask(_ptr(lengths), _ptr(v), _ptr(p), B, T, _stream()), "osb_sequence_mask")
    return v, p


def segment_starts(rand: torch.Tensor, lengths: torch.Tensor, segment_size: int, margin: int = 4) -> torch.Tensor:
    """start = floor(rand * max(len - margin - S, 0)) as int64 (utils/segments.py:29-35)."""
    B = lengths.shape[0]
    out = torch.empty((B,), device=lengths.device, dtype=torch.int64)
    _lib.check(_lib.load().osb_segment_starts(_ptr(_f32(rand.contiguous())), _ptr(lengths.to(torch.int64).contiguous()), _ptr(out), B,
                                              int(margin), int(segment_size), _stream()), "osb_segment_starts")
    return out


def gather_segments(x: torch.Tensor, start: torch.Tensor, segment_size: int, scale: int = 1) -> torch.Tensor:
    """x (B,T,C) fp32 channels-last -> (B,S,C): rows start*scale .. +S of every sample, zero past T (utils/segments.py:38-60)."""
    B, T, Cc = x.shape
    out = torch.empty((B, segment_size, Cc), device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().osb_gather_segments(_ptr(_f32(x)), _ptr(start), _ptr(out), B, T, Cc, int(segment_size), int(scale), _stream()),
               "osb_gather_segments")
    return out


def crop_segments(wav: torch.Tensor, start_frames: torch.Tensor, n: int, hop: int) -> torch.Tensor:
    """wav (B,Tw) fp32 -> (B,n): samples start*hop .. +n, zero past the end (utils/segments.py:63-72)."""
    B, Tw = wav.shape
    out = torch.empty((B, n), device=wav.device, dtype=torch.float32)
    _lib.check(_lib.load().osb_gather_segments(_ptr(_f32(wav)), _ptr(start_frames.to(torch.int64).contiguous()), _ptr(out), B, Tw, 1, int(n),
                                               int(hop), _stream()), "osb_gather_segments")
    return out


# ------------------------------------------------------------------------------------------------
# period discriminators (osb_disc.cu): flat fp16 sequence matrices, see include/osb200.h
# ------------------------------------------------------------------------------------------------
def gemm_wgrad_strided(dy: torch.Tensor, a: torch.Tensor, dw: torch.Tensor, *, taps: int, pad: int, stride: int):
    """dw[tap,n,k] += sum_t dy[t,n] a[t*stride + tap - pad, k]; dy (T,N), a (T_in,K) fp16 flat matrices; dw fp32 (taps,N,K)."""
    assert dy.dtype == torch.float16 and a.dtype == torch.float16 and dw.dtype == torch.float32 and dy.dim() == 2 and a.dim() == 2
    T, N = dy.shape
    T_in, K = a.shape
    _lib.check(_lib.load().osb_gemm_wgrad_strided(_ptr(dy), N, _ptr(a), K, _ptr(dw), 1, T, T_in, N, K, taps, pad, stride, _stream()),
               "osb_gemm_wgrad_strided")
    return dw


def mpd_first_fwd(wav, w, bias, period: int, L1: int, P1: int, CP: int, stride: int, slope: float):
    NS, T = wav.shape
    out = torch.empty((NS * period * P1, CP), device=wav.device, dtype=torch.float16)
    _lib.check(_lib.load().osb_mpd_first_fwd(_ptr(_f32(wav)), _ptr(_f32(w)), _ptr(_f32(bias)), _ptr(out), NS, T, period, L1, P1, CP, stride,
                                             slope, _stream()), "osb_mpd_first_fwd")
    return out


def mpd_first_bwd(g, wav, w, period: int, L1: int, P1: int, stride: int, inv_scale: float, want_dwav: bool, want_dw: bool):
    NS, T = wav.shape
    dwav = torch.zeros_like(wav) if want_dwav else None
    dwb = torch.zeros((32 * 5 + 32,), device=wav.device, dtype=torch.float32) if want_dw else None
    _lib.check(_lib.load().osb_mpd_first_bwd(_ptr(g), _ptr(_f32(wav)), _ptr(_f32(w)), _ptr(dwav), _ptr(dwb),
                                             (dwb.data_ptr() + 4 * 160) if want_dw else None, NS, T, period, L1, P1, g.shape[1], stride,
                                             inv_scale, _stream()), "osb_mpd_first_bwd")
    return dwav, (dwb[:160].view(32, 5) if want_dw else None), (dwb[160:] if want_dw else None)


def mpd_post_fwd(x, w, bias, period: int, L: int, P: int):
    rows, Cc = x.shape
    NSEQ = rows // P
    out = torch.empty((NSEQ // period, L * period), device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().osb_mpd_post_fwd(_ptr(x), _ptr(_f32(w)), _ptr(_f32(bias)), _ptr(out), NSEQ, period, L, P, Cc, _stream()),
               "osb_mpd_post_fwd")
    return out


def mpd_post_bwd(dscore, x, w, period: int, L: int, P: int, scale: float, want_dx: bool, want_dw: bool):
    rows, Cc = x.shape
    NSEQ = rows // P
    dx = torch.empty_like(x) if want_dx else None
    dwb = torch.zeros((Cc * 3 + 4,), device=x.device, dtype=torch.float32) if want_dw else None
    _lib.check(_lib.load().osb_mpd_post_bwd(_ptr(_f32(dscore.contiguous())), _ptr(x), _ptr(_f32(w)), _ptr(dx), _ptr(dwb),
                                            (dwb.data_ptr() + 4 * Cc * 3) if want_dw else None, NSEQ, period, L, P, Cc, scale, _stream()),
               "osb_mpd_post_bwd")
    return dx, (dwb[:Cc * 3].view(Cc, 3) if want_dw else None), (dwb[Cc * 3:Cc * 3 + 1] if want_dw else None)


def lrelu_bwd_h16(dy, y, P: int, L: int, slope: float, colsum_scale: Optional[float] = None):
    """Gated gradient g; with `colsum_scale` also returns colsum_scale * column sums of g (the bias gradient), same launch."""
    rows, Cc = y.shape
    g = torch.empty_like(y)
    cs = torch.zeros((Cc,), device=y.device, dtype=torch.float32) if colsum_scale is not None else None
    _lib.check(_lib.load().osb_lrelu_bwd_h16(_ptr(dy), _ptr(y), _ptr(g), rows, Cc, P, L, slope, _ptr(cs), float(colsum_scale or 0.0), _stream()),
               "osb_lrelu_bwd_h16")
    return g if colsum_scale is None else (g, cs)


def col2im_h16(col, rows_in: int, Cc: int, taps: int, pad: int, stride: int, reversed_taps: bool):
    rows_out = col.shape[0]
    dx = torch.empty((rows_in, Cc), device=col.device, dtype=torch.float16)
    _lib.check(_lib.load().osb_col2im_h16(_ptr(col), _ptr(dx), rows_in, rows_out, Cc, taps, pad, stride, int(reversed_taps), _stream()),
               "osb_col2im_h16")
    return dx


def l1_pair_fwd(a, b):
    out = torch.zeros((1,), device=a.device, dtype=torch.float32)
    _lib.check(_lib.load().osb_l1_pair_fwd(_ptr(a), _ptr(b), _ptr(out), a.numel(), _stream()), "osb_l1_pair_fwd")
    return out


def l1_pair_bwd(a, b, coef, scale: float):
    db = torch.empty_like(b)
    _lib.check(_lib.load().osb_l1_pair_bwd(_ptr(a), _ptr(b), _ptr(_f32(coef)), scale, _ptr(db), a.numel(), _stream()), "osb_l1_pair_bwd")
    return db


# ------------------------------------------------------------------------------------------------
# resolution discriminators (osb_disc.cu)
# ------------------------------------------------------------------------------------------------
def mrd_first_fwd(spec, w, bias, H1: int, W1: int, P1: int, slope: float):
    NS, Fq, W = spec.shape
    out = torch.empty((NS * W1 * P1, 64), device=spec.device, dtype=torch.float16)
    _lib.check(_lib.load().osb_mrd_first_fwd(_ptr(_f32(spec)), _ptr(_f32(w)), _ptr(_f32(bias)), _ptr(out), NS, Fq, W, H1, W1, P1, slope, _stream()),
               "osb_mrd_first_fwd")
    return out


def mrd_first_bwd(g, spec, w, H1: int, W1: int, P1: int, inv_scale: float, want_dspec: bool, want_dw: bool):
    NS, Fq, W = spec.shape
    dspec = torch.empty_like(spec) if want_dspec else None
    dwb = torch.zeros((64 * 35 + 64,), device=spec.device, dtype=torch.float32) if want_dw else None
    _lib.check(_lib.load().osb_mrd_first_bwd(_ptr(g), _ptr(_f32(spec)), _ptr(_f32(w)), _ptr(dspec), _ptr(dwb),
                                             (dwb.data_ptr() + 4 * 64 * 35) if want_dw else None, NS, Fq, W, H1, W1, P1, inv_scale, _stream()),
               "osb_mrd_first_bwd")
    return dspec, (dwb[:64 * 35].view(64, 35) if want_dw else None), (dwb[64 * 35:] if want_dw else None)


def wim2col_h16(x, NS: int, W_in: int, W_out: int, P: int, KW: int, pw: int, sw: int):
    Cc = x.shape[1]
    xcol = torch.empty((NS * W_out * P, KW * Cc), device=x.device, dtype=torch.float16)
    _lib.check(_lib.load().osb_wim2col_h16(_ptr(x), _ptr(xcol), NS, W_in, W_out, P, Cc, KW, pw, sw, _stream()), "osb_wim2col_h16")
    return xcol


def wcol2im_h16(dxcol, NS: int, W_in: int, W_out: int, P: int, Cc: int, KW: int, pw: int, sw: int):
    dx = torch.empty((NS * W_in * P, Cc), device=dxcol.device, dtype=torch.float16)
    _lib.check(_lib.load().osb_wcol2im_h16(_ptr(dxcol), _ptr(dx), NS, W_in, W_out, P, Cc, KW, pw, sw, _stream()), "osb_wcol2im_h16")
    return dx


def mrd_post_fwd(x, w, bias, NS: int, W: int, H: int, P: int):
    out = torch.empty((NS, H * W), device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().osb_mrd_post_fwd(_ptr(x), _ptr(_f32(w)), _ptr(_f32(bias)), _ptr(out), NS, W, H, P, _stream()), "osb_mrd_post_fwd")
    return out


def mrd_post_bwd(dscore, x, w, NS: int, W: int, H: int, P: int, scale: float, want_dx: bool, want_dw: bool):
    dx = torch.empty_like(x) if want_dx else None
    dwb = torch.zeros((64 * 9 + 4,), device=x.device, dtype=torch.float32) if want_dw else None
    _lib.check(_lib.load().osb_mrd_post_bwd(_ptr(_f32(dscore.contiguous())), _ptr(x), _ptr(_f32(w)), _ptr(dx), _ptr(dwb),
                                            (dwb.data_ptr() + 4 * 64 * 9) if want_dw else None, NS, W, H, P, scale, _stream()),
               "osb_mrd_post_bwd")
    return dx, (dwb[:64 * 9].view(64, 9) if want_dw else None), (dwb[64 * 9:64 * 9 + 1] if want_dw else None)


def spec_im2col_h16(spec, H1: int, W1: int, P1: int):
    NS, Fq, W = spec.shape
    xcol = torch.empty((NS * W1 * P1, 64), device=spec.device, dtype=torch.float16)
    _lib.check(_lib.load().osb_spec_im2col_h16(_ptr(_f32(spec)), _ptr(xcol), NS, Fq, W, H1, W1, P1, _stream()), "osb_spec_im2col_h16")
    return xcol


def spec_col2im(col, shape, H1: int, W1: int, P1: int, inv_scale: float):
    NS, Fq, W = shape
    dspec = torch.empty((NS, Fq, W), device=col.device, dtype=torch.float32)
    _lib.check(_lib.load().osb_spec_col2im(_ptr(col), _ptr(dspec), NS, Fq, W, H1, W1, P1, inv_scale, _stream()), "osb_spec_col2im")
    return dspec
