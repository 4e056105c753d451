"""torch.autograd glue for the training path: each Function's forward AND backward run libosb200 kernels
(tcgen05 dgrad / wgrad contractions plus the HBM-bound kernels of osb_backward.cu); autograd itself is only
the tape.  Activations cross Function boundaries in fp32; inside a Function the tensor-core operands are fp16
(the reference's own GPU default is `16-mixed`), and gradients carry the caller's static loss scale.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch.autograd import Function

from . import ops


# --------------------------------------------------------------------------------------------------
# weight packing helpers (fp32 master -> fp16 operands).  Only forward packs exist: data gradients read the same pack as an
# MN-major B operand (ops.gemm(..., w_mn=True[, tap_reverse=True])).
# --------------------------------------------------------------------------------------------------
def pack_nk(w: torch.Tensor, k_pad: Optional[int] = None) -> torch.Tensor:
    """(N, K) fp32 -> (1, N, Kp) fp16."""
    N, K = w.shape
    Kp = k_pad or K
    return ops.pack_h16(w.contiguous(), rows=N, cols=K, src_ld=K, dst_cols=Kp).view(1, N, Kp)


def pack_params_nk(ws, col_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Parameters (N_i, K) fp32 (stacked along N), optionally scaled per column -> (1, sum N_i, K) fp16, served by the step
    packer (model/packing.py: one launch for all the packs of a training step once the step has been seen)."""
    from .model.packing import current_packer

    ws = [w for w in ws]
    K = ws[0].shape[1]
    n_total = sum(w.shape[0] for w in ws)

    def direct():
        out = torch.empty((1, n_total, K), device=ws[0].device, dtype=torch.float16)
        r = 0
        for w in ws:
            ops.pack_h16(w.detach().contiguous(), rows=w.shape[0], cols=K, src_ld=K, dst_cols=K, col_scale=col_scale, out=out[0, r: r + w.shape[0]])
            r += w.shape[0]
        return out

    packer = current_packer()
    if packer is None or any(not w.is_contiguous() for w in ws):
        return direct()
    return packer.request("nk", ws, col_scale, None, direct)


def folded_bias(b1: torch.Tensor, w1: torch.Tensor, ln_b: torch.Tensor) -> torch.Tensor:
    """b1 + W1 @ ln_b (the LayerNorm bias folded into the bias of the Linear that follows it), through the step packer: inside
    a recorded step it costs nothing extra (a job of the one osb_pack_multi launch) instead of a copy + gemv per block."""
    from .model.packing import current_packer

    def direct():
        return torch.addmv(b1, w1, ln_b)

    packer = current_packer()
    if packer is None or not (w1.is_contiguous() and w1.shape[1] % 4 == 0):
        return direct()
    return packer.request("matvec", [w1], ln_b, None, direct, aux=b1)


def pack_param_conv(w: torch.Tensor, k_pad: Optional[int] = None) -> torch.Tensor:
    """Conv1d parameter (N, Cin, k) -> (k, N, Cin_pad) fp16 through the step packer."""
    from .model.packing import current_packer

    packer = current_packer()
    if packer is None or not w.is_contiguous():
        return ops.pack_conv_h16(w, k_pad=k_pad)
    return packer.request("conv", [w], None, k_pad, lambda: ops.pack_conv_h16(w, k_pad=k_pad))


def _conv_wgrad(g_h16: torch.Tensor, a_h16: torch.Tensor, N: int, Cin: int, k: int, pad: int) -> torch.Tensor:
    """-> Conv1d weight gradient in the parameter's (N, Cin, k) layout."""
    dw = ops.zeros((k, N, Cin), g_h16)
    ops.gemm_wgrad(g_h16, a_h16, dw, taps=k, pad=pad, K=Cin)
    return dw.permute(1, 2, 0).contiguous()


# --------------------------------------------------------------------------------------------------
# text embedding
# --------------------------------------------------------------------------------------------------
class EmbedTextFn(Function):
    @staticmethod
    def forward(ctx, ids, table, scale, inv_freq, padding_idx):
        ctx.save_for_backward(ids, inv_freq)
        ctx.n_vocab, ctx.padding_idx = table.shape[0], padding_idx
        return ops.embed_text(ids, table, inv_freq, scale)

    @staticmethod
    @ops.pooled
    def backward(ctx, dout):
        ids, inv_freq = ctx.saved_tensors
        dtable, dscale = ops.embed_text_bwd(dout.contiguous(), ids, inv_freq, ctx.n_vocab, ctx.padding_idx)
        return None, dtable, dscale, None, None


# --------------------------------------------------------------------------------------------------
# LayerNorm with affine (final_layer_norm of a backbone)
# --------------------------------------------------------------------------------------------------
class LayerNormFn(Function):
    @staticmethod
    def forward(ctx, x, w, b, eps):
        ctx.save_for_backward(x, w)
        ctx.eps = eps
        out, _ = ops.layernorm(x, w, b, eps, f32=True, h16=False)
        return out

    @staticmethod
    @ops.pooled
    def backward(ctx, dout):
        x, w = ctx.saved_tensors
        dx, dw, db = ops.layernorm_bwd(dout.contiguous(), x, w, ctx.eps)
        return dx, dw, db, None


# --------------------------------------------------------------------------------------------------
# ConvNeXt block
# --------------------------------------------------------------------------------------------------
class ConvNeXtBlockFn(Function):
    """out = (x + gamma * pw2(gelu(pw1(LN(dwconv7(x))))) * row_scale[b]) * keep      (channels-last)

    Supported shapes ((C, I) = (256, 1024), (384, 1152)) run ONE fused tcgen05 kernel forward (osb_convnext_block_fwd_train:
    dwconv + LN prologue, both pointwise convolutions, GELU, layer scale, DropPath, residual, mask; it also emits xhat / rstd /
    pre-GELU / GELU activations for the backward) and, backward, one fused kernel for both data-gradient contractions
    (osb_convnext_block_bwd) followed by one LayerNorm + depthwise-conv backward pass (osb_ln_dwconv_bwd).  The weight / bias /
    layer-scale gradients only meet the optimizer: they run on a side stream (ops.grad_side)."""

    FUSED = True   # tests flip this to compare the fused path with the three-kernel path below

    @staticmethod
    def forward(ctx, x, dw_w, dw_b, ln_w, ln_b, w1, b1, w2, b2, gamma, pad_mask, row_scale, eps):
        B, T, C = x.shape
        I = w1.shape[0]
        b1f = folded_bias(b1, w1, ln_b)      # fold the LN affine into pwconv1: bias here, weight as a column scale of the pack
        w1f_h, w2_h = pack_params_nk([w1], col_scale=ln_w), pack_params_nk([w2])
        ctx.fused = ConvNeXtBlockFn.FUSED and (C, I) in ops.FUSED_BLOCK_SHAPES
        if ctx.fused:
            out, xhat, rstd, pre, h = ops.convnext_block_fwd_train(x, dw_w.view(C, 7), dw_b, w1f_h[0], b1f, w2_h[0], b2, gamma, row_scale,
                                                                   pad_mask, eps)
            ctx.save_for_backward(x, dw_w, ln_w, ln_b, w1, w2, gamma, xhat, rstd, pre, h, out, pad_mask, row_scale, w1f_h, w2_h)
            return out
        xhat, rstd = ops.dwconv_ln(x, dw_w.view(C, 7), dw_b, eps, want_rstd=True)
        h, pre, _ = ops.gemm(xhat, w1f_h, epi=ops.EPI_GELU, flags=ops.FLAG_SAVE_PRE, bias=b1f)
        flags = ops.FLAG_SAVE_PRE | (ops.FLAG_KEEPMASK if pad_mask is not None else 0)
        out, z, _ = ops.gemm(h, w2_h, epi=ops.EPI_RESID, flags=flags, bias=b2, resid=x, gamma=gamma, row_scale=row_scale,
                             pad_mask=pad_mask)
        ctx.save_for_backward(x, dw_w, ln_w, ln_b, w1, w2, gamma, xhat, rstd, pre, h, z, pad_mask, row_scale, w1f_h, w2_h)
        return out

    @staticmethod
    @ops.pooled
    def backward(ctx, dout):
        x, dw_w, ln_w, ln_b, w1, w2, gamma, xhat, rstd, pre, h, z_or_out, pad_mask, row_scale, w1f_h, w2_h = ctx.saved_tensors
        B, T, C = x.shape
        I = w1.shape[0]
        dout = dout.contiguous()
        if ctx.fused:
            out = z_or_out
            dyg, dh, dxh = ops.convnext_block_bwd(dout, gamma, row_scale, pad_mask, pre, w2_h[0], w1f_h[0])
            dx, ddw, ddb = ops.ln_dwconv_bwd(dxh, xhat, rstd, dout, x, dw_w.view(C, 7), pad_mask)
            # parameter gradients: off the critical path, as two independent chains on two side streams (the five launches in a
            # row were ~100 us behind the LAST block of the backward pass, i.e. the tail of the step)
            with ops.grad_side(x, dout, dyg, h, out, x):
                dgamma, db2 = ops.resid_param_grad(dout, out, x, gamma, pad_mask, row_scale)
                dw2 = ops.zeros((1, C, I), x)
                ops.gemm_wgrad(dyg, h, dw2)
            with ops.grad_side(x, dh, xhat):
                dw1f = ops.zeros((1, I, C), x)
                ops.gemm_wgrad(dh, xhat, dw1f)
                db1 = ops.colsum_h16(dh)
                dw1, dln_w, dln_b = ops.ln_fold_bwd(dw1f.view(I, C), w1, ln_w, ln_b, db1)
            return dx, ddw.reshape(C, 1, 7), ddb, dln_w, dln_b, dw1, db1, dw2.view(C, I), db2, dgamma, None, None, None
        z = z_or_out
        dyg, dgamma, db2 = ops.resid_bwd_prep(dout, z, gamma, pad_mask, row_scale, T)
        # pwconv2: dgrad (with the GELU derivative fused) and wgrad
        db1 = ops.zeros((I,), x)                                           # bias gradient: column sums taken in the dgrad epilogue
        # data gradients run on the FORWARD weight packs (MN-major B operand): no transposed copies are made
        dpre, _, _ = ops.gemm(dyg, w2_h, epi=ops.EPI_GELU_BWD, aux_in=pre, colsum=db1, w_mn=True)
        dw2 = ops.zeros((1, C, I), x)
        ops.gemm_wgrad(dyg, h, dw2)
        # pwconv1: dgrad (LayerNorm backward fused: the CTA owns whole rows) and wgrad
        dd, _, _ = ops.gemm(dpre, w1f_h, epi=ops.EPI_LN_BWD, aux_in=xhat, row_stat=rstd, w_mn=True)
        dw1f = ops.zeros((1, I, C), x)
        ops.gemm_wgrad(dpre, xhat, dw1f)
        dw1, dln_w, dln_b = ops.ln_fold_bwd(dw1f.view(I, C), w1, ln_w, ln_b, db1)
        dx, ddw, ddb = ops.dwconv_bwd(dd, dout, x, dw_w.view(C, 7), pad_mask)
        return dx, ddw.view(C, 1, 7), ddb, dln_w, dln_b, dw1, db1, dw2.view(C, I), db2, dgamma, None, None, None


# --------------------------------------------------------------------------------------------------
# VariancePredictor: [Conv1d(k) + ReLU + LayerNorm] * L + Linear(->1) + masked_fill
# --------------------------------------------------------------------------------------------------
class VariancePredictorFn(Function):
    @staticmethod
    def forward(ctx, x, pad_mask, kernel_size, eps, dropout_p, dropout_seed, lin_w, lin_b, *layer_params):
        """layer_params = (conv_w, conv_b, ln_w, ln_b) per layer.  x fp32 (B,T,C) -> (B,T) fp32.
        dropout_p > 0 applies dropout after every LayerNorm (core.py:78); layer l uses seed dropout_seed + l."""
        L = len(layer_params) // 4
        pad = (kernel_size - 1) // 2
        a = ops.to_h16(x.contiguous())
        acts, pres, wps = [a], [], []
        out = None
        for l in range(L):
            cw, cb, lw, lb = layer_params[4 * l: 4 * l + 4]
            wp = pack_param_conv(cw)
            wps.append(wp)
            if l < L - 1:
                y, r, _ = ops.gemm(acts[-1], wp, epi=ops.EPI_RELU_LN, flags=ops.FLAG_SAVE_PRE, pad=pad, bias=cb, ln_w=lw, ln_b=lb, ln_eps=eps,
                                   dropout_p=dropout_p, dropout_seed=dropout_seed + l)
                acts.append(y)
            else:
                _, r, out = ops.gemm(acts[-1], wp, epi=ops.EPI_RELU_LN, flags=ops.FLAG_SAVE_PRE | ops.FLAG_DOT, pad=pad, bias=cb, ln_w=lw,
                                     ln_b=lb, ln_eps=eps, dot_w=lin_w.view(-1), dot_b=lin_b, pad_mask=pad_mask, dropout_p=dropout_p,
                                     dropout_seed=dropout_seed + l)
            pres.append(r)
        ctx.save_for_backward(pad_mask, lin_w, *layer_params, *acts, *pres, *wps)
        ctx.L, ctx.k, ctx.eps = L, kernel_size, eps
        ctx.drop_p, ctx.drop_seed = dropout_p, dropout_seed
        ctx.x_needs_grad = x.requires_grad
        return out

    @staticmethod
    @ops.pooled
    def backward(ctx, d_out):
        L, k, eps = ctx.L, ctx.k, ctx.eps
        saved = ctx.saved_tensors
        pad_mask, lin_w = saved[0], saved[1]
        layer_params = saved[2: 2 + 4 * L]
        acts = saved[2 + 4 * L: 2 + 5 * L]
        pres = saved[2 + 5 * L: 2 + 6 * L]
        wps = saved[2 + 6 * L: 2 + 7 * L]
        pad = (k - 1) // 2
        grads: List[Optional[torch.Tensor]] = [None] * (4 * L)
        cw, cb, lw, lb = layer_params[4 * (L - 1): 4 * L]
        g, dlin_w, dlin_b, dln_w, dln_b = ops.predictor_tail_bwd(d_out.contiguous(), pad_mask, pres[L - 1], lw, lb, lin_w.view(-1), eps,
                                                                 ctx.drop_p, ctx.drop_seed + L - 1)
        grads[4 * (L - 1) + 2], grads[4 * (L - 1) + 3] = dln_w, dln_b
        dx = None
        with ops.grad_side(g, g):   # parameter gradients leave the critical path (the data-gradient chain g -> g_prev)
            grads[4 * (L - 1) + 1] = ops.colsum_h16(g)
        for l in range(L - 1, -1, -1):
            cw = layer_params[4 * l]
            N, Cin, _ = cw.shape
            with ops.grad_side(g, g, acts[l]):
                grads[4 * l] = _conv_wgrad(g, acts[l], N, Cin, k, pad)
            if l > 0:
                lw_prev = layer_params[4 * (l - 1) + 2]
                grads[4 * (l - 1) + 1] = ops.zeros((Cin,), g)             # conv bias gradient of layer l-1, summed in the epilogue
                g_prev, gy, _ = ops.gemm(g, wps[l], epi=ops.EPI_RELU_LN_BWD, flags=ops.FLAG_OUT_H16, pad=k - 1 - pad,
                                         aux_in=pres[l - 1], ln_w=lw_prev, ln_eps=eps, dropout_p=ctx.drop_p,
                                         dropout_seed=ctx.drop_seed + l - 1, colsum=grads[4 * (l - 1) + 1], w_mn=True, tap_reverse=True,
                                         N=Cin)
                with ops.grad_side(gy, gy, pres[l - 1]):
                    grads[4 * (l - 1) + 2], grads[4 * (l - 1) + 3] = ops.ln_param_grad(gy, pres[l - 1], lw_prev, eps)
                g = g_prev
            elif ctx.x_needs_grad:
                dx, _, _ = ops.gemm(g, wps[0], epi=ops.EPI_BIAS, pad=k - 1 - pad, w_mn=True, tap_reverse=True, N=Cin)
        return (dx, None, None, None, None, None, dlin_w.view(1, -1), dlin_b, *grads)


# --------------------------------------------------------------------------------------------------
# stack of Conv1d layers with ReLU between them (AlignmentModule text / feature encoders)
# --------------------------------------------------------------------------------------------------
class ConvStackFn(Function):
    @staticmethod
    def forward(ctx, x, k_pad, *params):
        """params = (conv_w (N,Cin,k), conv_b) per layer; ReLU after every layer but the last.
        x fp32 (B,T,Cin) -> fp32 (B,T,N_last).  k_pad: channel padding of the first operand (e.g. 100 mel bins -> 128)."""
        L = len(params) // 2
        a = ops.to_h16(x.contiguous(), pad_to=k_pad or None)
        acts, wps = [a], []
        out = None
        for l in range(L):
            cw, cb = params[2 * l], params[2 * l + 1]
            k = cw.shape[2]
            wp = pack_param_conv(cw, k_pad=acts[-1].shape[-1])
            wps.append(wp)
            if l < L - 1:
                y, _, _ = ops.gemm(acts[-1], wp, epi=ops.EPI_RELU, pad=(k - 1) // 2, bias=cb)
                acts.append(y)
            else:
                out, _, _ = ops.gemm(acts[-1], wp, epi=ops.EPI_BIAS, pad=(k - 1) // 2, bias=cb)
        ctx.save_for_backward(*params, *acts, *wps)
        ctx.L = L
        ctx.x_needs_grad = x.requires_grad
        ctx.cin0 = x.shape[-1]
        return out

    @staticmethod
    @ops.pooled
    def backward(ctx, dout):
        L = ctx.L
        saved = ctx.saved_tensors
        params, acts, wps = saved[: 2 * L], saved[2 * L: 3 * L], saved[3 * L:]
        grads: List[Optional[torch.Tensor]] = [None] * (2 * L)
        g = ops.to_h16(dout.contiguous())
        dx = None
        with ops.grad_side(g, g):
            grads[2 * (L - 1) + 1] = ops.colsum_h16(g)
        for l in range(L - 1, -1, -1):
            cw = params[2 * l]
            N, Cin, k = cw.shape
            pad = (k - 1) // 2
            with ops.grad_side(g, g, acts[l]):
                grads[2 * l] = _conv_wgrad(g, acts[l], N, Cin, k, pad)
            if l > 0:
                grads[2 * (l - 1) + 1] = ops.zeros((Cin,), g)             # bias gradient of layer l-1, summed in the epilogue
                g, _, _ = ops.gemm(g, wps[l], epi=ops.EPI_RELU_BWD, pad=k - 1 - pad, aux_in=acts[l],
                                   colsum=grads[2 * (l - 1) + 1], w_mn=True, tap_reverse=True, N=Cin)
            elif ctx.x_needs_grad:
                dx, _, _ = ops.gemm(g, wps[0], epi=ops.EPI_BIAS, pad=k - 1 - pad, w_mn=True, tap_reverse=True, N=Cin)
        return (dx, None, *grads)


# --------------------------------------------------------------------------------------------------
# variance embedding: out = (x + bias + Conv1d(1->C,k)(val)) * keep
# --------------------------------------------------------------------------------------------------
class VarianceEmbedFn(Function):
    @staticmethod
    def forward(ctx, x, val, w, b, pad_mask, emb_scale=None):
        C = x.shape[-1]
        ctx.save_for_backward(val, pad_mask, emb_scale)
        ctx.k = w.shape[-1]
        ctx.x_needs_grad = x.requires_grad
        out, _ = ops.variance_embed(x.contiguous(), val.contiguous(), w.view(C, -1), b, pad_mask, f32=True, h16=False,
                                    emb_scale=emb_scale)
        return out

    @staticmethod
    @ops.pooled
    def backward(ctx, dout):
        val, pad_mask, emb_scale = ctx.saved_tensors
        dx, dw, db = ops.variance_embed_bwd(dout.contiguous(), val, pad_mask, ctx.k, want_dx=ctx.x_needs_grad, emb_scale=emb_scale)
        return dx, None, dw.view(dw.shape[0], 1, ctx.k), db, None, None


# --------------------------------------------------------------------------------------------------
# WaveNeXt head: clip(Linear2(Linear1(x))) with the two Linears folded into one GEMM
# --------------------------------------------------------------------------------------------------
class WaveNeXtHeadFn(Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2):
        """x fp32 (B,T,dim); linear_1 (l_fft, dim) + bias, linear_2 (hop, l_fft) -> audio (B, T*hop) in [-1, 1]."""
        wc = w2 @ w1                      # (hop, dim): no non-linearity between the two Linears (wavenext/__init__.py:43-45)
        bc = w2 @ b1
        a = ops.to_h16(x.contiguous())
        wc_h = pack_nk(wc)
        out, _, _ = ops.gemm(a, wc_h, epi=ops.EPI_BIAS, flags=ops.FLAG_CLIP, bias=bc.contiguous())
        ctx.save_for_backward(a, out, w1, b1, w2, wc_h)
        return out.view(out.shape[0], -1)

    @staticmethod
    @ops.pooled
    def backward(ctx, dout):
        a, out, w1, b1, w2, wc_h = ctx.saved_tensors
        B, T, hop = out.shape
        # clip passes gradient only strictly inside (-1, 1)  (torch.clip backward)
        g32 = (dout.reshape(B, T, hop) * ((out > -1.0) & (out < 1.0))).contiguous()
        g = ops.to_h16(g32)
        dx, _, _ = ops.gemm(g, wc_h, epi=ops.EPI_BIAS, w_mn=True)
        dwc = ops.zeros((1, hop, a.shape[-1]), a)
        ops.gemm_wgrad(g, a, dwc)
        dwc = dwc[0]
        dbc = ops.colsum_h16(g)
        # un-fold: wc = w2 @ w1, bc = w2 @ b1   (small fp32 matrix products on the weights)
        dw2 = dwc @ w1.t() + torch.outer(dbc, b1)
        dw1 = w2.t() @ dwc
        db1 = w2.t() @ dbc
        return dx, dw1, db1, dw2


# --------------------------------------------------------------------------------------------------
# attention log-probabilities of the alignment module
# --------------------------------------------------------------------------------------------------
class AttnLogProbFn(Function):
    """log_p_attn[b,t,n] = log_softmax_n(-||f_t - e_n||_2 over n < x_len[b]) + prior[b,t,n]   (alignments.py:66-81).

    ||f - e||^2 = |f|^2 + |e|^2 - 2 f.e with the dot product on the tensor cores in split precision (fp16 hi+lo,
    three passes, ~2^-21 relative): the (B,Tm,Tx,C) difference tensor of the reference is never formed, and the
    masked log-softmax + prior run in the GEMM epilogue (one thread owns one attention row)."""

    @staticmethod
    def forward(ctx, fe, te, prior, x_len, m_len):
        fe, te = fe.contiguous(), te.contiguous()
        B, Tm, C = fe.shape
        Tx = te.shape[1]
        f16 = ops.to_h16(fe, split=True)
        e16 = ops.to_h16(te, split=True)
        nf, ne = ops.rownorm_sq(fe), ops.rownorm_sq(te)
        lp, _, lse = ops.gemm(f16, e16, epi=ops.EPI_ATTN_LOGP, flags=ops.FLAG_SPLIT_IN, K=C, N=Tx, w_batched=True, row_stat=nf, bias=ne,
                              resid=prior, col_len=x_len)
        ctx.save_for_backward(fe, te, lp, lse, prior, x_len, m_len, f16)
        return lp

    @staticmethod
    @ops.pooled
    def backward(ctx, G):
        fe, te, lp, lse, prior, x_len, m_len, f16 = ctx.saved_tensors
        B, Tm, C = fe.shape
        Tx = te.shape[1]
        wn, neg_rsn, csn = ops.attn_bwd_prep(G.contiguous(), lp, prior, lse, x_len, m_len)
        # dF = -rowsum(Wn) * F + Wn @ E      (contraction over Tx; E transposed per batch is the K-major B operand)
        eT = ops.transpose_pack_h16(te)                                   # (B, C, Tx8)
        dF, _, _ = ops.gemm(wn, eT, epi=ops.EPI_AXPY, K=Tx, N=C, w_batched=True, resid=fe, row_stat=neg_rsn)
        # dE = -colsum(Wn) * E + Wn^T @ F    (contraction over Tm: both operands MN-major, per batch)
        dE = ops.scale_rows(te, csn, sign=-1.0)
        f_hi = f16.view(B, Tm, 2 * C)                                      # [hi | lo] rows; the hi halves are the operand
        ops.gemm_wgrad_batched(wn, f_hi, dE, N=Tx, K=C)
        return dF, dE, None, None, None


# --------------------------------------------------------------------------------------------------
# forward-sum (CTC) alignment loss
# --------------------------------------------------------------------------------------------------
class ForwardSumLossFn(Function):
    @staticmethod
    def forward(ctx, log_p_attn, x_len, m_len, blank_logit):
        per_sample, grad = ops.forward_sum(log_p_attn.contiguous(), x_len, m_len, blank_logit)
        ctx.save_for_backward(grad)
        return per_sample.sum() / log_p_attn.shape[0]

    @staticmethod
    @ops.pooled
    def backward(ctx, dloss):
        (grad,) = ctx.saved_tensors
        return grad * dloss, None, None, None


# --------------------------------------------------------------------------------------------------
# Transformer backbone (modules/transformer.py): scaled positional encoding and one pre-LN encoder layer
# --------------------------------------------------------------------------------------------------
class PosEncFn(Function):
    """out = (x + alpha * pe[:T]) * D   (ScaledPositionalEncoding.forward, _transformer/embedding.py:111-124)."""

    @staticmethod
    def forward(ctx, x, alpha, pe, p, seed):
        ctx.save_for_backward(pe)
        ctx.p, ctx.seed, ctx.T = p, seed, x.shape[1]
        return ops.add_posenc(x.contiguous(), pe, alpha.detach().reshape(1), p, seed)

    @staticmethod
    def backward(ctx, dout):
        (pe,) = ctx.saved_tensors
        dout = dout.contiguous()
        dx = ops.add_posenc(dout, None, None, ctx.p, ctx.seed) if ctx.p > 0.0 else dout
        dalpha = (dx * pe[: ctx.T]).sum()  # one scalar: a reduction of the (B,T,C) gradient against the table
        return dx, dalpha, None, None, None


class TransformerLayerFn(Function):
    """x1 = x + D1(linear_out(MHA(LN1 x)));  out = x1 + D3(w_2(D2(relu(w_1(LN2 x1)))))     (encoder_layer.py:88-116)

    Seeds: attention dropout seed+0, D1 seed+1, D2 seed+2, D3 seed+3 (counter-based masks are regenerated in backward)."""

    @staticmethod
    def forward(ctx, x, kv_len, heads, p_drop, p_attn, p_ffn, seed, ones, n1w, n1b, wq, bq, wk, bk, wv, bv, wo, bo, n2w, n2b, w1, b1, w2, b2):
        eps = 1e-12
        x = x.contiguous()
        _, xn = ops.layernorm(x, n1w, n1b, eps, f32=False, h16=True)
        bqkv = torch.cat([bq, bk, bv])
        wqkv_h, wo_h = pack_params_nk([wq, wk, wv]), pack_params_nk([wo])
        w1_h, w2_h = pack_params_nk([w1.view(w1.shape[0], -1)]), pack_params_nk([w2.view(w2.shape[0], -1)])
        _, qkv, _ = ops.gemm(xn, wqkv_h, epi=ops.EPI_BIAS, flags=ops.FLAG_OUT_H16 | ops.FLAG_NO_F32, bias=bqkv)
        att, rmax, rinv = ops.mha_fwd(qkv, heads, kv_len, save_stats=True, dropout_p=p_attn, dropout_seed=seed)
        x1, _, _ = ops.gemm(att, wo_h, epi=ops.EPI_RESID, bias=bo, resid=x, gamma=ones, dropout_p=p_drop, dropout_seed=seed + 1)
        _, xn2 = ops.layernorm(x1, n2w, n2b, eps, f32=False, h16=True)
        h, _, _ = ops.gemm(xn2, w1_h, epi=ops.EPI_RELU, bias=b1, dropout_p=p_ffn, dropout_seed=seed + 2)
        out, _, _ = ops.gemm(h, w2_h, epi=ops.EPI_RESID, bias=b2, resid=x1, gamma=ones, dropout_p=p_drop, dropout_seed=seed + 3)
        ctx.save_for_backward(x, kv_len, n1w, wqkv_h, wo_h, n2w, w1_h, w2_h, xn, qkv, att, rmax, rinv, x1, xn2, h)
        ctx.cfg = (heads, p_drop, p_attn, p_ffn, seed, eps)
        return out

    @staticmethod
    @ops.pooled
    def backward(ctx, dout):
        x, kv_len, n1w, wqkv_h, wo_h, n2w, w1_h, w2_h, xn, qkv, att, rmax, rinv, x1, xn2, h = ctx.saved_tensors
        heads, p_drop, p_attn, p_ffn, seed, eps = ctx.cfg
        B, T, D = x.shape
        U = w1_h.shape[1]
        dout = dout.contiguous()
        # ---- feed-forward branch ----
        g2 = ops.dropout_pack_h16(dout, p_drop, seed + 3)                      # grad wrt (w_2 h + b2)
        dw2 = ops.zeros((1, D, U), x)
        ops.gemm_wgrad(g2, h, dw2)
        db2 = ops.colsum_h16(g2)
        db1 = ops.zeros((U,), x)
        dh, _, _ = ops.gemm(g2, w2_h, epi=ops.EPI_RELU_BWD, aux_in=h, dropout_p=p_ffn, dropout_seed=seed + 2, colsum=db1, w_mn=True)
        dw1 = ops.zeros((1, U, D), x)
        ops.gemm_wgrad(dh, xn2, dw1)
        dxn2, _, _ = ops.gemm(dh, w1_h, epi=ops.EPI_BIAS, w_mn=True)
        dln2, dn2w, dn2b = ops.layernorm_bwd(dxn2, x1, n2w, eps)
        dx1 = dout + dln2
        # ---- attention branch ----
        g1 = ops.dropout_pack_h16(dx1, p_drop, seed + 1)                       # grad wrt (linear_out att + bo)
        dwo = ops.zeros((1, D, D), x)
        ops.gemm_wgrad(g1, att, dwo)
        dbo = ops.colsum_h16(g1)
        _, datt, _ = ops.gemm(g1, wo_h, epi=ops.EPI_BIAS, flags=ops.FLAG_OUT_H16 | ops.FLAG_NO_F32, w_mn=True)
        dqkv = ops.mha_bwd(qkv, heads, kv_len, att, datt, rmax, rinv, dropout_p=p_attn, dropout_seed=seed)
        dwqkv = ops.zeros((1, 3 * D, D), x)
        ops.gemm_wgrad(dqkv, xn, dwqkv)
        dbqkv = ops.colsum_h16(dqkv)
        dxn, _, _ = ops.gemm(dqkv, wqkv_h, epi=ops.EPI_BIAS, w_mn=True)
        dln1, dn1w, dn1b = ops.layernorm_bwd(dxn, x, n1w, eps)
        dx = dx1 + dln1
        dwq, dwk, dwv = dwqkv[0, :D], dwqkv[0, D:2 * D], dwqkv[0, 2 * D:]
        return (dx, None, None, None, None, None, None, None, dn1w, dn1b, dwq, dbqkv[:D], dwk, dbqkv[D:2 * D], dwv, dbqkv[2 * D:],
                dwo[0], dbo, dn2w, dn2b, dw1.view(U, D, 1), db1, dw2.view(D, U, 1), db2)


# --------------------------------------------------------------------------------------------------
# FastSpeech2 variance losses (duration / pitch / energy), one kernel for the three losses and their gradients
# --------------------------------------------------------------------------------------------------
class FastSpeech2LossFn(Function):
    @staticmethod
    def forward(ctx, d_hat, p_hat, e_hat, ds, p_tgt, e_tgt, x_len):
        losses, gd, gp, ge = ops.fs2_losses(d_hat.contiguous(), p_hat.contiguous(), e_hat.contiguous(), ds.float().contiguous(),
                                            p_tgt.float().contiguous(), e_tgt.float().contiguous(), x_len.contiguous())
        ctx.save_for_backward(gd, gp, ge)
        return losses[0], losses[1], losses[2]

    @staticmethod
    def backward(ctx, dl_d, dl_p, dl_e):
        gd, gp, ge = ctx.saved_tensors
        return gd * dl_d, gp * dl_p, ge * dl_e, None, None, None, None


class AlignLossFn(Function):
    """align_loss = ForwardSumLoss + bin loss (generator/__init__.py:174-175).  `per_sample_fs` / `fs_grad` come from
    ops.forward_sum (run on a side stream), `path` from ops.mas; the bin-loss gradient is folded into fs_grad in place."""

    @staticmethod
    def forward(ctx, log_p_attn, per_sample_fs, fs_grad, path, m_len):
        out = ops.align_loss_fold(log_p_attn.contiguous(), path, m_len.contiguous(), per_sample_fs, fs_grad)
        ctx.save_for_backward(fs_grad)
        fs_part, bin_part = out[1], out[2]
        ctx.mark_non_differentiable(fs_part, bin_part)
        return out[0], fs_part, bin_part

    @staticmethod
    def backward(ctx, g, _g1, _g2):
        (fs_grad,) = ctx.saved_tensors
        return fs_grad * g, None, None, None, None
