"""Functional fp32 restatement of the OptiSpeech generator hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function takes a flat `sd` (state_dict with
the reference's keys, see oracle/spec.py) and plain tensors; citations are reference file:line
(mush42/optispeech @ 3bdde20).  Data layout inside the oracle is channels-last (B, T, C) wherever
the reference's (B, C, T) is only a transposition artefact.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from .spec import ModelSpec, PredictorSpec

SD = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------------
# masks  (optispeech/utils/model.py:12-21)
# ----------------------------------------------------------------------------------------------
def sequence_mask(lengths: torch.Tensor, max_length: Optional[int] = None) -> torch.Tensor:
    """True where position < length."""
    if max_length is None:
        max_length = int(lengths.max())
    pos = torch.arange(int(max_length), dtype=lengths.dtype, device=lengths.device)
    return pos[None, :] < lengths[:, None]


# ----------------------------------------------------------------------------------------------
# text embedding  (generator/modules/core.py:10-31, layers.py:48-71)
# ----------------------------------------------------------------------------------------------
def text_embedding(sd: SD, x: torch.Tensor, spec: ModelSpec, prefix: str = "text_embedding"):
    """sqrt(dim) * E[x] + scale * [sin(pos * theta^-j/half) | cos(...)]; dropout is identity in eval."""
    dim = spec.dim
    embed = math.sqrt(dim) * F.embedding(x, sd[f"{prefix}.embed_tokens.weight"], padding_idx=0)
    half = dim // 2
    inv_freq = float(spec.max_source_positions) ** -(torch.arange(half).float() / half)
    ang = torch.arange(x.shape[1]).float()[:, None] * inv_freq[None, :]
    pos = torch.cat((ang.sin(), ang.cos()), dim=-1) * sd[f"{prefix}.embed_positions.scale"]
    return embed + pos, embed


# ----------------------------------------------------------------------------------------------
# ConvNeXt  (generator/modules/convnext.py:34-47, 92-103)
# ----------------------------------------------------------------------------------------------
def convnext_block(sd: SD, prefix: str, x: torch.Tensor, drop_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: (B, T, C) -> x + [drop] gamma * pw2(gelu(pw1(LN_1e-6(dwconv7(x)))))."""
    C = x.shape[-1]
    y = F.conv1d(x.transpose(1, 2), sd[f"{prefix}.dwconv.weight"], sd[f"{prefix}.dwconv.bias"], padding=3, groups=C)
    y = y.transpose(1, 2)
    y = F.layer_norm(y, (C,), sd[f"{prefix}.norm.weight"], sd[f"{prefix}.norm.bias"], 1e-6)
    y = F.linear(y, sd[f"{prefix}.pwconv1.weight"], sd[f"{prefix}.pwconv1.bias"])
    y = F.gelu(y)
    y = F.linear(y, sd[f"{prefix}.pwconv2.weight"], sd[f"{prefix}.pwconv2.bias"])
    y = sd[f"{prefix}.gamma"] * y
    if drop_scale is not None:  # DropPath with an injected per-sample scale (convnext.py:121-129)
        y = y * drop_scale[:, None, None]
    return x + y


def convnext_backbone(sd: SD, prefix: str, x: torch.Tensor, padding_mask: Optional[torch.Tensor], num_layers: int,
                      drop_scales=None) -> torch.Tensor:
    """x (B,T,C), padding_mask (B,T) True = pad.  Mask multiplies AFTER each block; final LN eps 1e-6."""
    keep = None if padding_mask is None else (1.0 - padding_mask.float())[..., None]
    for i in range(num_layers):
        x = convnext_block(sd, f"{prefix}.convnext.{i}", x, None if drop_scales is None else drop_scales[i])
        if keep is not None:
            x = x * keep
    C = x.shape[-1]
    return F.layer_norm(x, (C,), sd[f"{prefix}.final_layer_norm.weight"], sd[f"{prefix}.final_layer_norm.bias"], 1e-6)


# ----------------------------------------------------------------------------------------------
# Transformer backbone  (generator/modules/transformer.py:9-27; _transformer/encoder.py:271-313,
# encoder_layer.py:60-116, attention.py:38-125, multi_layer_conv.py:52-62, embedding.py:57-124, layer_norm.py:11-36)
# ----------------------------------------------------------------------------------------------
def positional_encoding_table(T: int, d_model: int) -> torch.Tensor:
    """embedding.py:57-73: pe[t, 2i] = sin(t * w_i), pe[t, 2i+1] = cos(t * w_i), w_i = exp(-2i ln(10000) / d)."""
    position = torch.arange(0, T, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
    pe = torch.zeros(T, d_model)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def multi_head_attention(sd: SD, prefix: str, x: torch.Tensor, key_valid: torch.Tensor, heads: int) -> torch.Tensor:
    """Self-attention of attention.py:107-125: softmax(masked_fill(QK^T / sqrt(d_k), finfo.min)).masked_fill(0) V, then
    linear_out.  key_valid (B, T) True = real key; only KEYS are masked (query rows at pads are computed like any other)."""
    B, T, D = x.shape
    dk = D // heads
    lin = lambda n, v: F.linear(v, sd[f"{prefix}.{n}.weight"], sd[f"{prefix}.{n}.bias"])  # noqa: E731
    q = lin("linear_q", x).view(B, T, heads, dk).transpose(1, 2)
    k = lin("linear_k", x).view(B, T, heads, dk).transpose(1, 2)
    v = lin("linear_v", x).view(B, T, heads, dk).transpose(1, 2)
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk)
    masked = ~key_valid[:, None, None, :]
    scores = scores.masked_fill(masked, torch.finfo(scores.dtype).min)
    attn = torch.softmax(scores, dim=-1).masked_fill(masked, 0.0)
    ctx = torch.matmul(attn, v).transpose(1, 2).contiguous().view(B, T, D)
    return lin("linear_out", ctx)


def transformer_backbone(sd: SD, prefix: str, x: torch.Tensor, padding_mask: torch.Tensor, blocks: int, heads: int) -> torch.Tensor:
    """x (B,T,C), padding_mask (B,T) True = pad.  Eval mode (all dropouts are identity): x + alpha * pe, then `blocks`
    pre-LN layers  x += MHA(LN(x));  x += w_2(relu(w_1(LN(x))))  (k=1 Conv1d = Linear), then after_norm.  LN eps 1e-12."""
    t = f"{prefix}.transformer"
    B, T, C = x.shape
    x = x + sd[f"{t}.embed.0.alpha"] * positional_encoding_table(T, C)[None]
    valid = ~padding_mask
    ln = lambda v, n: F.layer_norm(v, (C,), sd[f"{n}.weight"], sd[f"{n}.bias"], 1e-12)  # noqa: E731
    for i in range(blocks):
        p = f"{t}.encoders.{i}"
        x = x + multi_head_attention(sd, f"{p}.self_attn", ln(x, f"{p}.norm1"), valid, heads)
        h = F.relu(F.linear(ln(x, f"{p}.norm2"), sd[f"{p}.feed_forward.w_1.weight"][:, :, 0], sd[f"{p}.feed_forward.w_1.bias"]))
        x = x + F.linear(h, sd[f"{p}.feed_forward.w_2.weight"][:, :, 0], sd[f"{p}.feed_forward.w_2.bias"])
    return ln(x, f"{t}.after_norm")


def backbone(sd: SD, spec: ModelSpec, prefix: str, x: torch.Tensor, padding_mask: torch.Tensor, layers: int) -> torch.Tensor:
    """encoder / decoder dispatch on spec.backbone (configs/model/optispeech.yaml vs configs/model/transformer.yaml)."""
    if spec.backbone == "transformer":
        return transformer_backbone(sd, prefix, x, padding_mask, spec.tf_blocks, spec.tf_heads)
    return convnext_backbone(sd, prefix, x, padding_mask, layers)


# ----------------------------------------------------------------------------------------------
# variance predictors  (generator/modules/core.py:34-180, layers.py:26-45)
# ----------------------------------------------------------------------------------------------
def variance_predictor(sd: SD, prefix: str, x: torch.Tensor, padding_mask: torch.Tensor, ps: PredictorSpec):
    """[Conv1d(k, same) -> ReLU -> LN_1e-12 over channels]*L -> Linear(->1) -> 0 at pads.  x: (B,T,C)."""
    h = x.transpose(1, 2)
    for i in range(ps.num_layers):
        h = F.conv1d(h, sd[f"{prefix}.conv.{i}.0.weight"], sd[f"{prefix}.conv.{i}.0.bias"],
                     padding=(ps.kernel_size - 1) // 2)
        h = F.relu(h)
        c = h.shape[1]
        h = F.layer_norm(h.transpose(1, 2), (c,), sd[f"{prefix}.conv.{i}.2.weight"], sd[f"{prefix}.conv.{i}.2.bias"],
                         1e-12).transpose(1, 2)
    out = F.linear(h.transpose(1, 2), sd[f"{prefix}.linear.weight"], sd[f"{prefix}.linear.bias"]).squeeze(-1)
    return out.masked_fill(padding_mask, 0.0)


def duration_infer(sd: SD, x, padding_mask, ps: PredictorSpec, factor: float = 1.0, clip_val: float = 1e-8):
    """core.py:115-133: clamp(ceil((exp(logd) - clip) * factor).long(), 0), 0 at pads."""
    log_d = variance_predictor(sd, "duration_predictor", x, padding_mask, ps)
    d = torch.ceil((torch.exp(log_d) - clip_val) * factor)
    d = torch.clamp(d.long(), min=0)
    return d.masked_fill(padding_mask, 0), log_d


def variance_embed(sd: SD, prefix: str, x, padding_mask, value, ps: PredictorSpec):
    """core.py:152-176: x + Conv1d(1->dim, k9, same)(value), zeroed at pads."""
    emb = F.conv1d(value.unsqueeze(1), sd[f"{prefix}.embed.0.weight"], sd[f"{prefix}.embed.0.bias"],
                   padding=(ps.embed_kernel_size - 1) // 2)
    x = x + emb.transpose(1, 2)
    return x * (1.0 - padding_mask.float())[..., None]


# ----------------------------------------------------------------------------------------------
# length regulation  (generator/alignments.py:126-174, 283-297)
# ----------------------------------------------------------------------------------------------
def gaussian_upsampling(hs, ds, h_masks, d_masks, delta: float = 0.1):
    """hs (B,Tx,C), ds (B,Tx), h_masks (B,Tm) True=valid, d_masks (B,Tx) True=valid -> (B,Tm,C)."""
    ds = ds.clone()
    if ds.sum() == 0:  # alignments.py:152-157
        ds[ds.sum(dim=1).eq(0)] = 1
    B, Tm = h_masks.shape
    t = torch.arange(Tm, device=ds.device).float()[None, :].repeat(B, 1) * h_masks.float()
    c = ds.cumsum(dim=-1) - ds / 2
    energy = -delta * (t[:, :, None] - c[:, None, :]) ** 2
    energy = energy.masked_fill(~d_masks[:, None, :], -float("inf"))
    return torch.softmax(energy, dim=2) @ hs


def expand_by_duration(x, durations):
    """Hard repeat of x (B,Tx,C) by integer durations (B,Tx) -> ((B,Tm,C), lengths); zeros beyond each length."""
    lengths = durations.sum(dim=1)
    Tm = int(lengths.max())
    csum = torch.cumsum(F.pad(durations, (1, 0)), dim=1)  # (B, Tx+1)
    frame = torch.arange(Tm, device=x.device)[None, :, None]
    sel = (csum[:, None, :-1] <= frame) & (csum[:, None, 1:] > frame)
    return sel.to(x.dtype) @ x, lengths


def expand_indices(durations: torch.Tensor, Tm: int) -> torch.Tensor:
    """Integer form of expand_by_duration: source token of every frame, -1 beyond the length.  (bit-exact target)"""
    csum = torch.cumsum(durations, dim=1)
    frame = torch.arange(Tm, device=durations.device)[None, :]
    idx = torch.searchsorted(csum, frame.expand(durations.shape[0], -1).contiguous(), right=True)
    return torch.where(frame < csum[:, -1:], idx, torch.full_like(idx, -1))


# ----------------------------------------------------------------------------------------------
# alignment learning  (generator/alignments.py:14-123, 177-280; generator/loss.py:143-194)
# ----------------------------------------------------------------------------------------------
def beta_binomial_log_prior(T: int, N: int) -> np.ndarray:
    """log BetaBinomial(k; n=N, a=t, b=T-t+1) for t=1..T, k=0..N-1 -> (T, N) float64 (alignments.py:109-114).

    logpmf(k) = log C(n,k) + lbeta(k+a, n-k+b) - lbeta(a,b)   (scipy.stats.betabinom definition)."""
    from math import lgamma

    lg = np.vectorize(lgamma, otypes=[np.float64])
    t = np.arange(1, T + 1, dtype=np.float64)[:, None]
    a, b = t, T - t + 1.0
    k = np.arange(N, dtype=np.float64)[None, :]
    n = float(N)
    log_comb = lg(n + 1.0) - lg(k + 1.0) - lg(n - k + 1.0)

    def lbeta(p, q):
        return lg(p) + lg(q) - lg(p + q)

    return log_comb + lbeta(k + a, n - k + b) - lbeta(a, b)


def alignment_log_p_attn(sd: SD, text, feats, text_lengths, feats_lengths, x_masks, prefix="alignment_module"):
    """text (B,Tx,C), feats (B,Tm,F) -> log_p_attn (B,Tm,Tx) incl. beta-binomial prior (alignments.py:41-83)."""
    w = lambda n: (sd[f"{prefix}.{n}.weight"], sd[f"{prefix}.{n}.bias"])  # noqa: E731
    te = text.transpose(1, 2)
    te = F.relu(F.conv1d(te, *w("t_conv1"), padding=1))
    te = F.conv1d(te, *w("t_conv2")).transpose(1, 2)
    fe = feats.transpose(1, 2)
    fe = F.relu(F.conv1d(fe, *w("f_conv1"), padding=1))
    fe = F.relu(F.conv1d(fe, *w("f_conv2"), padding=1))
    fe = F.conv1d(fe, *w("f_conv3")).transpose(1, 2)
    dist = torch.norm(fe.unsqueeze(2) - te.unsqueeze(1), p=2, dim=3)
    score = (-dist).masked_fill(x_masks.unsqueeze(-2), -np.inf)
    log_p = F.log_softmax(score, dim=-1)
    B, Tm, Tx = log_p.shape
    prior = torch.full((B, Tm, Tx), -np.inf)
    for b in range(B):
        T, N = int(feats_lengths[b]), int(text_lengths[b])
        prior[b, :T, :N] = torch.from_numpy(beta_binomial_log_prior(T, N))
    return log_p + prior.to(log_p.dtype)


def monotonic_alignment_search(lp: np.ndarray) -> np.ndarray:
    """lp (T_mel, T_inp) float32 -> A (T_mel,) token of each frame (alignments.py:177-207).

    Q is float64; row 0 is a float32 running sum widened to float64; recursion
    Q[i,j] = max(Q[i-1,j-1], Q[i,j-1]) + lp[j,i]; backtrack prefers the lower token on ties (>=)."""
    T, N = lp.shape
    logp = np.ascontiguousarray(lp.T)  # (N, T)
    Q = np.full((N, T), -np.inf)
    Q[0, :] = np.cumsum(logp[0, :], dtype=np.float32)  # numba sums a float32 slice in float32
    for j in range(1, T):
        hi = min(j + 1, N)
        if hi > 1:
            Q[1:hi, j] = np.maximum(Q[0:hi - 1, j - 1], Q[1:hi, j - 1]) + logp[1:hi, j]
    A = np.full((T,), N - 1, dtype=np.int64)
    for j in range(T - 2, -1, -1):
        ib = A[j + 1]
        if ib == 0:
            A[j] = 0
        elif Q[ib - 1, j] >= Q[ib, j]:
            A[j] = ib - 1
        else:
            A[j] = ib
    return A


def viterbi_decode(log_p_attn, text_lengths, feats_lengths):
    """-> durations (B,Tx) float32, bin_loss scalar (alignments.py:210-239)."""
    B, _, Tx = log_p_attn.shape
    ds = torch.zeros((B, Tx))
    bin_loss = 0
    for b in range(B):
        T, N = int(feats_lengths[b]), int(text_lengths[b])
        cur = log_p_attn[b, :T, :N]
        A = monotonic_alignment_search(cur.detach().float().cpu().numpy())
        cnt = np.bincount(A)
        ds[b, : len(cnt)] = torch.from_numpy(cnt).float()
        bin_loss = bin_loss - cur[torch.arange(T), torch.from_numpy(A)].mean()
    return ds, bin_loss / B


def average_by_duration(ds, xs, text_lengths, feats_lengths):
    """Token-level mean of frame-level xs (B,Tm) over duration spans; 0 for empty spans (alignments.py:242-280)."""
    B, Tx = ds.shape
    out = torch.zeros((B, Tx), dtype=torch.float32)
    d_int = ds.to(torch.int32)
    for b in range(B):
        n, T = int(text_lengths[b]), int(feats_lengths[b])
        x = xs[b, :T].float()
        edges = torch.cat([torch.zeros(1, dtype=torch.int64), d_int[b, :n].to(torch.int64).cumsum(0)])
        for i in range(n):
            seg = x[int(edges[i]): int(edges[i + 1])]
            out[b, i] = seg.mean() if seg.numel() else 0.0
    return out


def forward_sum_loss(log_p_attn, ilens, olens, blank_logprob: float = -1.0):
    """CTC forward-sum alignment loss (loss.py:150-194): blank column log(e^-1), per-sample re-normalisation,
    F.ctc_loss(reduction='mean' -> / target length, zero_infinity=True), mean over the batch."""
    B = log_p_attn.size(0)
    padded = F.pad(log_p_attn, (1, 0, 0, 0, 0, 0), value=blank_logprob)
    loss = 0
    for b in range(B):
        n, T = int(ilens[b]), int(olens[b])
        lp = F.log_softmax(padded[b, :T, : n + 1].unsqueeze(1), dim=-1)
        loss = loss + F.ctc_loss(lp, torch.arange(1, n + 1).unsqueeze(0), input_lengths=olens[b: b + 1],
                                 target_lengths=ilens[b: b + 1], zero_infinity=True)
    return loss / B


def forward_sum_loss_explicit(log_p_attn, ilens, olens, blank_logprob: float = -1.0):
    """Same quantity from the explicit CTC alpha recursion (the published algorithm behind F.ctc_loss),
    used to pin the semantics the CUDA kernel implements: extended target [blank, 1, blank, 2, ..., N, blank]."""
    B = log_p_attn.size(0)
    total = 0
    for b in range(B):
        n, T = int(ilens[b]), int(olens[b])
        lp = F.log_softmax(F.pad(log_p_attn[b, :T, :n], (1, 0), value=blank_logprob), dim=-1)  # (T, n+1)
        S = 2 * n + 1
        ext = torch.zeros(S, dtype=torch.long)
        ext[1::2] = torch.arange(1, n + 1)
        ninf = torch.tensor(-float("inf"))
        alpha = torch.full((S,), -float("inf"))
        alpha[0] = lp[0, 0]
        if S > 1:
            alpha[1] = lp[0, 1]
        for t in range(1, T):
            a1 = torch.cat([ninf[None], alpha[:-1]])
            a2 = torch.cat([ninf[None], ninf[None], alpha[:-2]])
            a2 = torch.where(torch.arange(S) % 2 == 1, a2, ninf)  # skip only between distinct labels
            alpha = torch.logsumexp(torch.stack([alpha, a1, a2]), dim=0) + lp[t, ext]
        ll = torch.logsumexp(torch.stack([alpha[-1], alpha[-2]]) if S > 1 else alpha[-1:], dim=0)
        nll = -ll
        nll = torch.where(torch.isinf(nll), torch.zeros_like(nll), nll)
        total = total + nll / n
    return total / B


def fastspeech2_losses(d_outs, p_outs, e_outs, ds, ps, es, ilens):
    """loss.py:83-140 as the reference *actually evaluates it* (use_masking=True, (B,Tx,1) inputs).

    The reference's masks have a stray singleton axis (make_non_pad_mask -> (B,1,Tx), utils/model.py:19-21), so
    `masked_select` broadcasts instead of selecting the non-pad positions:
      * duration: d_outs (B,Tx,1) x mask (B,1,Tx) -> element (b,i) is taken len_b times, for EVERY i < Tx
        (padded positions included: prediction 0 vs target log(0 + 1e-8));
      * pitch / energy: outs (B,Tx,1) x mask (B,1,Tx,1) -> element (b,i) is taken once per sample a with len_a > i.
    Both reduce with 'mean', i.e. a weighted mean with those multiplicities.  MSE in the log domain for durations
    (loss.py:31-47), SmoothL1(beta=1) for pitch and energy (loss.py:77-78)."""
    B, Tx = d_outs.shape
    lens = ilens.to(torch.float32)
    d_err = (d_outs - torch.log(ds.float() + 1e-8)) ** 2                      # (B, Tx)
    dl = (d_err.sum(dim=1) * lens).sum() / (lens.sum() * Tx)
    w = (torch.arange(Tx)[None, :] < ilens[:, None]).to(torch.float32).sum(dim=0)  # (Tx,) samples longer than i
    denom = w.sum() * B
    pl = (F.smooth_l1_loss(p_outs, ps, reduction="none") * w[None, :]).sum() / denom
    el = (F.smooth_l1_loss(e_outs, es, reduction="none") * w[None, :]).sum() / denom
    return dl, pl, el


# ----------------------------------------------------------------------------------------------
# WaveNeXt  (vocoder/wavenext/__init__.py:31-48, 82-86)
# ----------------------------------------------------------------------------------------------
def wavenext(sd: SD, x: torch.Tensor, padding_mask: Optional[torch.Tensor], spec: ModelSpec, prefix: str = "vocoder",
             drop_scales=None) -> torch.Tensor:
    """x (B,T,dim) channels-last -> wav (B, T*hop), clipped to [-1, 1]."""
    h = F.conv1d(x.transpose(1, 2), sd[f"{prefix}.embed.weight"], sd[f"{prefix}.embed.bias"], padding=3).transpose(1, 2)
    h = F.layer_norm(h, (spec.voc_dim,), sd[f"{prefix}.norm.weight"], sd[f"{prefix}.norm.bias"], 1e-6)
    h = convnext_backbone(sd, f"{prefix}.backbone", h, padding_mask, spec.voc_layers, drop_scales)
    h = F.linear(h, sd[f"{prefix}.head.linear_1.weight"], sd[f"{prefix}.head.linear_1.bias"])
    h = F.linear(h, sd[f"{prefix}.head.linear_2.weight"])
    return torch.clip(h.reshape(h.shape[0], -1), min=-1.0, max=1.0)


# ----------------------------------------------------------------------------------------------
# segments  (utils/segments.py:12-72)
# ----------------------------------------------------------------------------------------------
def segment_starts(num_frames: torch.Tensor, segment_size: int, rand: torch.Tensor) -> torch.Tensor:
    """start = floor(rand * max(num_frames - segment, 0)) with `rand` ~ U[0,1) drawn by the caller."""
    max_start = (num_frames - segment_size).clamp(min=0)
    return (rand * max_start).to(torch.long)


def get_segments(x: torch.Tensor, starts: torch.Tensor, segment_size: int) -> torch.Tensor:
    """x (B,T,C) channels-last -> (B,segment,C)."""
    return torch.stack([x[b, int(s): int(s) + segment_size] for b, s in enumerate(starts)])


# ----------------------------------------------------------------------------------------------
# generator.synthesise / generator.forward  (generator/__init__.py:72-301)
# ----------------------------------------------------------------------------------------------
@torch.no_grad()
def synthesise(sd: SD, spec: ModelSpec, x, x_lengths, d_factor=1.0, p_factor=1.0, e_factor=1.0, durations=None):
    """-> dict(wav, wav_lengths, durations, pitch, energy, y (decoder out), f0_cond).  `durations` injects
    integer durations (tests use it to decouple waveform parity from +-1-frame rounding flips)."""
    x_mask = sequence_mask(x_lengths, int(x_lengths.max()))
    pad = ~x_mask
    h, _ = text_embedding(sd, x, spec)
    h = backbone(sd, spec, "encoder", h, pad, spec.enc_layers)
    d_pred, log_d = duration_infer(sd, h, pad, spec.duration, d_factor)
    if durations is None:
        durations = d_pred
    pitch = variance_predictor(sd, "pitch_predictor.predictor", h, pad, spec.pitch) * p_factor
    h = variance_embed(sd, "pitch_predictor", h, pad, pitch, spec.pitch)
    energy = variance_predictor(sd, "energy_predictor.predictor", h, pad, spec.energy) * e_factor
    h = variance_embed(sd, "energy_predictor", h, pad, energy, spec.energy)
    y_lengths = durations.sum(dim=1)
    y_mask = sequence_mask(y_lengths, int(y_lengths.max()))
    y = gaussian_upsampling(h, durations, y_mask, x_mask)
    y = backbone(sd, spec, "decoder", y, ~y_mask, spec.dec_layers)
    f0_cond, _ = expand_by_duration(pitch.unsqueeze(-1), durations)
    wav = wavenext(sd, y, ~y_mask, spec)
    return dict(wav=wav, wav_lengths=y_lengths * spec.hop_length, durations=durations, pitch=pitch, energy=energy,
                log_durations=log_d, y=y, f0_cond=f0_cond, encoder_out=h)


def generator_forward(sd: SD, spec: ModelSpec, x, x_lengths, mel, mel_lengths, pitches, energies, seg_rand):
    """Training forward (generator/__init__.py:72-192), eval-mode (no dropout / DropPath).

    mel is (B, F, Tm) as the collate function provides it; `seg_rand` (B,) replaces torch.rand in
    get_random_segments.  Returns the reference's dict plus intermediates used by parity tests."""
    x_mask = sequence_mask(x_lengths, int(x_lengths.max()))
    mel_mask = sequence_mask(mel_lengths, int(mel_lengths.max()))
    in_pad, tgt_pad = ~x_mask, ~mel_mask
    h, _ = text_embedding(sd, x, spec)
    h = backbone(sd, spec, "encoder", h, in_pad, spec.enc_layers)
    log_p_attn = alignment_log_p_attn(sd, h, mel.transpose(1, 2), x_lengths, mel_lengths, in_pad)
    durations, bin_loss = viterbi_decode(log_p_attn, x_lengths, mel_lengths)
    duration_hat = variance_predictor(sd, "duration_predictor", h.detach(), in_pad, spec.duration)
    p_avg = average_by_duration(durations, pitches, x_lengths, mel_lengths)
    e_avg = average_by_duration(durations, energies, x_lengths, mel_lengths)
    pitch_hat = variance_predictor(sd, "pitch_predictor.predictor", h, in_pad, spec.pitch)
    h = variance_embed(sd, "pitch_predictor", h, in_pad, p_avg, spec.pitch)
    energy_hat = variance_predictor(sd, "energy_predictor.predictor", h, in_pad, spec.energy)
    h = variance_embed(sd, "energy_predictor", h, in_pad, e_avg, spec.energy)
    y = gaussian_upsampling(h, durations, mel_mask, x_mask)
    y = backbone(sd, spec, "decoder", y, tgt_pad, spec.dec_layers)
    segment_size = min(spec.segment_size, y.shape[1])
    start_idx = segment_starts((mel_lengths - 4).to(y.dtype), segment_size, seg_rand)
    segment = get_segments(y, start_idx, segment_size)
    wav_hat = wavenext(sd, segment.detach(), None, spec)
    d_loss, p_loss, e_loss = fastspeech2_losses(duration_hat, pitch_hat, energy_hat, durations, p_avg, e_avg, x_lengths)
    fs_loss = forward_sum_loss(log_p_attn, x_lengths, mel_lengths)
    align_loss = fs_loss + bin_loss
    loss = (align_loss * spec.lambda_align + d_loss * spec.lambda_duration + p_loss * spec.lambda_pitch
            + e_loss * spec.lambda_energy)
    return dict(wav_hat=wav_hat, start_idx=start_idx, segment_size=segment_size, loss=loss, align_loss=align_loss,
                duration_loss=d_loss, pitch_loss=p_loss, energy_loss=e_loss, forwardsum_loss=fs_loss, bin_loss=bin_loss,
                log_p_attn=log_p_attn, durations=durations, pitch_avg=p_avg, energy_avg=e_avg, decoder_out=y,
                duration_hat=duration_hat, pitch_hat=pitch_hat, energy_hat=energy_hat)


def crop_wav_segments(wav: np.ndarray, start_idx: torch.Tensor, segment_size: int, hop: int) -> torch.Tensor:
    """Ground-truth crop (base_lightning_module.py:38-43, utils/segments.py:63-72): wav (B, Tw) numpy."""
    n = segment_size * hop
    out = np.zeros((wav.shape[0], n), dtype=np.float32)
    for b, s in enumerate(start_idx.tolist()):
        out[b] = wav[b, s * hop: s * hop + n]
    return torch.from_numpy(out)
