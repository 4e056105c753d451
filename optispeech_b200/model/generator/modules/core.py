"""Text embedding and variance predictors on the B200 path.

Mirrors optispeech/model/generator/modules/core.py (reference @ 3bdde20): class names, constructor
quirks (`PitchPredictor` reads kwargs["dim"] / kwargs["conv_layer_class"]; `DurationPredictor(*args,
clip_val=1e-8, **kwargs)`; `TextEmbedding.forward` returns a tuple) and state_dict keys
(`conv.{i}.0.*` conv, `conv.{i}.2.*` LayerNorm, `linear.*`, `embed.0.*`).

Each predictor layer Conv1d(k) + ReLU + LayerNorm(eps 1e-12) is ONE tcgen05 implicit-GEMM launch
(osb_gemm, EPI_RELU_LN; the CTA owns whole rows so the normalisation is thread-local); the final
Linear(->1) + masked_fill rides in the last layer's epilogue (FLAG_DOT).
"""
from __future__ import annotations

import math

import torch
from torch import nn

from .... import ops, precision
from ...packing import PackedCache, pack_conv
from .layers import LayerNorm, ScaledSinusoidalEmbedding

DEFAULT_MAX_SOURCE_POSITIONS = 2000



def _as_u8(mask: torch.Tensor) -> torch.Tensor:
    """Byte view of a boolean mask (no cast launch when it is already a contiguous bool tensor)."""
    if mask.dtype == torch.bool and mask.is_contiguous():
        return mask.view(torch.uint8)
    return mask.to(torch.uint8).contiguous()


class TextEmbedding(nn.Module):
    def __init__(self, dim: int, n_vocab: int, dropout: float = 0.0, padding_idx: int = 0,
                 max_source_positions: int = DEFAULT_MAX_SOURCE_POSITIONS):
        super().__init__()
        self.embed_scale = math.sqrt(dim)
        self.embed_tokens = nn.Embedding(n_vocab, dim, padding_idx)
        self.embed_positions = ScaledSinusoidalEmbedding(dim, theta=max_source_positions)
        self.emb_dropout = nn.Dropout(dropout)

    def forward(self, src_tokens):
        """-> (x, embed) like the reference; `embed` (token part only) is not materialised on this path."""
        if torch.is_grad_enabled():
            from ....autograd import EmbedTextFn

            x = EmbedTextFn.apply(src_tokens.contiguous(), self.embed_tokens.weight, self.embed_positions.scale,
                                  self.embed_positions.inv_freq, self.embed_tokens.padding_idx)
        else:
            x = ops.embed_text(src_tokens.contiguous(), self.embed_tokens.weight, self.embed_positions.inv_freq,
                               self.embed_positions.scale)
        x = self.emb_dropout(x)
        return x, None


NARROW_PREDICTOR_MAX_TILES = 37   # row tiles up to which a predictor layer runs as narrow-tile GEMM + row kernel (forward_h16)


class VariancePredictor(nn.Module):
    def __init__(self, dim: int, num_layers: int, intermediate_dim: int, kernel_size: int, dropout: float = 0.1,
                 conv_layer_class: type = torch.nn.Conv1d):
        super().__init__()
        self.dim = dim
        self.conv_layer_class = conv_layer_class
        self.kernel_size = kernel_size
        self.conv = torch.nn.ModuleList()
        for idx in range(num_layers):
            input_dim = dim if idx == 0 else intermediate_dim
            self.conv += [
                torch.nn.Sequential(
                    self.conv_layer_class(input_dim, intermediate_dim, kernel_size, padding=(kernel_size - 1) // 2),
                    torch.nn.ReLU(),
                    LayerNorm(intermediate_dim, dim=1),
                    torch.nn.Dropout(dropout),
                )
            ]
        self.linear = torch.nn.Linear(intermediate_dim, 1)
        self._packed = PackedCache()

    def packed(self):
        srcs = [layer[0].weight for layer in self.conv]
        return self._packed.get("fwd", srcs, lambda: [pack_conv(layer[0].weight) for layer in self.conv])

    def forward_h16(self, x_h16: torch.Tensor, pad_mask_u8: torch.Tensor, split: bool = False) -> torch.Tensor:
        """x fp16 (B,T,dim) (or split (B,T,2*dim)), pad mask (B,T) uint8 -> (B,T) fp32 predictions, 0 at pads."""
        ws = self.packed()
        fl = (ops.FLAG_SPLIT_IN | ops.FLAG_SPLIT_OUT) if split else 0
        pad = (self.kernel_size - 1) // 2
        h = x_h16
        n = len(self.conv)
        B, T = x_h16.shape[0], x_h16.shape[1]
        if split and B * ((T + 127) // 128) <= NARROW_PREDICTOR_MAX_TILES:
            # A few row tiles (synthesis of one or a few utterances): the fused ReLU + LayerNorm epilogue needs whole rows, i.e.
            # ONE CTA per 128 rows streams the layer's whole hi + lo weight matrix through one SM's L2 port (33 us per layer at
            # B=1).  A plain-bias GEMM in 64-wide tiles spreads the weight rows over 4-6x the SMs and a row kernel does
            # ReLU + LayerNorm (+ the final Linear) on the fp32 result — same arithmetic, no extra rounding.
            for i, layer in enumerate(self.conv):
                conv, ln = layer[0], layer[2]
                z, _, _ = ops.gemm(h, ws[i], epi=ops.EPI_BIAS, flags=ops.FLAG_SPLIT_IN, pad=pad, bias=conv.bias)
                if i < n - 1:
                    h, _ = ops.relu_layernorm(z, ln.weight, ln.bias, ln.eps, h16=True, split=True)
                else:
                    _, out = ops.relu_layernorm(z, ln.weight, ln.bias, ln.eps, h16=False, dot_w=self.linear.weight.view(-1),
                                                dot_b=self.linear.bias, pad_mask=pad_mask_u8)
            return out
        for i, layer in enumerate(self.conv):
            conv, ln = layer[0], layer[2]
            last = i == n - 1
            if not last:
                h, _, _ = ops.gemm(h, ws[i], epi=ops.EPI_RELU_LN, flags=fl, pad=pad, bias=conv.bias, ln_w=ln.weight, ln_b=ln.bias,
                                   ln_eps=ln.eps)
            else:
                _, _, out = ops.gemm(h, ws[i], epi=ops.EPI_RELU_LN, flags=ops.FLAG_DOT | (ops.FLAG_SPLIT_IN if split else 0), pad=pad,
                                     bias=conv.bias,
                                     ln_w=ln.weight, ln_b=ln.bias, ln_eps=ln.eps, dot_w=self.linear.weight.view(-1),
                                     dot_b=self.linear.bias,
                                     pad_mask=pad_mask_u8)
        return out

    def forward(self, x: torch.Tensor, padding_mask) -> torch.Tensor:
        """Reference signature: x (B,T,dim) fp32, padding_mask (B,T) bool -> (B,T)."""
        mask_u8 = _as_u8(padding_mask)
        if torch.is_grad_enabled():
            from ....autograd import VariancePredictorFn

            params = []
            for layer in self.conv:
                params += [layer[0].weight, layer[0].bias, layer[2].weight, layer[2].bias]
            p_drop = float(self.conv[0][3].p) if self.training else 0.0
            seed = int(torch.randint(0, 2 ** 62, (), dtype=torch.int64)) if p_drop > 0.0 else 0  # CPU generator: no device sync
            return VariancePredictorFn.apply(x, mask_u8, self.kernel_size, self.conv[0][2].eps, p_drop, seed, self.linear.weight,
                                             self.linear.bias, *params)
        split = precision.use_split(self.training)
        return self.forward_h16(ops.to_h16(x.contiguous(), split=split), mask_u8, split)


class DurationPredictor(VariancePredictor):
    def __init__(self, *args, clip_val=1e-8, **kwargs):
        super().__init__(*args, **kwargs)
        self.clip_val = clip_val

    @torch.inference_mode()
    def infer(self, x, mask, factor=1.0, x_h16=None):
        """-> (durations int64 (B,T), lengths int64 (B,)); reference returns durations only (core.py:115-133).
        `x_h16`, when given, must be in the current inference operand format (see precision.use_split)."""
        mask_u8 = _as_u8(mask)
        split = precision.use_split(False)
        log_d = self.forward_h16(x_h16 if x_h16 is not None else ops.to_h16(x.contiguous(), split=split), mask_u8, split)
        return ops.durations(log_d, mask_u8, factor, self.clip_val)


class PitchPredictor(nn.Module):
    def __init__(self, *args, embed_kernel_size=9, embed_dropout=0.1, **kwargs):
        super().__init__()
        self.predictor = VariancePredictor(*args, **kwargs)
        self.dim = kwargs["dim"]
        self.conv_layer_class = kwargs["conv_layer_class"]
        self.embed = torch.nn.Sequential(
            self.conv_layer_class(in_channels=1, out_channels=self.dim, kernel_size=embed_kernel_size,
                                  padding=(embed_kernel_size - 1) // 2),
            torch.nn.Dropout(embed_dropout),
        )

    def _embed_add(self, x, values, mask_u8, want_h16, split=False):
        conv = self.embed[0]
        return ops.variance_embed(x.contiguous(), values.contiguous(), conv.weight.view(self.dim, -1), conv.bias, mask_u8,
                                  f32=True, h16=want_h16, split=split)

    def predict_ahead(self, x: torch.Tensor, padding_mask: torch.Tensor, side_stream):
        """The prediction stack alone, queued on `side_stream` forked from the current stream NOW.  The stack reads `x` only,
        not the targets, so the training forward calls this as soon as `x` exists — before the alignment search that produces
        the targets — and hands the result to forward(preds=...); the CALLER joins the stream before using it."""
        side_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side_stream):
            return self.predictor(x, padding_mask)

    def forward(self, x: torch.Tensor, padding_mask: torch.Tensor, target: torch.Tensor, side_stream=None, preds=None):
        """Teacher-forced: returns (x + embed(target), preds) (reference core.py:152-166), eval-mode numerics.
        The prediction only meets the rest of the step at the loss (the embedding is driven by `target`), so with
        `side_stream` it is enqueued there, forked from the current stream; the CALLER joins (`wait_stream`) before using it.
        `preds`: the prediction computed earlier by predict_ahead() on the same `x`."""
        mask_u8 = _as_u8(padding_mask)
        if torch.is_grad_enabled():
            from ....autograd import VarianceEmbedFn

            if preds is not None:
                pass
            elif side_stream is not None and x.is_cuda:
                side_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side_stream):
                    preds = self.predictor(x, padding_mask)
            else:
                preds = self.predictor(x, padding_mask)
            conv, drop = self.embed[0], self.embed[1]
            emb_scale = None
            if self.training and drop.p > 0.0:  # Dropout on the embedding branch (core.py:143-150)
                emb_scale = torch.empty(x.shape, device=x.device, dtype=torch.float32).bernoulli_(1.0 - drop.p).div_(1.0 - drop.p)
            out = VarianceEmbedFn.apply(x, target.contiguous(), conv.weight, conv.bias, mask_u8, emb_scale)
            return out, preds
        split = precision.use_split(self.training)
        preds = self.predictor.forward_h16(ops.to_h16(x.contiguous(), split=split), mask_u8, split)
        out, _ = self._embed_add(x, target, mask_u8, False)
        return out, preds

    @torch.inference_mode()
    def infer(self, x, padding_mask, factor=1.0, x_h16=None, want_h16=False):
        mask_u8 = _as_u8(padding_mask)
        split = precision.use_split(False)
        preds = self.predictor.forward_h16(x_h16 if x_h16 is not None else ops.to_h16(x.contiguous(), split=split), mask_u8, split)
        preds = preds * factor
        o32, o16 = self._embed_add(x, preds, mask_u8, want_h16, split)
        if want_h16:
            return o32, preds, o16
        return o32, preds


class EnergyPredictor(PitchPredictor):
    pass
