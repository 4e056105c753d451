"""Transformer backbone on the B200 path (BASELINE config 4, configs/model/transformer.yaml).

Mirrors the reference's wrapper `Transformer` (optispeech/model/generator/modules/transformer.py:9-27) around the espnet
`Encoder` (modules/_transformer/encoder.py:24-313) for the configuration the YAMLs select — `input_layer=None`, pre-LN
(`normalize_before`), no `concat_after`, `conv1d` position-wise layers with kernel size 1, scaled positional encoding,
`selfattn` — with the reference's module tree, so `state_dict` keys are identical:

    transformer.embed.0.alpha
    transformer.encoders.{i}.self_attn.linear_{q,k,v,out}.{weight,bias}
    transformer.encoders.{i}.feed_forward.w_{1,2}.{weight,bias}        (Conv1d weights (out, in, 1))
    transformer.encoders.{i}.norm{1,2}.{weight,bias}
    transformer.after_norm.{weight,bias}

The nn.Linear / nn.Conv1d / nn.LayerNorm children are parameter containers; compute goes through libosb200:

    LayerNorm (eps 1e-12) -> fp16 operand            osb_layernorm
    fused q|k|v projection (N = 3*dim)                osb_gemm EPI_BIAS (fp16 output only)
    softmax(QK^T/sqrt(d_k)) V, key mask, dropout      osb_mha_fwd  (tcgen05; scores never reach HBM)
    linear_out + dropout + residual                   osb_gemm EPI_RESID
    w_1 + ReLU + dropout, w_2 + dropout + residual    osb_gemm EPI_RELU / EPI_RESID
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from .... import ops, precision
from ...packing import PackedCache, pack_linear


def _seed(p: float) -> int:
    return int(torch.randint(0, 2 ** 62, (), dtype=torch.int64)) if p > 0.0 else 0  # CPU generator: no device sync


class LayerNorm(nn.LayerNorm):
    """espnet LayerNorm (modules/_transformer/layer_norm.py:11-36): eps 1e-12, over the last dimension."""

    def __init__(self, nout: int, dim: int = -1):
        super().__init__(nout, eps=1e-12)
        self.dim = dim


class PositionalEncoding(nn.Module):
    """modules/_transformer/embedding.py:24-89: x * sqrt(d) + pe (not selected by the shipped configs; kept for the class path)."""

    def __init__(self, d_model: int, dropout_rate: float, max_len: int = 5000, reverse: bool = False):
        super().__init__()
        self.d_model = d_model
        self.xscale = math.sqrt(d_model)
        self.dropout_rate = dropout_rate
        self._pe: Optional[torch.Tensor] = None

    def table(self, T: int, device) -> torch.Tensor:
        """pe[t, 2i] = sin(t w_i), pe[t, 2i+1] = cos(t w_i): built exactly as the reference does (fp32, CPU), cached on the device."""
        if self._pe is None or self._pe.shape[0] < T or self._pe.device != device:
            with torch.inference_mode(False):  # a cached constant: must stay usable by autograd after synthesise() built it
                n = max(T, 1024)
                position = torch.arange(0, n, dtype=torch.float32).unsqueeze(1)
                div_term = torch.exp(torch.arange(0, self.d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / self.d_model))
                pe = torch.zeros(n, self.d_model)
                pe[:, 0::2] = torch.sin(position * div_term)
                pe[:, 1::2] = torch.cos(position * div_term)
                self._pe = pe.to(device)
        return self._pe


class ScaledPositionalEncoding(PositionalEncoding):
    """modules/_transformer/embedding.py:91-124: x + alpha * pe, then dropout."""

    def __init__(self, d_model: int, dropout_rate: float, max_len: int = 5000):
        super().__init__(d_model=d_model, dropout_rate=dropout_rate, max_len=max_len)
        self.alpha = nn.Parameter(torch.tensor(1.0))

    def reset_parameters(self):
        self.alpha.data = torch.tensor(1.0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        p = self.dropout_rate if self.training else 0.0
        pe = self.table(x.shape[1], x.device)
        if torch.is_grad_enabled() and (x.requires_grad or self.alpha.requires_grad):
            from ....autograd import PosEncFn

            return PosEncFn.apply(x, self.alpha, pe, p, _seed(p))
        return ops.add_posenc(x.contiguous(), pe, self.alpha.detach().reshape(1), p, _seed(p))


class MultiHeadedAttention(nn.Module):
    """Parameter container of modules/_transformer/attention.py:14-37 (the compute is in EncoderLayer)."""

    def __init__(self, n_head: int, n_feat: int, dropout_rate: float):
        super().__init__()
        assert n_feat % n_head == 0
        self.d_k = n_feat // n_head
        self.h = n_head
        if self.d_k != ops.MHA_DK:
            raise NotImplementedError(f"osb_mha kernels are built for d_k = {ops.MHA_DK} (got {self.d_k})")
        self.linear_q = nn.Linear(n_feat, n_feat)
        self.linear_k = nn.Linear(n_feat, n_feat)
        self.linear_v = nn.Linear(n_feat, n_feat)
        self.linear_out = nn.Linear(n_feat, n_feat)
        self.dropout_rate = dropout_rate


class MultiLayeredConv1d(nn.Module):
    """Parameter container of modules/_transformer/multi_layer_conv.py:11-62 (kernel size 1 = per-position Linear)."""

    def __init__(self, in_chans: int, hidden_chans: int, kernel_size: int, dropout_rate: float):
        super().__init__()
        if kernel_size != 1:
            raise NotImplementedError("position-wise Conv1d kernel size 1 only (configs/model/transformer.yaml)")
        self.w_1 = nn.Conv1d(in_chans, hidden_chans, kernel_size, stride=1, padding=(kernel_size - 1) // 2)
        self.w_2 = nn.Conv1d(hidden_chans, in_chans, kernel_size, stride=1, padding=(kernel_size - 1) // 2)
        self.dropout_rate = dropout_rate


class EncoderLayer(nn.Module):
    """modules/_transformer/encoder_layer.py:16-116 with normalize_before=True, concat_after=False, stochastic depth 0."""

    def __init__(self, size: int, self_attn: MultiHeadedAttention, feed_forward: MultiLayeredConv1d, dropout_rate: float):
        super().__init__()
        self.self_attn = self_attn
        self.feed_forward = feed_forward
        self.norm1 = LayerNorm(size)
        self.norm2 = LayerNorm(size)
        self.dropout_rate = dropout_rate
        self.size = size
        self.register_buffer("_ones", torch.ones(size), persistent=False)
        self._packed = PackedCache()

    def packed(self):
        a, f = self.self_attn, self.feed_forward
        srcs = [a.linear_q.weight, a.linear_q.bias, a.linear_k.weight, a.linear_k.bias, a.linear_v.weight, a.linear_v.bias,
                a.linear_out.weight, f.w_1.weight, f.w_2.weight]

        def build():
            wqkv = pack_linear(torch.cat([a.linear_q.weight, a.linear_k.weight, a.linear_v.weight], dim=0))
            bqkv = torch.cat([a.linear_q.bias, a.linear_k.bias, a.linear_v.bias]).contiguous()
            return wqkv, bqkv, pack_linear(a.linear_out.weight), pack_linear(f.w_1.weight[:, :, 0]), pack_linear(f.w_2.weight[:, :, 0])

        return self._packed.get("fwd", srcs, build)

    def forward_cl(self, x: torch.Tensor, kv_len: torch.Tensor, split: bool) -> torch.Tensor:
        """Inference kernels (no autograd): x fp32 (B,T,C) -> fp32 (B,T,C).  Dropout is drawn when the module is in train mode
        (the reference's decoder runs in train mode but receives no gradient, generator/__init__.py:161)."""
        a, f = self.self_attn, self.feed_forward
        wqkv, bqkv, wo, w1, w2 = self.packed()
        pd = self.dropout_rate if self.training else 0.0
        pa = a.dropout_rate if self.training else 0.0
        pf = f.dropout_rate if self.training else 0.0
        fin = ops.FLAG_SPLIT_IN if split else 0
        fout = ops.FLAG_SPLIT_OUT if split else 0
        _, xn = ops.layernorm(x, self.norm1.weight, self.norm1.bias, self.norm1.eps, f32=False, h16=True, split=split)
        _, qkv, _ = ops.gemm(xn, wqkv, epi=ops.EPI_BIAS, flags=fin | ops.FLAG_OUT_H16 | ops.FLAG_NO_F32, bias=bqkv)
        ctx, _, _ = ops.mha_fwd(qkv, a.h, kv_len, split_out=split, dropout_p=pa, dropout_seed=_seed(pa))
        x1, _, _ = ops.gemm(ctx, wo, epi=ops.EPI_RESID, flags=fin, bias=a.linear_out.bias, resid=x, gamma=self._ones, dropout_p=pd,
                            dropout_seed=_seed(pd))
        _, xn2 = ops.layernorm(x1, self.norm2.weight, self.norm2.bias, self.norm2.eps, f32=False, h16=True, split=split)
        h, _, _ = ops.gemm(xn2, w1, epi=ops.EPI_RELU, flags=fin | fout, bias=f.w_1.bias, dropout_p=pf, dropout_seed=_seed(pf))
        x2, _, _ = ops.gemm(h, w2, epi=ops.EPI_RESID, flags=fin, bias=f.w_2.bias, resid=x1, gamma=self._ones, dropout_p=pd,
                            dropout_seed=_seed(pd))
        return x2

    def forward_train(self, x: torch.Tensor, kv_len: torch.Tensor) -> torch.Tensor:
        from ....autograd import TransformerLayerFn

        a, f = self.self_attn, self.feed_forward
        pd = self.dropout_rate if self.training else 0.0
        pa = a.dropout_rate if self.training else 0.0
        pf = f.dropout_rate if self.training else 0.0
        return TransformerLayerFn.apply(x, kv_len, a.h, pd, pa, pf, _seed(max(pd, pa, pf)), self._ones,
                                        self.norm1.weight, self.norm1.bias, a.linear_q.weight, a.linear_q.bias, a.linear_k.weight,
                                        a.linear_k.bias, a.linear_v.weight, a.linear_v.bias, a.linear_out.weight, a.linear_out.bias,
                                        self.norm2.weight, self.norm2.bias, f.w_1.weight, f.w_1.bias, f.w_2.weight, f.w_2.bias)


class Encoder(nn.Module):
    """The subset of modules/_transformer/encoder.py:24-313 that configs/model/**/transformer.yaml instantiates."""

    def __init__(self, idim, attention_dim=256, attention_heads=4, linear_units=2048, num_blocks=6, dropout_rate=0.1,
                 positional_dropout_rate=0.1, attention_dropout_rate=0.0, input_layer="conv2d", pos_enc_class=PositionalEncoding,
                 normalize_before=True, concat_after=False, positionwise_layer_type="linear", positionwise_conv_kernel_size=1,
                 selfattention_layer_type="selfattn", stochastic_depth_rate=0.0, **unsupported):
        super().__init__()
        if input_layer is not None:
            raise NotImplementedError("input_layer must be None (the Transformer wrapper forces it, modules/transformer.py:18)")
        if not normalize_before or concat_after or positionwise_layer_type != "conv1d" or selfattention_layer_type != "selfattn" \
                or stochastic_depth_rate != 0.0:
            raise NotImplementedError("only the pre-LN / conv1d position-wise / selfattn encoder of configs/model/transformer.yaml is built")
        if pos_enc_class is not ScaledPositionalEncoding:
            raise NotImplementedError("use_scaled_pos_enc must be true (configs/model/transformer.yaml)")
        self.embed = nn.Sequential(pos_enc_class(attention_dim, positional_dropout_rate))
        self.normalize_before = normalize_before
        self.encoders = nn.ModuleList([
            EncoderLayer(attention_dim, MultiHeadedAttention(attention_heads, attention_dim, attention_dropout_rate),
                         MultiLayeredConv1d(attention_dim, linear_units, positionwise_conv_kernel_size, dropout_rate), dropout_rate)
            for _ in range(num_blocks)
        ])
        self.after_norm = LayerNorm(attention_dim)


def initialize(model: nn.Module, init: str):
    """modules/_transformer/initialize.py:65-90 (non-chainer branch)."""
    fns = {"xavier_uniform": nn.init.xavier_uniform_, "xavier_normal": nn.init.xavier_normal_,
           "kaiming_uniform": lambda p: nn.init.kaiming_uniform_(p, nonlinearity="relu"),
           "kaiming_normal": lambda p: nn.init.kaiming_normal_(p, nonlinearity="relu")}
    if init not in fns:
        raise ValueError("Unknown initialization: " + init)
    for p in model.parameters():
        if p.dim() > 1:
            fns[init](p.data)
    for name, p in model.named_parameters():
        if ".bias" in name and p.dim() == 1:
            p.data.zero_()
    for m in model.modules():
        if isinstance(m, (nn.Embedding, nn.LayerNorm, nn.GroupNorm)):
            m.reset_parameters()


class Transformer(nn.Module):
    """Wraps the espnet transformer encoder (reference modules/transformer.py:9-27)."""

    def __init__(self, dim, **kwargs):
        super().__init__()
        use_scaled_pos_enc = kwargs.pop("use_scaled_pos_enc")
        init_alpha = kwargs.pop("init_alpha")
        init_type = kwargs.pop("init_type")
        pos_enc_class = ScaledPositionalEncoding if use_scaled_pos_enc else PositionalEncoding
        kwargs.update(dict(idim=0, attention_dim=dim, input_layer=None, pos_enc_class=pos_enc_class))
        self.transformer = Encoder(**kwargs)
        initialize(self, init_type)
        if use_scaled_pos_enc:
            self.transformer.embed[-1].alpha.data = torch.tensor(init_alpha)

    def forward(self, x: torch.Tensor, padding_mask: torch.Tensor, want_h16: bool = False, split: Optional[bool] = None):
        """x (B,T,C) fp32, padding_mask (B,T) bool True = pad (a prefix mask built from lengths, as every caller passes)
        -> (B,T,C) fp32 [, fp16 operand copy].  Keys at padded positions are masked; nothing else is (reference semantics)."""
        enc = self.transformer
        x = x.contiguous()
        kv_len = (~padding_mask).sum(dim=1).to(torch.int64).contiguous()
        ln = enc.after_norm
        if torch.is_grad_enabled():
            from ....autograd import LayerNormFn

            x = enc.embed[0](x)
            for layer in enc.encoders:
                x = layer.forward_train(x, kv_len)
            out = LayerNormFn.apply(x, ln.weight, ln.bias, ln.eps)
            return (out, None) if want_h16 else out
        split = precision.use_split(self.training) if split is None else split
        x = enc.embed[0](x)
        for layer in enc.encoders:
            x = layer.forward_cl(x, kv_len, split)
        o32, o16 = ops.layernorm(x, ln.weight, ln.bias, ln.eps, f32=True, h16=want_h16, split=split)
        return (o32, o16) if want_h16 else o32
