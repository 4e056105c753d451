"""WaveNeXt vocoder on the B200 path.

Mirrors optispeech/model/vocoder/wavenext/__init__.py (reference @ 3bdde20): `WaveNeXt(input_channels,
dim, intermediate_dim, num_layers, n_fft, hop_length, sample_rate, drop_path, layer_scale_init_value)`
with children `embed` (Conv1d k7), `norm`, `backbone` (ConvNeXtBackbone) and `head` (`linear_1`,
`linear_2` without bias); `forward(x (B,C,T), f0, padding_mask=None) -> (B, T*hop)`; `f0` is ignored
exactly as in the reference (:82-86).

Kernels: embed conv + LayerNorm is one implicit-GEMM launch (7 taps, EPI_BIAS_LN); the head's two
bias-free-composable Linears are folded into one (hop x dim) matrix at packing time (there is no
non-linearity between them, :43-45), so the waveform is written by a single GEMM with the clip fused.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .... import ops, precision
from ...generator.modules import ConvNeXtBackbone
from ...packing import PackedCache, pack_conv, pack_linear


class WaveNeXtHead(nn.Module):
    def __init__(self, dim: int, n_fft: int, hop_length: int):
        super().__init__()
        l_fft = n_fft + 2
        l_shift = hop_length
        self.linear_1 = torch.nn.Linear(dim, l_fft)
        self.linear_2 = torch.nn.Linear(l_fft, l_shift, bias=False)
        nn.init.trunc_normal_(self.linear_1.weight, std=0.02)
        nn.init.trunc_normal_(self.linear_2.weight, std=0.02)
        self._packed = PackedCache()

    def packed(self):
        srcs = [self.linear_1.weight, self.linear_1.bias, self.linear_2.weight]

        def build():
            w = (self.linear_2.weight @ self.linear_1.weight).contiguous()   # (hop, dim)
            b = (self.linear_2.weight @ self.linear_1.bias).contiguous()     # (hop,)
            return pack_linear(w), b

        return self._packed.get("fold", srcs, build)

    def forward_h16(self, x_h16: torch.Tensor, split: bool = False) -> torch.Tensor:
        """x fp16 (B,T,dim) (or split (B,T,2*dim)) -> clipped audio (B, T*hop) fp32."""
        w, b = self.packed()
        out, _, _ = ops.gemm(x_h16, w, epi=ops.EPI_BIAS, flags=ops.FLAG_CLIP | (ops.FLAG_SPLIT_IN if split else 0), bias=b)
        return out.view(out.shape[0], -1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        split = precision.use_split(self.training)
        return self.forward_h16(ops.to_h16(x.contiguous(), split=split), split)


class WaveNeXt(nn.Module):
    def __init__(
        self,
        input_channels: int,
        dim: int,
        intermediate_dim: int,
        num_layers: int,
        n_fft: int,
        hop_length: int,
        sample_rate: int,
        drop_path: float = 0.0,
        layer_scale_init_value: Optional[float] = None,
    ):
        super().__init__()
        self.dim = dim
        self.embed = nn.Conv1d(input_channels, dim, kernel_size=7, padding=3)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.backbone = ConvNeXtBackbone(dim=dim, intermediate_dim=intermediate_dim, num_layers=num_layers,
                                         drop_path=drop_path, layer_scale_init_value=layer_scale_init_value)
        self.head = WaveNeXtHead(dim=dim, n_fft=n_fft, hop_length=hop_length)
        self._packed = PackedCache()

    def forward_cl(self, x_h16: torch.Tensor, padding_mask: Optional[torch.Tensor] = None, split: bool = False) -> torch.Tensor:
        """Channels-last entry: x fp16 (B,T,input_channels) (or split (B,T,2*input_channels)) -> (B, T*hop)."""
        w = self._packed.get("embed", [self.embed.weight], lambda: pack_conv(self.embed.weight))
        h, _, _ = ops.gemm(x_h16, w, epi=ops.EPI_BIAS_LN, flags=ops.FLAG_SPLIT_IN if split else 0, pad=3, bias=self.embed.bias,
                           ln_w=self.norm.weight, ln_b=self.norm.bias, ln_eps=self.norm.eps)
        _, h16 = self.backbone(h, padding_mask, want_h16=True, split=split)
        return self.head.forward_h16(h16, split)

    def forward_train(self, x: torch.Tensor, padding_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Autograd path, channels-last fp32 input (B,T,input_channels) -> (B, T*hop)."""
        if not torch.is_grad_enabled():
            return self.forward_cl(ops.to_h16(x.contiguous()), padding_mask, False)
        from ....autograd import ConvStackFn, LayerNormFn, WaveNeXtHeadFn

        h = ConvStackFn.apply(x, 0, self.embed.weight, self.embed.bias)
        h = LayerNormFn.apply(h, self.norm.weight, self.norm.bias, self.norm.eps)
        h = self.backbone(h, padding_mask)
        return WaveNeXtHeadFn.apply(h, self.head.linear_1.weight, self.head.linear_1.bias, self.head.linear_2.weight)

    def forward(self, x, f0, padding_mask=None):
        """Reference signature: x (B, C, T)."""
        if torch.is_grad_enabled():
            return self.forward_train(x.transpose(1, 2).contiguous(), padding_mask)
        split = precision.use_split(self.training)
        return self.forward_cl(ops.to_h16(x.transpose(1, 2).contiguous(), split=split), padding_mask, split)
