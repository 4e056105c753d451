"""ConvNeXt blocks on the B200 path.

Mirrors optispeech/model/generator/modules/convnext.py (reference @ 3bdde20): same class names,
constructor arguments, parameter names/shapes (`gamma`, `dwconv.*`, `norm.*`, `pwconv1.*`,
`pwconv2.*`, `final_layer_norm.*`) and forward signatures.  The nn.Conv1d / nn.LayerNorm /
nn.Linear children are parameter containers only; compute goes through libosb200:

    dwconv7 + LayerNorm statistics   -> osb_dwconv_ln   (fp16 xhat; LN affine folded into pwconv1)
    pwconv1 + bias + erf-GELU        -> osb_gemm (tcgen05, EPI_GELU)
    pwconv2 + bias, gamma, DropPath, residual, pad mask -> osb_gemm (tcgen05, EPI_RESID)
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .... import ops, precision
from ...packing import PackedCache, pack_linear


class DropPath(nn.Module):
    """Per-sample stochastic depth (reference convnext.py:106-132).  On this path the Bernoulli
    scale is drawn per sample on the device and applied inside the pwconv2 epilogue."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def sample_scale(self, batch: int, device) -> Optional[torch.Tensor]:
        if self.drop_prob == 0.0 or not self.training:
            return None
        pre = self.__dict__.pop("_presampled", None)   # drawn for all blocks of the step at once (presample_drop_paths)
        if pre is not None and pre.shape[0] == batch and pre.device == device:
            return pre
        keep = 1.0 - self.drop_prob
        scale = torch.empty(batch, device=device, dtype=torch.float32).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            scale.div_(keep)
        return scale

    def forward(self, x):
        scale = self.sample_scale(x.shape[0], x.device)
        return x if scale is None else x * scale.view(-1, *([1] * (x.ndim - 1)))

    def extra_repr(self):
        return f"drop_prob={round(self.drop_prob, 3):0.3f}"


def presample_drop_paths(root: nn.Module, batch: int, device) -> None:
    """Draw the per-sample stochastic-depth scales of every DropPath under `root` for one forward with three kernels instead
    of two per block (same distribution: Bernoulli(keep_l) / keep_l per sample and block; reference convnext.py:121-129)."""
    mods = root.__dict__.get("_drop_path_mods")
    if mods is None:
        mods = [m for m in root.modules() if isinstance(m, DropPath) and m.drop_prob > 0.0 and m.scale_by_keep]
        root.__dict__["_drop_path_mods"] = mods
        root.__dict__["_drop_path_keep"] = {}
    active = [m for m in mods if m.training]
    if not active:
        return
    cache = root.__dict__["_drop_path_keep"]
    key = (tuple(id(m) for m in active), str(device))
    keep = cache.get(key)
    if keep is None:
        keep = torch.tensor([1.0 - m.drop_prob for m in active], dtype=torch.float32).to(device).view(-1, 1)
        cache[key] = keep
    scales = torch.bernoulli(keep.expand(len(active), batch)) / keep
    for i, m in enumerate(active):
        m.__dict__["_presampled"] = scales[i]


class ConvNeXtBlock(nn.Module):
    def __init__(self, dim: int, intermediate_dim: int, drop_path: float = 0.0, layer_scale_init_value: float = None):
        super().__init__()
        self.dim = dim
        self.intermediate_dim = intermediate_dim
        self.dwconv = nn.Conv1d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, intermediate_dim)
        self.act = nn.GELU()
        self.pwconv2 = nn.Linear(intermediate_dim, dim)
        self.gamma = (
            nn.Parameter(layer_scale_init_value * torch.ones(dim), requires_grad=True)
            if layer_scale_init_value > 0
            else None
        )
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self._packed = PackedCache()

    # -- packed operands -------------------------------------------------------------------
    def packed(self):
        srcs = [self.norm.weight, self.norm.bias, self.pwconv1.weight, self.pwconv1.bias, self.pwconv2.weight]

        def build():
            w1 = pack_linear(self.pwconv1.weight, col_scale=self.norm.weight.detach().contiguous())
            b1 = (self.pwconv1.bias + self.pwconv1.weight @ self.norm.bias).contiguous()
            w2 = pack_linear(self.pwconv2.weight)
            return w1, b1, w2

        return self._packed.get("fwd", srcs, build)

    def forward_train(self, x: torch.Tensor, pad_mask_u8: Optional[torch.Tensor]) -> torch.Tensor:
        """Autograd path (fp16 operands, activations saved for the hand-written backward)."""
        from ....autograd import ConvNeXtBlockFn

        scale = self.drop_path.sample_scale(x.shape[0], x.device) if isinstance(self.drop_path, DropPath) else None
        gamma = self.gamma if self.gamma is not None else torch.ones(self.dim, device=x.device)
        return ConvNeXtBlockFn.apply(x, self.dwconv.weight, self.dwconv.bias, self.norm.weight, self.norm.bias, self.pwconv1.weight,
                                     self.pwconv1.bias, self.pwconv2.weight, self.pwconv2.bias, gamma, pad_mask_u8, scale, self.norm.eps)

    def forward_cl(self, x: torch.Tensor, pad_mask_u8: Optional[torch.Tensor], split: bool = False) -> torch.Tensor:
        """Channels-last forward: x (B,T,C) fp32 -> (B,T,C) fp32, pad mask (B,T) uint8 applied to the output
        (the reference's ConvNeXtBackbone multiplies by the mask right after each block, convnext.py:98-101)."""
        w1, b1, w2 = self.packed()
        if not split and (self.dim, self.intermediate_dim) in ops.FUSED_BLOCK_SHAPES:
            # single-pass fp16 operands: the whole block is ONE kernel (osb_convnext.cu)
            scale = self.drop_path.sample_scale(x.shape[0], x.device) if isinstance(self.drop_path, DropPath) else None
            gamma = self.gamma if self.gamma is not None else torch.ones(self.dim, device=x.device)
            return ops.convnext_block_fwd(x, self.dwconv.weight.view(self.dim, 7), self.dwconv.bias, w1[0, 0], b1, w2[0, 0],
                                          self.pwconv2.bias, gamma, scale, pad_mask_u8, self.norm.eps)
        fin = ops.FLAG_SPLIT_IN if split else 0
        xhat, _ = ops.dwconv_ln(x, self.dwconv.weight.view(self.dim, 7), self.dwconv.bias, self.norm.eps, split=split)
        h, _, _ = ops.gemm(xhat, w1, epi=ops.EPI_GELU, bias=b1, flags=fin | (ops.FLAG_SPLIT_OUT if split else 0))
        scale = self.drop_path.sample_scale(x.shape[0], x.device) if isinstance(self.drop_path, DropPath) else None
        gamma = self.gamma if self.gamma is not None else torch.ones(self.dim, device=x.device)
        out, _, _ = ops.gemm(h, w2, epi=ops.EPI_RESID, bias=self.pwconv2.bias, resid=x, gamma=gamma, row_scale=scale,
                             pad_mask=pad_mask_u8, flags=fin | (ops.FLAG_KEEPMASK if pad_mask_u8 is not None else 0))
        return out

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Reference signature: x (B, C, T) -> (B, C, T)."""
        return self.forward_cl(x.transpose(1, 2).contiguous(), None, precision.use_split(self.training)).transpose(1, 2)


class ConvNeXtBackbone(nn.Module):
    """Reference: modules/convnext.py:50-103."""

    def __init__(
        self,
        dim: int,
        intermediate_dim: int,
        num_layers: int,
        drop_path: float = 0.0,
        layer_scale_init_value: Optional[float] = None,
    ):
        super().__init__()
        layer_scale_init_value = layer_scale_init_value or 1 / num_layers
        rates = [r.item() for r in torch.linspace(0, drop_path, num_layers)]
        self.convnext = nn.ModuleList(
            [
                ConvNeXtBlock(dim=dim, intermediate_dim=intermediate_dim, drop_path=r,
                              layer_scale_init_value=layer_scale_init_value)
                for r in rates
            ]
        )
        self.final_layer_norm = nn.LayerNorm(dim, eps=1e-6)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, (nn.Conv1d, nn.Linear)):
            nn.init.trunc_normal_(m.weight, std=0.02)
            nn.init.constant_(m.bias, 0)

    def forward(self, x: torch.Tensor, padding_mask: Optional[torch.Tensor] = None, want_h16: bool = False,
                split: Optional[bool] = None):
        """x (B,T,C) fp32, padding_mask (B,T) bool True = pad -> (B,T,C) fp32 [, fp16 operand copy]."""
        x = x.contiguous()
        mask_u8 = None
        if padding_mask is not None:   # a contiguous bool mask IS the byte mask the kernels read: no cast launch
            mask_u8 = (padding_mask.view(torch.uint8) if padding_mask.dtype == torch.bool and padding_mask.is_contiguous()
                       else padding_mask.to(torch.uint8).contiguous())
        if torch.is_grad_enabled():
            from ....autograd import LayerNormFn

            for blk in self.convnext:
                x = blk.forward_train(x, mask_u8)
            ln = self.final_layer_norm
            out = LayerNormFn.apply(x, ln.weight, ln.bias, ln.eps)
            return (out, None) if want_h16 else out
        split = precision.use_split(self.training) if split is None else split
        for blk in self.convnext:
            x = blk.forward_cl(x, mask_u8, split)
        ln = self.final_layer_norm
        o32, o16 = ops.layernorm(x, ln.weight, ln.bias, ln.eps, f32=True, h16=want_h16, split=split)
        return (o32, o16) if want_h16 else o32
