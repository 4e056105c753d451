"""Small layers shared by the generator modules (reference: modules/layers.py:26-71)."""
from __future__ import annotations

import torch
from torch import nn


class LayerNorm(nn.LayerNorm):
    """LayerNorm over an arbitrary dim with eps 1e-12 (reference layers.py:26-45).  Parameter container on
    the B200 path: the normalisation itself runs inside the conv GEMM epilogue (EPI_RELU_LN)."""

    def __init__(self, nout, dim=-1):
        super().__init__(nout, eps=1e-12)
        self.dim = dim


class ScaledSinusoidalEmbedding(nn.Module):
    """scale * [sin(pos * inv_freq) | cos(pos * inv_freq)], inv_freq = theta^-(j/half) (reference layers.py:48-71).
    `inv_freq` is a non-persistent buffer exactly as in the reference; evaluated by osb_embed_text."""

    def __init__(self, dim, theta=10000):
        super().__init__()
        assert (dim % 2) == 0
        self.scale = nn.Parameter(torch.ones(1) * dim**-0.5)
        half_dim = dim // 2
        freq_seq = torch.arange(half_dim).float() / half_dim
        inv_freq = theta ** -freq_seq.float()
        self.register_buffer("inv_freq", inv_freq, persistent=False)
