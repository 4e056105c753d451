"""Random segment selection with the reference's names (optispeech/utils/segments.py:12-72).

Start indices are drawn on the CPU generator exactly like the reference (`torch.rand([B])` on the default
CPU generator, :32) so a seeded run picks the same crops; the crop itself is one batched gather on the
device instead of a per-sample Python loop.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch


def get_segments(x: torch.Tensor, start_idxs: torch.Tensor, segment_size: int) -> torch.Tensor:
    """x (B, C, T) -> (B, C, segment_size); positions past T read as zero."""
    b, c, t = x.shape
    pos = start_idxs.to(x.device).view(b, 1) + torch.arange(segment_size, device=x.device).view(1, -1)
    valid = pos < t
    gathered = torch.gather(x, 2, pos.clamp(max=t - 1).view(b, 1, -1).expand(b, c, segment_size))
    return gathered * valid.view(b, 1, -1).to(x.dtype)


def get_random_segments(x: torch.Tensor, x_lengths: torch.Tensor, segment_size: int) -> Tuple[torch.Tensor, torch.Tensor]:
    batches = x.shape[0]
    max_start_idx = (x_lengths - segment_size).clamp(min=0)
    start_idxs = (torch.rand([batches]).to(x.device) * max_start_idx).to(dtype=torch.long)
    return get_segments(x, start_idxs, segment_size), start_idxs


def get_segments_numpy(x: np.ndarray, start_idxs: np.ndarray, segment_size: int) -> np.ndarray:
    b, c, _ = x.shape
    segments = np.zeros((b, c, segment_size), dtype=np.float32)
    for i, start_idx in enumerate(np.asarray(start_idxs).tolist()):
        segments[i] = x[i, :, start_idx: start_idx + segment_size]
    return segments
