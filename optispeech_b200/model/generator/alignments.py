"""Length regulation and alignment learning on the B200 path.

Mirrors optispeech/model/generator/alignments.py (reference @ 3bdde20): `GaussianUpsampling`,
`expand_by_duration`, `AlignmentModule`, `viterbi_decode`, `average_by_duration` keep their names
and call signatures.
"""
from __future__ import annotations

import logging

import torch
from torch import nn

from ... import ops


class GaussianUpsampling(torch.nn.Module):
    """Gaussian upsampling with fixed temperature (reference alignments.py:126-174).  The (B,Tm,Tx)
    attention is never written to HBM: osb_gaussian_upsample builds each softmax row in shared memory."""

    def __init__(self, delta=0.1):
        super().__init__()
        self.delta = delta

    def forward(self, hs, ds, h_masks=None, d_masks=None, *, x_lengths=None, y_lengths=None, want_h16=False):
        """hs (B,Tx,C) fp32; ds (B,Tx) int64 or fp32; h_masks (B,Tm) / d_masks (B,Tx) True = valid.
        Masks must be prefix masks (they are `sequence_mask`s in every reference call site); lengths may be
        passed directly to skip the reductions."""
        if x_lengths is None:
            x_lengths = d_masks.sum(dim=1).to(torch.int64) if d_masks is not None else torch.full(
                (hs.shape[0],), hs.shape[1], device=hs.device, dtype=torch.int64)
        if h_masks is None:
            # reference: T_feats = ds.sum() (alignments.py:159-160)
            Tm = int(ds.sum().item())
            y_lengths = torch.full((hs.shape[0],), Tm, device=hs.device, dtype=torch.int64)
        else:
            Tm = h_masks.size(-1)
            if y_lengths is None:
                y_lengths = h_masks.sum(dim=1).to(torch.int64)
        if ds.dtype not in (torch.int64, torch.float32):
            ds = ds.float()
        c, _ = ops.centres(ds.contiguous(), want_csum=False)
        o32, o16 = ops.gaussian_upsample(hs.contiguous(), c, x_lengths.contiguous(), y_lengths.contiguous(), Tm, self.delta,
                                         f32=True, h16=want_h16)
        return (o32, o16) if want_h16 else o32

    def forward_window(self, hs, ds, x_lengths, y_lengths, win_start, Tm: int, W: int, halo: int):
        """Frames win_start[b] - halo .. + W of forward()'s result for every sample (zero outside [0, Tm)): (B, W, C)."""
        if ds.dtype not in (torch.int64, torch.float32):
            ds = ds.float()
        c, _ = ops.centres(ds.contiguous(), want_csum=False)
        return ops.gaussian_upsample_window(hs.contiguous(), c, x_lengths.contiguous(), y_lengths.contiguous(), win_start, Tm, W, halo,
                                            self.delta)

    @staticmethod
    def patch_all_zero(ds):
        """alignments.py:152-157: an all-zero duration batch gets its empty rows set to 1 (with a warning)."""
        if ds.sum() == 0:
            logging.warning("predicted durations includes all 0 sequences. fill the first element with 1.")
            ds[ds.sum(dim=1).eq(0)] = 1
        return ds


def expand_by_duration(x, durations, max_len=None):
    """Hard repeat (reference alignments.py:283-297): x (B,Tx,C), integer durations (B,Tx) ->
    (expanded (B,Tm,C), lengths (B,)).  `max_len` avoids a device->host read when the caller knows it."""
    durations = durations.to(torch.int64).contiguous()
    lengths = durations.sum(dim=1)
    if max_len is None:
        max_len = int(lengths.max().item())
    _, csum = ops.centres(durations, want_csum=True)
    out, _ = ops.expand_gather(x.contiguous().float(), csum, max_len)
    return out.to(x.dtype), lengths


def expand_indices(durations, max_len):
    """Source-token index of every output frame (int32, -1 past the end): the integer form of the length
    regulator used for bit-exact parity checks."""
    durations = durations.to(torch.int64).contiguous()
    _, csum = ops.centres(durations, want_csum=True)
    dummy = torch.zeros((durations.shape[0], durations.shape[1], 1), device=durations.device)
    _, idx = ops.expand_gather(dummy, csum, max_len)
    return idx
