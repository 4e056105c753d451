"""Mask / normalisation helpers with the reference's names (optispeech/utils/model.py:12-21,74-116,168-217)."""
from __future__ import annotations

import numpy as np
import torch


def sequence_mask(length, max_length=None):
    """(B,) lengths -> (B, max_length) bool, True inside the sequence (utils/model.py:12-16)."""
    if max_length is None:
        max_length = length.max()
    if length.is_cuda and length.dim() == 1:   # one launch (osb_sequence_mask) instead of arange + compare
        from .. import ops

        return ops.sequence_masks(length, int(max_length), valid=True, pad=False)[0]
    x = torch.arange(int(max_length), dtype=length.dtype, device=length.device)
    return x.unsqueeze(0) < length.unsqueeze(1)


def make_non_pad_mask(lengths):
    return sequence_mask(lengths, lengths.max()).unsqueeze(1).bool()


def make_pad_mask(lengths, max_len=None):
    max_length = max_len if max_len is not None else lengths.max()
    if lengths.is_cuda and lengths.dim() == 1:
        from .. import ops

        return ops.sequence_masks(lengths, int(max_length), valid=False, pad=True)[1]
    return ~sequence_mask(lengths, max_length).bool()


def _as_stat(v, data):
    if isinstance(v, (float, int)):
        return v
    if isinstance(v, list):
        v = torch.tensor(v, dtype=data.dtype, device=data.device)
    elif isinstance(v, np.ndarray):
        v = torch.from_numpy(v).to(data.device)
    elif isinstance(v, torch.Tensor):
        v = v.to(data.device)
    return v.unsqueeze(-1)


def normalize(data, mu, std):
    return (data - _as_stat(mu, data)) / _as_stat(std, data)


def denormalize(data, mu, std):
    return data * _as_stat(std, data) + _as_stat(mu, data)


def safe_log(x: torch.Tensor, clip_val: float = 1e-7) -> torch.Tensor:
    return torch.log(torch.clip(x, min=float(clip_val)))


def pad_list(xs, pad_value, max_len=None):
    n_batch = len(xs)
    if max_len is None:
        max_len = max(x.size(0) for x in xs)
    pad = xs[0].new(n_batch, max_len, *xs[0].size()[1:]).fill_(pad_value)
    for i in range(n_batch):
        pad[i, : min(xs[i].size(0), max_len)] = xs[i][:max_len]
    return pad
