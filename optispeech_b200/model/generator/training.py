"""Training-only machinery of the generator (alignment learning, losses, training forward).

Mirrors optispeech/model/generator/alignments.py:14-123,177-280 and generator/__init__.py:72-192.
"""
from __future__ import annotations

import torch
from torch import nn


class AlignmentModule(nn.Module):
    """Alignment learning framework (reference alignments.py:14-123): parameter container with the
    reference's layer names; compute lives in `generator_training_forward`."""

    def __init__(self, adim, odim, cache_prior=True):
        super().__init__()
        self.cache_prior = cache_prior
        self._cache = {}
        self.t_conv1 = nn.Conv1d(adim, adim, kernel_size=3, padding=1)
        self.t_conv2 = nn.Conv1d(adim, adim, kernel_size=1, padding=0)
        self.f_conv1 = nn.Conv1d(odim, adim, kernel_size=3, padding=1)
        self.f_conv2 = nn.Conv1d(adim, adim, kernel_size=3, padding=1)
        self.f_conv3 = nn.Conv1d(adim, adim, kernel_size=1, padding=0)


def generator_training_forward(gen, x, x_lengths, mel, mel_lengths, pitches, energies, sids, lids):
    raise NotImplementedError("training forward is being brought up; see DESIGN.md")
