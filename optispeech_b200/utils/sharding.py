"""Synthesis over several GPUs: replicas only.

Utterances are independent through the whole synthesis path, so a multi-GPU synthesis job shards the UTTERANCES — sorted by
length, dealt round-robin, so that every rank gets the same mix of long and short ones — and every rank runs
`OptiSpeech.synthesise` / `OptiSpeechGenerator.synthesise` on its share with no collective on the data path.  The only
exchange is the optional gather of the finished waveforms on the host (`gather_outputs`: `all_gather_object`, any backend).
The reference has no multi-GPU synthesis; its `infer.py` loops over sentences on one device (optispeech/infer.py:38).
"""
from __future__ import annotations

from typing import Any, Dict, List, Sequence

import torch


def shard_by_length(lengths: Sequence[int], world: int) -> List[List[int]]:
    """-> per-rank lists of utterance indices.  Longest first, dealt round-robin in a snake order (0..w-1, w-1..0, ...) so that
    the total number of phonemes per rank differs by at most one utterance's length."""
    if world <= 0:
        raise ValueError("world must be positive")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    shards: List[List[int]] = [[] for _ in range(world)]
    for pos, idx in enumerate(order):
        lap, k = divmod(pos, world)
        shards[k if lap % 2 == 0 else world - 1 - k].append(idx)
    return shards


def synthesise_sharded(model, x: torch.Tensor, x_lengths: torch.Tensor, rank: int, world: int, max_batch: int = 8, **kwargs) -> Dict[int, Dict[str, Any]]:
    """Runs `model.synthesise` (an OptiSpeechGenerator) on this rank's share of the utterances `x` (N, Tmax) int64 ids /
    `x_lengths` (N,), in batches of at most `max_batch` utterances of similar length.  -> {utterance index: {"wav": 1-D CPU tensor
    cut to its length, "durations": ...}}.  No communication."""
    mine = shard_by_length([int(v) for v in x_lengths], world)[rank]
    dev = next(model.parameters()).device
    out: Dict[int, Dict[str, Any]] = {}
    for s in range(0, len(mine), max_batch):
        idx = mine[s:s + max_batch]
        lens = x_lengths[idx]
        tmax = int(lens.max())
        ids = x[idx, :tmax].contiguous()
        res = model.synthesise(ids.to(dev, non_blocking=True), lens, **kwargs)
        for j, i in enumerate(idx):
            n = int(res["wav_lengths"][j])
            out[i] = {"wav": res["wav"][j, :n].clone(), "durations": res["durations"][j, : int(lens[j])].clone()}
    return out


def gather_outputs(local: Dict[int, Any], world: int) -> Dict[int, Any]:
    """Host-side gather of the per-rank result dictionaries (every rank receives all of them).  world == 1: returns `local`."""
    if world == 1 or not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return dict(local)
    parts: List[Any] = [None] * world
    torch.distributed.all_gather_object(parts, local)
    merged: Dict[int, Any] = {}
    for p in parts:
        merged.update(p)
    return merged
