"""Data path of the training loop (reference optispeech/dataset/text_wav_datamodule.py:46-266).

* `TextWavDataset` reads the reference's preprocessed utterances: `<data>/<stem>.json` (`phoneme_ids`, `text`, optional
  `sid` / `lid`) + `<data>/<stem>.npz` (`wav`, `mel`, `energy`, `pitch`), listed in a filelist whose parent directory holds
  `data/` (:133-176).
* `TextWavBatchCollate` has the reference's semantics (:196-266): zero padding to the batch maxima, `wav` as a numpy array
  clipped to [-1, 1], mel / energy / pitch normalised AFTER padding (so padded positions hold `(0 - mean) / std`).
* `BatchPrefetcher` is what replaces the DataLoader's `pin_memory` thread + Lightning's transfer hook: a background thread
  collates into pinned staging buffers and keeps `depth` batches ahead, so `training_step` starts its host -> device copies
  (BaseModule.stage_batch / _upload_early) from page-locked memory without waiting for the collate.
* `feature_extractors.CommonFeatureExtractor` computes log-mel + energy of a ragged batch on the GPU (osb_mel_energy).

Out of scope (CPU-side third-party algorithms of the preprocessing CLI): audio decoding / resampling, silence trimming,
loudness normalisation, pitch extraction, the text front-end.
"""
from __future__ import annotations

import json
import queue
import random
import threading
from pathlib import Path
from typing import Dict, Iterable, Iterator, List, Optional

import numpy as np
import torch

from ..utils.model import normalize


def parse_filelist(filelist_path) -> List[str]:
    """One utterance stem per line, blank lines dropped (reference :46-49)."""
    lines = Path(filelist_path).read_text(encoding="utf-8").splitlines()
    return [ln.strip() for ln in lines if ln.strip()]


class TextWavDataset(torch.utils.data.Dataset):
    """Reference `TextWavDataset` (:133-194) without the on-the-fly preprocessing hooks."""

    def __init__(self, filelist_path, uv_threshold: float = 0.0, seed: Optional[int] = None, **_unused):
        self.file_paths = parse_filelist(filelist_path)
        self.data_dir = Path(filelist_path).parent.joinpath("data")
        self.uv_threshold = uv_threshold
        random.seed(seed)
        random.shuffle(self.file_paths)

    def get_datapoint(self, filepath) -> Dict:
        stem = Path(filepath)
        if not stem.is_absolute():
            stem = self.data_dir / stem
        with open(stem.with_suffix(".json"), encoding="utf-8") as fh:
            meta = json.load(fh)
        arrays = np.load(stem.with_suffix(".npz"), allow_pickle=False)
        pitch = torch.from_numpy(arrays["pitch"])
        pitch[pitch <= self.uv_threshold] = 0.0
        return dict(x=torch.LongTensor(meta["phoneme_ids"]), wav=torch.from_numpy(arrays["wav"]), mel=torch.from_numpy(arrays["mel"]),
                    energy=torch.from_numpy(arrays["energy"]), pitch=pitch, sid=meta.get("sid"), lid=meta.get("lid"), text=meta["text"],
                    filepath=str(filepath))

    def __getitem__(self, index):
        return self.get_datapoint(self.file_paths[index])

    def __len__(self):
        return len(self.file_paths)


class TextWavBatchCollate:
    """Reference `TextWavBatchCollate.__call__` (:196-266), same keys, dtypes and padding values."""

    def __init__(self, n_feats: int, data_statistics: Dict[str, float], do_normalize: bool = True):
        self.n_feats, self.data_statistics, self.do_normalize = n_feats, data_statistics, do_normalize

    def __call__(self, batch):
        B = len(batch)
        x_max = max(item["x"].shape[-1] for item in batch)
        mel_max = max(item["mel"].shape[-1] for item in batch)
        wav_max = max(item["wav"].shape[-1] for item in batch)
        x = torch.zeros((B, x_max), dtype=torch.long)
        wav = np.zeros((B, wav_max), dtype=np.float32)
        mel = torch.zeros((B, self.n_feats, mel_max), dtype=torch.float32)
        pitches = torch.zeros((B, mel_max), dtype=torch.float32)
        energies = torch.zeros((B, mel_max), dtype=torch.float32)
        x_lengths, wav_lengths, mel_lengths, sids, lids, texts, paths = [], [], [], [], [], [], []
        for i, item in enumerate(batch):
            nx, nw, nm = item["x"].shape[-1], item["wav"].shape[-1], item["mel"].shape[-1]
            x_lengths.append(nx); wav_lengths.append(nw); mel_lengths.append(nm)
            x[i, :nx] = item["x"]
            wav[i, :nw] = item["wav"]
            mel[i, :, :nm] = item["mel"]
            energies[i, : item["energy"].shape[-1]] = item["energy"].float()
            pitches[i, : item["pitch"].shape[-1]] = item["pitch"].float()
            if item["sid"] is not None:
                sids.append(item["sid"])
            if item["lid"] is not None:
                lids.append(item["lid"])
            texts.append(item["text"]); paths.append(item["filepath"])
        sids = torch.LongTensor(sids) if sids else None
        lids = torch.LongTensor(lids) if lids else None
        if sids is not None:
            assert sids.shape[0] == B, "Not all speaker IDs are provided"
        if lids is not None:
            assert lids.shape[0] == B, "Not all language IDs are provided"
        if self.do_normalize:
            st = self.data_statistics
            wav = wav.clip(-1, 1)
            mel = normalize(mel, st["mel_mean"], st["mel_std"])
            energies = normalize(energies, st["energy_mean"], st["energy_std"])
            pitches = normalize(pitches, st["pitch_mean"], st["pitch_std"])
        return dict(x=x, wav=wav, mel=mel, x_lengths=torch.tensor(x_lengths, dtype=torch.long),
                    wav_lengths=torch.tensor(wav_lengths, dtype=torch.long), mel_lengths=torch.tensor(mel_lengths, dtype=torch.long),
                    energies=energies, pitches=pitches, sids=sids, lids=lids, x_texts=texts, filepaths=paths)


def pin_batch(batch: Dict) -> Dict:
    """Page-locked copies of a collated batch's tensors (numpy `wav` stays numpy: the crop is cut on the host, stage_batch)."""
    if not torch.cuda.is_available():
        return batch
    return {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}


class BatchPrefetcher:
    """Iterate over `loader` (any iterable of collated batches) from a background thread, `depth` pinned batches ahead."""

    _END = object()

    def __init__(self, loader: Iterable[Dict], depth: int = 2, pin: bool = True):
        self.loader, self.depth, self.pin = loader, max(1, depth), pin

    def __iter__(self) -> Iterator[Dict]:
        q: "queue.Queue" = queue.Queue(maxsize=self.depth)
        stop = threading.Event()

        def work():
            try:
                for batch in self.loader:
                    if stop.is_set():
                        return
                    q.put(pin_batch(batch) if self.pin else batch)
                q.put(self._END)
            except BaseException as exc:  # surfaced in the consumer
                q.put(exc)

        t = threading.Thread(target=work, daemon=True)
        t.start()
        try:
            while True:
                item = q.get()
                if item is self._END:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            stop.set()
            while t.is_alive():   # unblock a producer waiting on a full queue
                try:
                    q.get_nowait()
                except queue.Empty:
                    t.join(timeout=0.05)


class TextWavDataModule:
    """The reference `TextWavDataModule`'s loaders (:52-131) without Lightning: `train_dataloader()` / `val_dataloader()`
    return prefetching iterables of collated batches."""

    def __init__(self, *, n_feats: int, data_statistics: Dict[str, float], train_filelist_path, valid_filelist_path, batch_size: int,
                 num_workers: int = 0, pin_memory: bool = True, seed: Optional[int] = None, uv_threshold: float = 0.0, prefetch: int = 2,
                 **_unused):
        self.n_feats, self.data_statistics = n_feats, data_statistics
        self.train_filelist_path, self.valid_filelist_path = train_filelist_path, valid_filelist_path
        self.batch_size, self.num_workers, self.pin_memory, self.seed = batch_size, num_workers, pin_memory, seed
        self.uv_threshold, self.prefetch = uv_threshold, prefetch
        self.trainset = self.validset = None

    def setup(self, stage: Optional[str] = None):
        self.trainset = TextWavDataset(self.train_filelist_path, uv_threshold=self.uv_threshold, seed=self.seed)
        self.validset = TextWavDataset(self.valid_filelist_path, uv_threshold=self.uv_threshold, seed=self.seed)

    def _loader(self, ds, shuffle: bool, do_normalize: bool = True):
        dl = torch.utils.data.DataLoader(ds, batch_size=self.batch_size, num_workers=self.num_workers, shuffle=shuffle,
                                         collate_fn=TextWavBatchCollate(self.n_feats, self.data_statistics, do_normalize=do_normalize))
        return BatchPrefetcher(dl, depth=self.prefetch, pin=self.pin_memory)

    def train_dataloader(self, do_normalize: bool = True):
        if self.trainset is None:
            self.setup()
        return self._loader(self.trainset, shuffle=True, do_normalize=do_normalize)

    def val_dataloader(self):
        if self.validset is None:
            self.setup()
        return self._loader(self.validset, shuffle=False)
