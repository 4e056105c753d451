"""Data path on the host: dataset reader, the collate function's padding / normalisation semantics against
the oracle restatement (reference text_wav_datamodule.py:133-266), the prefetcher, and the Slaney mel basis against an
independent implementation."""
import json

import numpy as np
import pytest
import torch

from oracle import data as O

STATS = dict(mel_mean=-5.5, mel_std=2.1, energy_mean=20.0, energy_std=15.0, pitch_mean=180.0, pitch_std=60.0)


def _make_corpus(root, n=7, n_feats=20, hop=256, seed=0, some_sids=True):
    rng = np.random.default_rng(seed)
    data = root / "data"
    data.mkdir()
    stems = []
    for i in range(n):
        frames = int(rng.integers(5, 40))
        stem = f"utt{i:03d}"
        wav = (rng.standard_normal(frames * hop) * 0.6).astype(np.float32)          # some samples beyond [-1, 1]
        pitch = rng.uniform(-20, 300, frames).astype(np.float32)                    # some at / below the voicing threshold
        np.savez(data / f"{stem}.npz", wav=wav, mel=rng.standard_normal((n_feats, frames)).astype(np.float32),
                 energy=rng.uniform(0, 50, frames).astype(np.float64), pitch=pitch)
        meta = dict(phoneme_ids=[int(v) for v in rng.integers(1, 150, int(rng.integers(3, 30)))], text=f"text {i}")
        if some_sids and i % 2 == 0:
            meta["sid"] = 1
        (data / f"{stem}.json").write_text(json.dumps(meta), encoding="utf-8")
        stems.append(stem)
    (root / "train.txt").write_text("\n".join(stems) + "\n\n", encoding="utf-8")
    return stems


def test_dataset_reads_reference_layout(tmp_path):
    from optispeech_b200.dataset import TextWavDataset, parse_filelist

    stems = _make_corpus(tmp_path)
    assert parse_filelist(tmp_path / "train.txt") == stems           # blank lines dropped
    ds = TextWavDataset(tmp_path / "train.txt", uv_threshold=10.0, seed=3)
    assert len(ds) == len(stems) and sorted(ds.file_paths) == stems and ds.file_paths != stems   # seeded shuffle
    item = ds[0]
    raw = np.load(tmp_path / "data" / f"{item['filepath']}.npz")
    assert item["x"].dtype == torch.long and item["mel"].shape == raw["mel"].shape
    assert torch.equal(item["wav"], torch.from_numpy(raw["wav"]))
    want_pitch = raw["pitch"].copy()
    want_pitch[want_pitch <= 10.0] = 0.0                              # unvoiced frames zeroed (reference :163-165)
    assert np.array_equal(item["pitch"].numpy(), want_pitch) and (want_pitch == 0).any()
    assert item["sid"] in (1, None) and item["lid"] is None


@pytest.mark.parametrize("do_normalize", [True, False])
def test_collate_matches_reference_semantics(tmp_path, do_normalize):
    from optispeech_b200.dataset import TextWavBatchCollate, TextWavDataset

    _make_corpus(tmp_path, n=5)
    ds = TextWavDataset(tmp_path / "train.txt", seed=0)
    items = [ds[i] for i in range(5)]
    for it in items:
        it["sid"] = None
    got = TextWavBatchCollate(20, STATS, do_normalize=do_normalize)(items)
    want = O.collate(items, 20, STATS, do_normalize=do_normalize)
    for k, v in want.items():
        if isinstance(v, np.ndarray):
            assert isinstance(got[k], np.ndarray) and got[k].dtype == np.float32 and np.array_equal(got[k], v), k
        else:
            assert got[k].dtype == v.dtype and torch.equal(got[k], v), k
    assert got["sids"] is None and got["lids"] is None and len(got["x_texts"]) == 5
    if do_normalize:
        assert np.abs(got["wav"]).max() <= 1.0
        pad = got["mel"][got["mel_lengths"].argmin(), :, -1]        # padded frames hold (0 - mean) / std, not 0
        assert torch.allclose(pad, torch.full_like(pad, (0 - STATS["mel_mean"]) / STATS["mel_std"]))
    # speaker ids: all or none
    items[0]["sid"] = 1
    with pytest.raises(AssertionError):
        TextWavBatchCollate(20, STATS)(items)


def test_prefetcher_preserves_order_and_propagates_errors(tmp_path):
    from optispeech_b200.dataset import BatchPrefetcher, TextWavDataModule

    _make_corpus(tmp_path, n=7, some_sids=False)
    (tmp_path / "valid.txt").write_text((tmp_path / "train.txt").read_text())
    dm = TextWavDataModule(n_feats=20, data_statistics=STATS, train_filelist_path=tmp_path / "train.txt",
                           valid_filelist_path=tmp_path / "valid.txt", batch_size=3, seed=1, pin_memory=False)
    batches = list(dm.val_dataloader())
    assert [b["x"].shape[0] for b in batches] == [3, 3, 1]
    assert [p for b in batches for p in b["filepaths"]] == dm.validset.file_paths
    assert sum(b["x"].shape[0] for b in dm.train_dataloader()) == 7

    def broken():
        yield {"x": torch.zeros(1)}
        raise RuntimeError("reader failed")

    it = iter(BatchPrefetcher(broken(), pin=False))
    next(it)
    with pytest.raises(RuntimeError, match="reader failed"):
        next(it)


@pytest.mark.parametrize("sr,n_fft,n_mels,fmin,fmax", [(22050, 1024, 80, 0, 8000), (22050, 1024, 100, 0, 11025), (24000, 2048, 100, 20, 12000)])
def test_slaney_mel_basis_matches_independent_implementation(sr, n_fft, n_mels, fmin, fmax):
    from transformers.audio_utils import mel_filter_bank

    from optispeech_b200.dataset.feature_extractors import slaney_mel_basis

    ours = slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax)
    ref = mel_filter_bank(n_fft // 2 + 1, n_mels, float(fmin), float(fmax), sr, norm="slaney", mel_scale="slaney").T
    assert ours.shape == ref.shape == (n_mels, n_fft // 2 + 1)
    assert np.abs(ours - ref).max() <= 1e-6 * np.abs(ref).max()
