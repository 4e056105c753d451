"""TEST INFRASTRUCTURE — CPU restatement of the reference data path (only tests/ may import this).

collate:   optispeech/dataset/text_wav_datamodule.py:196-266 (`TextWavBatchCollate.__call__`)
features:  optispeech/dataset/feature_extractors/__init__.py:114-146 (`get_energy`), :151-200 (`CommonFeatureExtractor.get_mel`),
           optispeech/utils/audio.py `spectral_normalize_torch` = log(clamp(x, 1e-5))
The mel basis of the reference comes from librosa (`librosa.filters.mel`, Slaney scale + Slaney normalisation), which is not
installed here: the tests take it from `transformers.audio_utils.mel_filter_bank(norm="slaney", mel_scale="slaney")`, an
independent implementation of the same published construction ("parity unpinned" against librosa itself for this one table).
"""
import numpy as np
import torch


def collate(batch, n_feats, stats, do_normalize=True):
    B = len(batch)
    xm = max(i["x"].shape[-1] for i in batch)
    mm = max(i["mel"].shape[-1] for i in batch)
    wm = max(i["wav"].shape[-1] for i in batch)
    x = torch.zeros((B, xm), dtype=torch.long)
    wav = np.zeros((B, wm), dtype=np.float32)
    mel = torch.zeros((B, n_feats, mm))
    pit, ene = torch.zeros((B, mm)), torch.zeros((B, mm))
    for i, it in enumerate(batch):
        x[i, : it["x"].shape[-1]] = it["x"]
        wav[i, : it["wav"].shape[-1]] = it["wav"]
        mel[i, :, : it["mel"].shape[-1]] = it["mel"]
        ene[i, : it["energy"].shape[-1]] = it["energy"].float()
        pit[i, : it["pitch"].shape[-1]] = it["pitch"].float()
    if do_normalize:
        wav = wav.clip(-1, 1)
        mel = (mel - stats["mel_mean"]) / stats["mel_std"]
        ene = (ene - stats["energy_mean"]) / stats["energy_std"]
        pit = (pit - stats["pitch_mean"]) / stats["pitch_std"]
    return dict(x=x, wav=wav, mel=mel, energies=ene, pitches=pit,
                x_lengths=torch.tensor([i["x"].shape[-1] for i in batch]), mel_lengths=torch.tensor([i["mel"].shape[-1] for i in batch]),
                wav_lengths=torch.tensor([i["wav"].shape[-1] for i in batch]))


def mel_and_energy(wav_1d: torch.Tensor, mel_basis: torch.Tensor, n_fft: int, hop: int, win: int):
    """One utterance on the CPU exactly as the reference does it (reflect pad (n_fft-hop)/2, center=False, Hann)."""
    y = wav_1d.reshape(1, 1, -1).float()
    pad = int((n_fft - hop) / 2)
    y = torch.nn.functional.pad(y, (pad, pad), mode="reflect").squeeze(1)
    spec = torch.view_as_real(torch.stft(y, n_fft, hop_length=hop, win_length=win, window=torch.hann_window(win), center=False,
                                         pad_mode="reflect", normalized=False, onesided=True, return_complex=True))
    mag = torch.sqrt(spec.pow(2).sum(-1) + 1e-9)
    mel = torch.log(torch.clamp(torch.matmul(mel_basis, mag), min=1e-5))
    energy = torch.norm(mag, dim=1)
    return mel.squeeze(0), energy.squeeze(0)
