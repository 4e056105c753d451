"""GPU feature extraction for the data path (reference optispeech/dataset/feature_extractors/__init__.py:14-200).

`CommonFeatureExtractor` mirrors the reference class's constructor and `get_mel` / `get_energy`, but takes a RAGGED BATCH of
waveforms and runs ONE kernel launch (osb_mel_energy) for log-mel and energy together; the reference loops over utterances on
the CPU (`torch.stft` + `matmul` per file).  Audio decoding, filtering, loudness normalisation, silence trimming and pitch
extraction (third-party CPU algorithms behind `__call__`, :57-110) are outside this package's scope.

The mel basis is librosa's `filters.mel(sr, n_fft, n_mels, fmin, fmax)` default (Slaney scale, Slaney area normalisation),
restated here because librosa is not a dependency; tests pin it to `transformers.audio_utils.mel_filter_bank`.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .. import _lib


def _hz_to_mel_slaney(f: np.ndarray) -> np.ndarray:
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mel = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mel)


def _mel_to_hz_slaney(m: np.ndarray) -> np.ndarray:
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def slaney_mel_basis(sample_rate: int, n_fft: int, n_mels: int, f_min: float, f_max: Optional[float]) -> np.ndarray:
    """(n_mels, n_fft/2+1) float32 triangular filters on the Slaney mel scale, each normalised to unit area in Hz."""
    f_max = float(sample_rate) / 2 if f_max is None else float(f_max)
    fft_freqs = np.linspace(0.0, sample_rate / 2.0, n_fft // 2 + 1)
    mel_pts = _mel_to_hz_slaney(np.linspace(_hz_to_mel_slaney(f_min), _hz_to_mel_slaney(f_max), n_mels + 2))
    fdiff = np.diff(mel_pts)
    ramps = mel_pts[:, None] - fft_freqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0.0, np.minimum(lower, upper))
    weights *= (2.0 / (mel_pts[2:n_mels + 2] - mel_pts[:n_mels]))[:, None]
    return weights.astype(np.float32)


class CommonFeatureExtractor:
    """log-mel ("compatible with most popular neural vocoders", reference :148-200) and energy (:114-146) on the GPU."""

    def __init__(self, sample_rate: int, n_feats: int, n_fft: int, hop_length: int, win_length: int, f_min: int, f_max: int,
                 center: bool = False, pitch_extractor=None, **_unused):
        if center:
            raise ValueError("the reference configs extract features with center=False; centred framing is not implemented")
        self.sample_rate, self.n_feats, self.n_fft, self.hop_length, self.win_length = sample_rate, n_feats, n_fft, hop_length, win_length
        self.f_min, self.f_max, self.center = f_min, f_max, center
        self._dev_tables = {}

    def _tables(self, device):
        key = str(device)
        t = self._dev_tables.get(key)
        if t is None:
            fb = slaney_mel_basis(self.sample_rate, self.n_fft, self.n_feats, self.f_min, self.f_max)
            nz = fb > 0
            klo = np.where(nz.any(1), nz.argmax(1), 0).astype(np.int32)
            khi = np.where(nz.any(1), fb.shape[1] - 1 - nz[:, ::-1].argmax(1), -1).astype(np.int32)
            t = (torch.from_numpy(fb).to(device).contiguous(), torch.from_numpy(klo).to(device), torch.from_numpy(khi).to(device),
                 torch.hann_window(self.win_length, device=device))
            self._dev_tables[key] = t
        return t

    def mel_and_energy(self, wav: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """wav (B, Lmax) fp32 on the GPU, lengths (B) samples -> (mel (B, n_feats, Fmax), energy (B, Fmax), frames (B))."""
        if not wav.is_cuda:
            raise _lib.OsbError("feature extraction runs on the CUDA kernels only (no CPU path)")
        wav = wav.float().contiguous()
        B, Lmax = wav.shape
        lengths = torch.full((B,), Lmax, dtype=torch.int64, device=wav.device) if lengths is None else lengths.to(wav.device, torch.int64)
        Fmax = Lmax // self.hop_length
        fb, klo, khi, window = self._tables(wav.device)
        mel = torch.empty((B, self.n_feats, Fmax), device=wav.device, dtype=torch.float32)
        energy = torch.empty((B, Fmax), device=wav.device, dtype=torch.float32)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.load().osb_mel_energy(wav.data_ptr(), lengths.contiguous().data_ptr(), window.data_ptr(), fb.data_ptr(), klo.data_ptr(),
                                              khi.data_ptr(), mel.data_ptr(), energy.data_ptr(), B, Lmax, Fmax, self.n_feats, self.n_fft,
                                              self.hop_length, self.win_length, 1e-9, 1e-5, stream), "osb_mel_energy")
        return mel, energy, lengths // self.hop_length

    def get_mel(self, wav: Union[np.ndarray, torch.Tensor]) -> torch.Tensor:
        """One utterance, as the reference's signature: (L,) -> (n_feats, frames)."""
        w = torch.as_tensor(wav).reshape(1, -1).cuda()
        return self.mel_and_energy(w)[0][0]

    def get_energy(self, wav: Union[np.ndarray, torch.Tensor], mel_length: Optional[int] = None) -> torch.Tensor:
        w = torch.as_tensor(wav).reshape(1, -1).cuda()
        e = self.mel_and_energy(w)[1][0]
        if mel_length is not None:   # trim_or_pad_to_target_length (utils/model.py:155-165)
            e = e[:mel_length] if e.shape[0] >= mel_length else torch.cat([e, e.new_zeros(mel_length - e.shape[0])])
        return e
