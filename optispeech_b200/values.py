"""Value containers of the public surface (reference: optispeech/values.py:22-171): `InferenceInputs`,
`InferenceOutputs` and the numpy pad / unpad helpers, same field names and conversions."""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass
from typing import Any

import numpy as np
import torch
from torch.nn.utils.rnn import unpad_sequence as torch_unpad_sequence


@dataclass
class BaseValueContainer:
    def as_tuple(self):
        return dataclasses.astuple(self)

    def as_dict(self):
        return {f.name: getattr(self, f.name) for f in dataclasses.fields(self)}

    def _map(self, fn):
        return type(self)(**{k: fn(v) for k, v in self.as_dict().items()})

    def as_torch(self):
        return self._map(lambda v: torch.as_tensor(v) if isinstance(v, (np.ndarray, torch.Tensor)) else v)

    def as_numpy(self):
        def conv(v):
            if isinstance(v, torch.Tensor):
                return v.detach().cpu().numpy()
            return np.asarray(v) if isinstance(v, np.ndarray) else v

        return self._map(conv)

    def to(self, device: str):
        return self._map(lambda v: v.to(device) if isinstance(v, torch.Tensor) else v)


@dataclass(kw_only=True)
class InferenceInputs(BaseValueContainer):
    clean_text: Any
    x: Any
    x_lengths: Any
    sids: Any = None
    lids: Any = None
    d_factor: float = 1.0
    p_factor: float = 1.0
    e_factor: float = 1.0

    @classmethod
    def from_ids_and_lengths(cls, ids, lengths, **kwargs) -> "InferenceInputs":
        x = numpy_pad_sequences(ids).astype(np.int64)
        x_lengths = np.array(lengths, dtype=np.int64)
        return cls(x=x, x_lengths=x_lengths, **kwargs).as_numpy()


@dataclass(kw_only=True)
class InferenceOutputs(BaseValueContainer):
    wav: Any
    wav_lengths: Any
    latency: float
    rtf: float
    durations: Any = None
    pitch: Any = None
    energy: Any = None
    am_rtf: float | None = None
    v_rtf: float | None = None

    def __iter__(self):
        return iter(self.unbatched_wavs())

    def unbatched_wavs(self):
        if isinstance(self.wav, np.ndarray):
            return numpy_unpad_sequences(self.wav, self.wav_lengths)
        if isinstance(self.wav, torch.Tensor):
            return torch_unpad_sequence(self.wav, self.wav_lengths, batch_first=True)
        raise RuntimeError("Unsupported operation")


def numpy_pad_sequences(sequences, maxlen=None, value=0):
    if maxlen is None:
        maxlen = max(len(seq) for seq in sequences)
    padded = np.full((len(sequences), maxlen), value)
    for i, seq in enumerate(sequences):
        padded[i, : len(seq)] = seq
    return padded


def numpy_unpad_sequences(sequences, lengths):
    if not isinstance(lengths, np.ndarray) or len(lengths.shape) != 1:
        raise ValueError("lengths must be a 1D numpy array")
    if np.any(lengths < 0) or np.any(lengths > sequences.shape[-1]):
        raise ValueError("lengths must be between 0 and max_len")
    return [sequences[i, : lengths[i]] for i in range(sequences.shape[0])]
