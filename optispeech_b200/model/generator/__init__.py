"""OptiSpeechGenerator on the B200 path.

Mirrors optispeech/model/generator/__init__.py (reference @ 3bdde20): constructor signature (sub-modules
arrive as partials and are called with `dim=`; the vocoder partial with `input_channels, sample_rate,
n_fft, hop_length`; extra **kwargs are swallowed), attribute names (== state_dict prefixes),
`synthesise(...)` and `forward(...)` signatures and result dictionaries.
"""
from __future__ import annotations

import torch
from torch import nn

from ... import ops, precision
from ...utils import sequence_mask
from .alignments import GaussianUpsampling, expand_by_duration


class OptiSpeechGenerator(nn.Module):
    def __init__(
        self,
        dim: int,
        segment_size,
        text_embedding,
        encoder,
        duration_predictor,
        pitch_predictor,
        energy_predictor,
        decoder,
        vocoder,
        loss_coeffs,
        feature_extractor,
        num_speakers,
        num_languages,
        data_statistics,
        **kwargs,
    ):
        super().__init__()
        from .training import AlignmentModule  # local import: training-only machinery

        self.segment_size = segment_size
        self.loss_coeffs = loss_coeffs
        self.n_feats = feature_extractor.n_feats
        self.n_fft = feature_extractor.n_fft
        self.hop_length = feature_extractor.hop_length
        self.sample_rate = feature_extractor.sample_rate
        self.data_statistics = data_statistics
        self.num_speakers = num_speakers
        self.num_languages = num_languages

        self.text_embedding = text_embedding(dim=dim)
        self.encoder = encoder(dim=dim)
        self.duration_predictor = duration_predictor(dim=dim)
        self.alignment_module = AlignmentModule(adim=dim, odim=self.n_feats)
        self.pitch_predictor = pitch_predictor(dim=dim)
        self.energy_predictor = energy_predictor(dim=dim)
        self.feature_upsampler = GaussianUpsampling()
        self.decoder = decoder(dim=dim)
        self.vocoder = vocoder(input_channels=dim, sample_rate=self.sample_rate, n_fft=self.n_fft, hop_length=self.hop_length)
        if self.num_speakers > 1:
            self.sid_embed = torch.nn.Embedding(self.num_speakers, dim)
        if self.num_languages > 1:
            self.lid_embed = torch.nn.Embedding(self.num_languages, dim)

    # ------------------------------------------------------------------------------------------
    def _speaker_language(self, x, sids, lids):
        if sids is not None:
            x = x + self.sid_embed(sids.view(-1)).unsqueeze(1)
        if lids is not None:
            x = x + self.lid_embed(lids.view(-1)).unsqueeze(1)
        return x

    def forward(self, x, x_lengths, mel, mel_lengths, pitches, energies, sids, lids, seg_rand=None):
        """Training forward (reference generator/__init__.py:72-192).  `seg_rand` (optional, (B,) in [0,1)) supplies the draw of
        get_random_segments when the caller made it already (BaseModule.stage_batch cuts the waveform crop on the host)."""
        from .training import generator_training_forward

        return generator_training_forward(self, x, x_lengths, mel, mel_lengths, pitches, energies, sids, lids, seg_rand=seg_rand)

    @torch.inference_mode()
    def synthesise(self, x, x_lengths, sids=None, lids=None, d_factor=1.0, p_factor=1.0, e_factor=1.0, durations=None):
        """Reference generator/__init__.py:194-301.  Returns the same dictionary (CPU tensors + timing scalars).
        `durations` (optional, int64 (B,Tx)) overrides the predicted durations — used by parity tests to hold the
        integer length decisions fixed."""
        dev = x.device
        t0, t1, t2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        t0.record()
        x_lengths = x_lengths.to(dev)
        x_mask = sequence_mask(x_lengths, x.shape[1])
        in_pad = ~x_mask

        split = precision.use_split(False)
        h, _ = self.text_embedding(x)
        h, h16 = self.encoder(h, in_pad, want_h16=True, split=split)

        if (self.num_speakers > 1) and sids is None:
            sids = torch.zeros(x.shape[0], dtype=torch.long, device=dev)
        if (self.num_languages > 1) and lids is None:
            lids = torch.zeros(x.shape[0], dtype=torch.long, device=dev)
        if sids is not None or lids is not None:
            h = self._speaker_language(h, sids, lids).contiguous()
            h16 = ops.to_h16(h, split=split)

        d_pred, y_lengths = self.duration_predictor.infer(h, in_pad, factor=d_factor, x_h16=h16)
        if durations is None:
            durations = d_pred
        else:
            durations = durations.to(dev).to(torch.int64).contiguous()
            y_lengths = durations.sum(dim=1)

        h, pitch, h16 = self.pitch_predictor.infer(h, in_pad, p_factor, x_h16=h16, want_h16=True)
        if self.energy_predictor is not None:
            h, energy = self.energy_predictor.infer(h, in_pad, e_factor, x_h16=h16)
        else:
            energy = None

        y_max_length = int(y_lengths.max().item())  # the one unavoidable device->host read: output length
        if y_max_length == 0:
            # reference alignments.py:152-157: all-zero durations are patched to one frame per row
            durations = GaussianUpsampling.patch_all_zero(durations.clone())
            y_lengths = durations.sum(dim=1)
            y_max_length = int(y_lengths.max().item())
        y_mask = sequence_mask(y_lengths, y_max_length)
        tgt_pad = ~y_mask

        y = self.feature_upsampler(hs=h, ds=durations, h_masks=y_mask, d_masks=x_mask, x_lengths=x_lengths, y_lengths=y_lengths)
        y, y16 = self.decoder(y, tgt_pad, want_h16=True, split=split)
        t1.record()

        f0_cond, _ = expand_by_duration(pitch.unsqueeze(-1), durations, max_len=y_max_length)
        wav = self.vocoder.forward_cl(y16, tgt_pad, split)
        wav_lengths = y_lengths * self.hop_length
        t2.record()

        out = {
            "wav": wav.detach().cpu(),
            "wav_lengths": wav_lengths.detach().cpu(),
            "durations": durations.detach().cpu(),
            "pitch": pitch.detach().cpu(),
            "energy": energy.detach().cpu() if energy is not None else None,
        }
        t2.synchronize()
        am_infer, v_infer = t0.elapsed_time(t1), t1.elapsed_time(t2)
        wav_t = wav.shape[-1] / (self.sample_rate * 1e-3)
        out.update(am_rtf=am_infer / wav_t, v_rtf=v_infer / wav_t, rtf=(am_infer + v_infer) / wav_t, latency=am_infer + v_infer)
        out["_device"] = {"wav": wav, "decoder_out": y, "f0_cond": f0_cond, "y_lengths": y_lengths}
        return out
