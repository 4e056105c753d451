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

    # ------------------------------------------------------------------------------------------
    # synthesis.  Stages around the one unavoidable device -> host read (the output length): A = embedding, encoder,
    # duration / pitch / energy predictors; B = upsampler, decoder; C = vocoder.  Each stage is ~30-100 small launches, so for a
    # repeated input SHAPE (B, Tx[, Tm]) the stage is captured into a CUDA graph the second time the shape is seen and
    # replayed afterwards (`synthesis_graphs`; exact shapes only: padding Tx or Tm to a bucket would change what the
    # unmasked convolution inputs at the sequence ends see, i.e. the result).  Graphs are dropped when any weight changes.
    # ------------------------------------------------------------------------------------------
    synthesis_graphs = True    # measured on B200, B=1 x 120 phonemes 1.91 -> 1.81 ms, B=8 x 512 2.59 -> 2.39 ms device time
    #                            (the stages are bound by the in-kernel latency of ~120 dependent small kernels, not by launches)
    _SYNTH_MAX_GRAPHS = 16

    def _synth_stage_a(self, x, x_lengths, sids, lids, d_factor, p_factor, e_factor):
        dev = x.device
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
        # the duration stack reads the encoder output only: on a side stream next to the pitch / energy stacks (in a captured
        # stage the fork and join are graph edges).  A B=1 utterance is ONE 128-row tile per predictor layer.
        main = torch.cuda.current_stream(dev)
        side = ops.side_stream(dev, 2)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            d_pred, y_lengths = self.duration_predictor.infer(h, in_pad, factor=d_factor, x_h16=h16)
            # the output length is the one value the host has to read before it can size the next stage: it leaves through
            # page-locked memory as part of this stage (no separate reduction launch + blocking read behind it)
            y_max_pin = self._y_max_pin(dev)
            y_max_pin.copy_(y_lengths.max().reshape(1), non_blocking=True)
        h, pitch, h16 = self.pitch_predictor.infer(h, in_pad, p_factor, x_h16=h16, want_h16=True)
        if self.energy_predictor is not None:
            h, energy = self.energy_predictor.infer(h, in_pad, e_factor, x_h16=h16)
        else:
            energy = None
        main.wait_stream(side)
        return {"h": h, "d_pred": d_pred, "y_lengths": y_lengths, "pitch": pitch, "energy": energy, "x_mask": x_mask,
                "y_max_pin": y_max_pin}

    def _y_max_pin(self, dev) -> torch.Tensor:
        """A page-locked int64 scalar per call site of stage A (a captured stage keeps writing to the one it was captured with)."""
        return torch.zeros(1, dtype=torch.int64, pin_memory=True)

    def _synth_stage_b(self, h, durations, x_mask, x_lengths, y_lengths, y_max_length: int):
        split = precision.use_split(False)
        y_mask = sequence_mask(y_lengths, y_max_length)
        tgt_pad = ~y_mask
        y = self.feature_upsampler(hs=h, ds=durations, h_masks=y_mask, d_masks=x_mask, x_lengths=x_lengths, y_lengths=y_lengths)
        y, y16 = self.decoder(y, tgt_pad, want_h16=True, split=split)
        return {"y": y, "y16": y16, "tgt_pad": tgt_pad}

    def _synth_stage_c(self, y16, tgt_pad, pitch, durations, y_max_length: int):
        f0_cond, _ = expand_by_duration(pitch.unsqueeze(-1), durations, max_len=y_max_length)
        wav = self.vocoder.forward_cl(y16, tgt_pad, precision.use_split(False))
        return {"wav": wav, "f0_cond": f0_cond}

    def _weights_token(self) -> int:
        """Changes whenever a parameter is modified in place (optimizer step, load_state_dict).  The parameter list is cached:
        walking the module tree costs ~0.2 ms per call, more than a B=1 stage."""
        ps = self.__dict__.get("_synth_params")
        if ps is None:
            ps = self.__dict__["_synth_params"] = list(self.parameters())
        return sum([p._version for p in ps])

    def _graphed_stage(self, key, fn, inputs):
        """Run `fn(**inputs)` through the graph cache: eager the first time a key is seen, captured the second time, replayed
        afterwards (inputs are copied into the capture's static tensors)."""
        cache = self.__dict__.setdefault("_synth_cache", {"token": None, "seen": {}, "graphs": {}})
        entry = cache["graphs"].get(key)
        if entry is None:
            if not cache["seen"].get(key):
                if len(cache["seen"]) > 4096:
                    cache["seen"].clear()
                cache["seen"][key] = True
                return fn(**inputs)
            if cache["seen"][key] == "eager":     # a capture of this shape failed before: stay on the launch path
                return fn(**inputs)
            # an input that IS the output of an earlier captured stage (fixed address, rewritten by that stage's replay) is
            # read in place; everything else gets a static copy that is refreshed before each replay
            stable = cache.setdefault("stable", set())
            static = {k: (v if (isinstance(v, torch.Tensor) and v.data_ptr() in stable) else (v.clone() if isinstance(v, torch.Tensor) else v))
                      for k, v in inputs.items()}
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    fn(**static)                      # warm-up on the capture stream (allocator, tensor maps, packs)
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    out = fn(**static)
            except RuntimeError:                      # e.g. another thread of the process is capturing: not an error of the request
                torch.cuda.synchronize()
                cache["seen"][key] = "eager"
                return fn(**inputs)
            entry = (graph, static, out)
            while len(cache["graphs"]) >= self._SYNTH_MAX_GRAPHS:
                cache["graphs"].pop(next(iter(cache["graphs"])))
            cache["graphs"][key] = entry
            for t in out.values():
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    stable.add(t.data_ptr())
        graph, static, out = entry
        for k, v in inputs.items():
            if isinstance(v, torch.Tensor) and static[k].data_ptr() != v.data_ptr():
                static[k].copy_(v, non_blocking=True)
        graph.replay()
        return out

    @torch.inference_mode()
    def synthesise(self, x, x_lengths, sids=None, lids=None, d_factor=1.0, p_factor=1.0, e_factor=1.0, durations=None):
        """Reference generator/__init__.py:194-301.  Returns the same dictionary (CPU tensors + timing scalars).
        `durations` (optional, int64 (B,Tx)) overrides the predicted durations — used by parity tests to hold the
        integer length decisions fixed."""
        dev = x.device
        t0, t1, t2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        t0.record()
        x_lengths = x_lengths.to(dev)
        use_graphs = self.synthesis_graphs and x.is_cuda and not self.training
        if use_graphs:   # captured stages hold fp16 packs of the weights: any in-place change of a parameter drops them
            cache = self.__dict__.setdefault("_synth_cache", {"token": None, "seen": {}, "graphs": {}})
            token = self._weights_token()
            if cache["token"] != token:
                cache.update(token=token, seen={}, graphs={}, stable=set())
        a_in = dict(x=x, x_lengths=x_lengths, sids=sids, lids=lids, d_factor=float(d_factor), p_factor=float(p_factor), e_factor=float(e_factor))
        if use_graphs:
            key_a = ("A", tuple(x.shape), sids is not None, lids is not None, float(d_factor), float(p_factor), float(e_factor))
            a = self._graphed_stage(key_a, self._synth_stage_a, a_in)
        else:
            a = self._synth_stage_a(**a_in)
        if durations is None:
            durations, y_lengths = a["d_pred"], a["y_lengths"]
            torch.cuda.current_stream(dev).synchronize()    # the one unavoidable device->host read: output length
            y_max_length = int(a["y_max_pin"])
        elif not durations.is_cuda:
            # durations handed over in host memory: the lengths are host arithmetic, nothing waits for the device
            durations = durations.to(torch.int64).contiguous()
            y_lengths_host = durations.sum(dim=1)
            y_max_length = int(y_lengths_host.max())
            durations = durations.to(dev, non_blocking=True)
            y_lengths = y_lengths_host.to(dev, non_blocking=True)
        else:
            durations = durations.to(torch.int64).contiguous()
            y_lengths = durations.sum(dim=1)
            y_max_length = int(y_lengths.max().item())
        if y_max_length == 0:
            # reference alignments.py:152-157: all-zero durations are patched to one frame per row
            durations = GaussianUpsampling.patch_all_zero(durations.clone())
            y_lengths = durations.sum(dim=1)
            y_max_length = int(y_lengths.max().item())
        b_in = dict(h=a["h"], durations=durations, x_mask=a["x_mask"], x_lengths=x_lengths, y_lengths=y_lengths, y_max_length=y_max_length)
        b = self._graphed_stage(("B", tuple(x.shape), y_max_length), self._synth_stage_b, b_in) if use_graphs else self._synth_stage_b(**b_in)
        t1.record()    # acoustic model done (reference :262), vocoder next
        c_in = dict(y16=b["y16"], tgt_pad=b["tgt_pad"], pitch=a["pitch"], durations=durations, y_max_length=y_max_length)
        c = self._graphed_stage(("C", tuple(x.shape), y_max_length), self._synth_stage_c, c_in) if use_graphs else self._synth_stage_c(**c_in)
        wav, y = c["wav"], b["y"]
        wav_lengths = y_lengths * self.hop_length
        t2.record()

        pitch, energy = a["pitch"], a["energy"]

        def to_host(t):   # device -> page-locked host memory (torch's caching host allocator recycles the blocks), asynchronously
            h_ = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h_.copy_(t.detach(), non_blocking=True)
            return h_

        out = {
            "wav": to_host(wav),
            "wav_lengths": to_host(wav_lengths),
            "durations": to_host(durations),
            "pitch": to_host(pitch),
            "energy": to_host(energy) if energy is not None else None,
        }
        torch.cuda.current_stream().synchronize()   # the copies above (and with them t2) have completed
        t2.synchronize()
        am_infer, v_infer = t0.elapsed_time(t1), t1.elapsed_time(t2)
        wav_t = wav.shape[-1] / (self.sample_rate * 1e-3)
        out.update(am_rtf=am_infer / wav_t, v_rtf=v_infer / wav_t, rtf=(am_infer + v_infer) / wav_t, latency=am_infer + v_infer)
        out["_device"] = {"wav": wav, "decoder_out": y, "f0_cond": c["f0_cond"], "y_lengths": y_lengths}
        return out
