"""Training-only machinery of the generator (alignment learning, losses, training forward).

Mirrors optispeech/model/generator/alignments.py:14-123,177-280, generator/loss.py and
generator/__init__.py:72-192 (reference @ 3bdde20).
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import nn

from ... import ops
from ...autograd import AttnLogProbFn, ConvStackFn, ForwardSumLossFn


class AlignmentModule(nn.Module):
    """Alignment learning framework (reference alignments.py:14-123), same layer names.  The five Conv1d layers
    run as tcgen05 implicit GEMMs (ConvStackFn); the beta-binomial prior is evaluated from a log-factorial table
    (all its arguments are integers), cached per (T, N) like the reference."""

    def __init__(self, adim, odim, cache_prior=True):
        super().__init__()
        self.cache_prior = cache_prior
        self._cache = {}
        self.t_conv1 = nn.Conv1d(adim, adim, kernel_size=3, padding=1)
        self.t_conv2 = nn.Conv1d(adim, adim, kernel_size=1, padding=0)
        self.f_conv1 = nn.Conv1d(odim, adim, kernel_size=3, padding=1)
        self.f_conv2 = nn.Conv1d(adim, adim, kernel_size=3, padding=1)
        self.f_conv3 = nn.Conv1d(adim, adim, kernel_size=1, padding=0)

    def encode_text(self, text):
        """t_conv1 -> ReLU -> t_conv2 (reference alignments.py:55-58)."""
        return ConvStackFn.apply(text, 0, self.t_conv1.weight, self.t_conv1.bias, self.t_conv2.weight, self.t_conv2.bias)

    def encode_feats(self, feats):
        """f_conv1 -> ReLU -> f_conv2 -> ReLU -> f_conv3 (reference alignments.py:60-64); depends on the mel input only."""
        odim = feats.shape[-1]
        return ConvStackFn.apply(feats, ((odim + 63) // 64) * 64, self.f_conv1.weight, self.f_conv1.bias, self.f_conv2.weight,
                                 self.f_conv2.bias, self.f_conv3.weight, self.f_conv3.bias)

    def attend(self, fe, te, text_lengths, feats_lengths, prior=None):
        if prior is None:
            prior = self._generate_prior(text_lengths, feats_lengths, te.shape[1], fe.shape[1])
        # x_masks is the prefix mask of text_lengths in every reference call site (generator/__init__.py:120-126)
        return AttnLogProbFn.apply(fe, te, prior, text_lengths.contiguous(), feats_lengths.contiguous())

    def forward(self, text, feats, text_lengths, feats_lengths, x_masks=None):
        """text (B,Tx,adim), feats (B,Tm,odim) -> log_p_attn (B,Tm,Tx) with the beta-binomial prior added."""
        return self.attend(self.encode_feats(feats), self.encode_text(text), text_lengths, feats_lengths)

    @staticmethod
    def _log_prior(T: int, N: int) -> np.ndarray:
        """(T, N) float64 log BetaBinomial(k; N, t, T-t+1), t = 1..T, k = 0..N-1 (reference :109-114, scipy definition);
        every gamma-function argument is an integer, so a log-factorial table is exact."""
        lf = np.concatenate([[0.0], np.cumsum(np.log(np.arange(1, T + N + 2, dtype=np.float64)))])  # lf[m] = log(m!)
        t = np.arange(1, T + 1)[:, None]
        k = np.arange(N)[None, :]
        return (lf[N] - lf[k] - lf[N - k] + lf[k + t - 1] + lf[N - k + T - t] - lf[N + T] - lf[t - 1] - lf[T - t] + lf[T])

    def _generate_prior(self, text_lengths, feats_lengths, T_text=None, T_feats=None, w=1) -> torch.Tensor:
        """(B, T_feats, T_text) log-prior on the device (osb_beta_binomial_prior): no host round trip for the lengths (the
        reference loops over .item() per sample and calls scipy, :104-114)."""
        _require_cuda(text_lengths, "AlignmentModule._generate_prior")
        T_text = T_text or int(text_lengths.max())
        T_feats = T_feats or int(feats_lengths.max())
        need = T_text + T_feats + 2
        lf = self._cache.get("lf")
        if lf is None or lf.numel() < need or lf.device != text_lengths.device:
            tab = np.concatenate([[0.0], np.cumsum(np.log(np.arange(1, max(need, 4096), dtype=np.float64)))])
            lf = torch.from_numpy(tab).to(text_lengths.device)
            self._cache["lf"] = lf
        return ops.beta_binomial_prior(lf, text_lengths.contiguous(), feats_lengths.contiguous(), T_feats, T_text)


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        from ... import _lib

        raise _lib.OsbError(f"{what}: the training path runs on libosb200 kernels only (CUDA tensors required; no CPU fallback)")


def viterbi_decode(log_p_attn, text_lengths, feats_lengths):
    """-> (durations (B,Tx) fp32, bin_loss).  Reference alignments.py:210-239; the search itself runs on the device
    (osb_mas, bit-exact with the reference's float64 numba code) instead of one host round trip per sample.  The training
    forward below uses the fused form (AlignLossFn: bin loss + forward-sum loss and their gradients in one kernel)."""
    _require_cuda(log_p_attn, "viterbi_decode")
    B, Tm, Tx = log_p_attn.shape
    path, ds = ops.mas(log_p_attn.detach().contiguous(), text_lengths.contiguous(), feats_lengths.contiguous())
    valid = path >= 0
    picked = torch.gather(log_p_attn, 2, path.clamp(min=0).long().unsqueeze(-1)).squeeze(-1)
    picked = torch.where(valid, picked, torch.zeros((), device=picked.device, dtype=picked.dtype))  # padded frames hold -inf
    per_sample = picked.sum(dim=1) / feats_lengths.to(picked.dtype)
    bin_loss = -(per_sample.sum()) / B
    return ds, bin_loss


def average_by_duration(ds, xs, text_lengths, feats_lengths):
    """Reference alignments.py:262-280, on the device (osb_average_by_duration)."""
    xs = xs.reshape(xs.shape[0], -1) if xs.dim() == 3 else xs
    return ops.average_by_duration(ds.detach().float().contiguous(), xs.detach().float().contiguous(), text_lengths.contiguous(),
                                   feats_lengths.contiguous())


def forward_sum_loss(log_p_attn, ilens, olens, blank_logprob: float = -1.0):
    """ForwardSumLoss (reference loss.py:150-194): blank column log(e^-1), per-sample log_softmax over its own
    (N_b + 1) columns, CTC with targets 1..N_b, 'mean' reduction (divide by N_b), zero_infinity, mean over batch —
    forward and gradient by the osb_forward_sum kernels."""
    _require_cuda(log_p_attn, "forward_sum_loss")
    return ForwardSumLossFn.apply(log_p_attn, ilens.contiguous(), olens.contiguous(), blank_logprob)


def fastspeech2_losses(d_outs, p_outs, e_outs, ds, ps, es, ilens):
    """FastSpeech2Loss exactly as the reference evaluates it (loss.py:83-140).  Its masks carry a stray singleton
    axis, so masked_select broadcasts: the duration term takes element (b,i) len_b times for EVERY i < Tx (padded
    positions included, target log(0 + 1e-8)); the pitch / energy terms take element (b,i) once per sample whose
    length exceeds i.  'mean' reduction over those multisets.  One kernel for the three losses and their gradients
    (osb_loss.cu)."""
    from ...autograd import FastSpeech2LossFn

    _require_cuda(d_outs, "fastspeech2_losses")
    return FastSpeech2LossFn.apply(d_outs, p_outs, e_outs, ds, ps, es, ilens)


def generator_training_forward(gen, x, x_lengths, mel, mel_lengths, pitches, energies, sids, lids, seg_rand=None):
    """Reference generator/__init__.py:72-192.  `seg_rand` (B,) in [0,1) replaces the CPU `torch.rand` draw of
    get_random_segments (utils/segments.py:32) when given (parity tests); otherwise it is drawn the same way."""
    dev = x.device
    from ..packing import step_packs
    if gen.training:
        from .modules.convnext import presample_drop_paths
        presample_drop_paths(gen, x.shape[0], dev)   # all DropPath draws of the step at once
    with step_packs(gen, dev):   # every weight pack of this forward in one launch (once the step has been recorded)
        return _generator_training_forward(gen, x, x_lengths, mel, mel_lengths, pitches, energies, sids, lids, seg_rand)


PITCH_AHEAD = True
WINDOWED_DECODER = True   # training: upsampler + ConvNeXt decoder on the frames the vocoder segment depends on only


def _decoder_halo(decoder):
    """Frames of context on either side that one output frame of the decoder depends on, or None when the decoder is not a
    stack of local blocks (Transformer decoder: every frame sees the whole sequence)."""
    from .modules.convnext import ConvNeXtBackbone

    if not isinstance(decoder, ConvNeXtBackbone):
        return None
    return sum((blk.dwconv.kernel_size[0] - 1) // 2 for blk in decoder.convnext)


def _generator_training_forward(gen, x, x_lengths, mel, mel_lengths, pitches, energies, sids, lids, seg_rand=None):
    dev = x.device
    _require_cuda(x, "OptiSpeechGenerator.forward")
    # valid + padding masks of a length vector in one launch each (osb_sequence_mask)
    x_mask, in_pad = ops.sequence_masks(x_lengths, x.shape[1])
    mel_mask, tgt_pad = ops.sequence_masks(mel_lengths, mel.shape[-1])

    # Independent branches run on side streams (forked from / joined to the current stream, so a CUDA-graph capture records
    # them as parallel branches and autograd replays their backward on the same streams): the 6144-row encoder-side kernels
    # fill 64 of 148 SMs, the 27 648-row mel-side kernels of the alignment module's feature encoder take the rest.
    main = torch.cuda.current_stream()
    am = gen.alignment_module
    s_feat = ops.side_stream(dev, 1)
    s_feat.wait_stream(main)
    with torch.cuda.stream(s_feat):
        # the beta-binomial prior depends on the lengths only: computed here, at the start of the step, instead of between the
        # text convolutions and the attention on the critical chain (30 us)
        prior = am._generate_prior(x_lengths, mel_lengths, x.shape[1], mel.shape[-1])
        fe = am.encode_feats(mel.transpose(1, 2))     # depends on the mel input only
    h, _ = gen.text_embedding(x)
    h = gen.encoder(h, in_pad)
    h = gen._speaker_language(h, sids, lids)
    s_dur = ops.side_stream(dev, 2)
    s_dur.wait_stream(main)
    with torch.cuda.stream(s_dur):
        duration_hat = gen.duration_predictor(h.detach(), in_pad)   # detached input: meets the rest only at the loss
    # the pitch prediction stack (5 convolutions) reads the encoder output only: forked here, it runs under the alignment
    # search instead of behind it (the pitch EMBEDDING below needs the search's averaged targets, the prediction does not)
    s_pitch = ops.side_stream(dev, 3)
    pitch_hat = gen.pitch_predictor.predict_ahead(h, in_pad, s_pitch) if PITCH_AHEAD else None
    # text-side alignment convs + attention on their own stream: in the backward pass (autograd replays a node on its forward
    # stream) the attention gradient chain then runs next to the predictors' instead of queueing behind them on the main stream
    s_attn = ops.side_stream(dev, ops.ATTN_SLOT)
    s_attn.wait_stream(main)
    s_attn.wait_stream(s_feat)
    with torch.cuda.stream(s_attn):
        te = am.encode_text(h)
        log_p_attn = am.attend(fe, te, x_lengths, mel_lengths, prior=prior)
    main.wait_stream(s_attn)
    # The forward-sum loss only meets the rest of the step at the final sum: its sequential recursion (one CTA per sample)
    # runs on a side stream, next to the alignment search, the predictors and the decoder.
    fs_side = ops.side_stream(dev)
    fs_side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(fs_side):   # per-sample losses and d(fs_loss)/d(log_p_attn); joined with the bin loss below
        fs_per_sample, fs_grad = ops.forward_sum(log_p_attn.detach().contiguous(), x_lengths.contiguous(), mel_lengths.contiguous(), -1.0)
    mas_path, durations = ops.mas(log_p_attn.detach().contiguous(), x_lengths.contiguous(), mel_lengths.contiguous())

    p_avg = average_by_duration(durations, pitches, x_lengths, mel_lengths)
    e_avg = average_by_duration(durations, energies, x_lengths, mel_lengths)

    # teacher forcing: the embeddings are driven by the targets, the predictions only feed the loss -> side streams
    s_energy = ops.side_stream(dev, 4)
    h, pitch_hat = gen.pitch_predictor(h, in_pad, p_avg, side_stream=s_pitch, preds=pitch_hat)
    h, energy_hat = gen.energy_predictor(h, in_pad, e_avg, side_stream=s_energy)

    # Upsampler, decoder and vocoder input carry no gradient in the reference (the vocoder is fed segment.detach(),
    # generator/__init__.py:161, and nothing else consumes the decoder output), so they run on the inference kernels — and on
    # their own stream: the acoustic-model losses do not depend on them.  In the pre-training phase nothing in the step
    # consumes wav_hat at all, so the caller may defer the join to the end of the step (`gen.defer_vocoder_join`): the
    # decoder / vocoder forward then overlaps the backward pass.
    s_voc = ops.side_stream(dev, 5)
    s_voc.wait_stream(main)
    with torch.cuda.stream(s_voc):
        with torch.no_grad():
            Tm = mel.shape[-1]
            segment_size = min(gen.segment_size, Tm)
            if seg_rand is None:
                # inside a CUDA-graph capture the draw has to live on the device (a pageable H2D copy cannot be captured)
                capturing = x.is_cuda and torch.cuda.is_current_stream_capturing()
                seg_rand = torch.rand([x.shape[0]], device=dev) if capturing else torch.rand([x.shape[0]])
            # start = floor(rand * max(len - 4 - S, 0)) (osb_glue.cu); the f0_cond crop of generator/__init__.py:156-158 is not
            # taken: WaveNeXt ignores it (wavenext/__init__.py:82)
            start_idx = ops.segment_starts(seg_rand.to(dev, non_blocking=True).float(), mel_lengths, segment_size, margin=4)
            halo = _decoder_halo(gen.decoder) if WINDOWED_DECODER else None
            if halo is not None and segment_size + 2 * halo < Tm:
                # The step consumes `segment_size` decoder frames per sample and nothing else of the decoder output (the
                # reference crops it right away, generator/__init__.py:146-152, and returns only wav_hat and the losses), and a
                # ConvNeXt decoder is local: frame t of its output depends on upsampler frames t-halo .. t+halo (3 per block).
                # So upsampler and decoder run on the window [start - halo, start + S + halo) of every sample — 88 of 864 frames
                # at the benchmark shape — with rows outside [0, Tm) zero and masked exactly as the convolutions' zero padding
                # and the padding mask present them in the full-length computation: same values for the segment.
                W = segment_size + 2 * halo
                y = gen.feature_upsampler.forward_window(h.detach(), durations, x_lengths, mel_lengths, start_idx, Tm, W, halo)
                t_full = start_idx[:, None] - halo + torch.arange(W, device=dev)[None, :]
                win_pad = ((t_full < 0) | (t_full >= mel_lengths[:, None])).contiguous()
                y = gen.decoder(y, win_pad, split=False)
                segment = y[:, halo:halo + segment_size].contiguous()
            else:
                y = gen.feature_upsampler(hs=h.detach(), ds=durations, h_masks=mel_mask, d_masks=x_mask, x_lengths=x_lengths,
                                          y_lengths=mel_lengths)
                y = gen.decoder(y, tgt_pad, split=False)
                segment = ops.gather_segments(y.contiguous(), start_idx, segment_size)

        if getattr(gen, "vocoder_needs_grad", True):
            wav_hat = gen.vocoder.forward_train(segment)
        else:
            with torch.no_grad():
                wav_hat = gen.vocoder.forward_train(segment)
    pending = []
    if getattr(gen, "defer_vocoder_join", False) and not getattr(gen, "vocoder_needs_grad", True):
        pending.append(s_voc)
    else:
        main.wait_stream(s_voc)

    main.wait_stream(s_dur)
    main.wait_stream(s_pitch)
    main.wait_stream(s_energy)
    d_loss, p_loss, e_loss = fastspeech2_losses(duration_hat, pitch_hat, energy_hat, durations, p_avg, e_avg, x_lengths)
    from ...autograd import AlignLossFn

    torch.cuda.current_stream().wait_stream(fs_side)
    align_loss, fs_loss, bin_loss = AlignLossFn.apply(log_p_attn, fs_per_sample, fs_grad, mas_path, mel_lengths)
    lc = gen.loss_coeffs
    loss = align_loss * lc.lambda_align + d_loss * lc.lambda_duration + p_loss * lc.lambda_pitch + e_loss * lc.lambda_energy
    return {
        "_pending_streams": pending,   # streams the caller has to join before the step ends (pre-training: the vocoder branch)
        # main-stream tensors that branch reads: they must not be freed (and their memory re-used by this stream) before the join
        "_pending_keepalive": [h, durations, mel_mask, x_mask, tgt_pad, in_pad, seg_rand] if pending else [],
        "wav_hat": wav_hat,
        "start_idx": start_idx,
        "segment_size": segment_size,
        "loss": loss,
        "align_loss": align_loss.detach(),
        "duration_loss": d_loss.detach(),
        "pitch_loss": p_loss.detach(),
        "energy_loss": e_loss.detach(),
        "_aux": {"log_p_attn": log_p_attn, "durations": durations, "pitch_avg": p_avg, "energy_avg": e_avg,
                 "duration_hat": duration_hat, "pitch_hat": pitch_hat, "energy_hat": energy_hat, "decoder_out": y,
                 "forwardsum_loss": fs_loss.detach(), "bin_loss": bin_loss.detach()},
    }
