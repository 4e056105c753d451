"""VocosDiscriminator: loss orchestration of the GAN phase (reference: disc/__init__.py:16-111)."""
from __future__ import annotations

from ....discriminator import BaseVocoderDiscriminator
from ._discriminators import MultiPeriodDiscriminator, MultiResolutionDiscriminator
from .loss import DiscriminatorLoss, FeatureMatchingLoss, GeneratorLoss, MelSpecReconstructionLoss, MultiResolutionSTFTLoss


class VocosDiscriminator(BaseVocoderDiscriminator):
    def __init__(self, feature_extractor, loss_coeffs):
        super().__init__()
        self.feature_extractor = feature_extractor
        self.loss_coeffs = loss_coeffs
        self.lambda_mel = self.loss_coeffs.lambda_mel
        self.lambda_mr_stft = self.loss_coeffs.lambda_mr_stft
        self.multiperioddisc = MultiPeriodDiscriminator()
        self.multiresddisc = MultiResolutionDiscriminator()
        self.gen_loss = GeneratorLoss()
        self.disc_loss = DiscriminatorLoss()
        self.feat_matching_loss = FeatureMatchingLoss()
        fe = self.feature_extractor
        self.melspec_loss = MelSpecReconstructionLoss(sample_rate=fe.sample_rate, n_fft=fe.n_fft, hop_length=fe.hop_length,
                                                      win_length=fe.win_length, n_mels=fe.n_feats, f_min=fe.f_min, f_max=fe.f_max)
        self.mr_stft_loss = MultiResolutionSTFTLoss()

    @staticmethod
    def _all_discriminators(wav):
        """On the device the eight discriminators run side by side on their own streams, joined when the block ends."""
        if wav.is_cuda:
            from . import native
            return native.deferred_join()
        import contextlib
        return contextlib.nullcontext()

    def prefetch_real(self, wav):
        """Queues every discriminator's forward pass over the real signals on the discriminator's stream, without joining:
        `forward_gen` on the same `wav` tensor picks the results up (native._pair).  Called before the generator forward."""
        if not wav.is_cuda:
            return
        from . import native
        native.reset_pack_memo()
        native.MEMO_IS_FRESH = True
        native.prefetch_real(self.multiperioddisc.discriminators, wav, native.period_forward, native.effective_weights,
                             native.DISC_STREAM_SLOT0)
        native.prefetch_real(self.multiresddisc.discriminators, wav, native.resolution_forward, native._resolution_weights,
                             native.DISC_STREAM_SLOT0 + 8)

    # Log dictionaries hold device scalars (detached); the reference calls .item() on every term (one host sync each).
    def forward_disc(self, wav, wav_hat):
        with self._all_discriminators(wav):
            real_mp, gen_mp, _, _ = self.multiperioddisc(y=wav, y_hat=wav_hat)
            real_mrd, gen_mrd, _, _ = self.multiresddisc(y=wav, y_hat=wav_hat)
        loss_mp, parts_mp, _ = self.disc_loss(disc_real_outputs=real_mp, disc_generated_outputs=gen_mp)
        loss_mrd, parts_mrd, _ = self.disc_loss(disc_real_outputs=real_mrd, disc_generated_outputs=gen_mrd)
        loss_mp = loss_mp / len(parts_mp)
        loss_mrd = loss_mrd / len(parts_mrd)
        loss = loss_mp + loss_mrd * self.loss_coeffs.lambda_mrd
        return loss, dict(loss_mp=loss_mp.detach(), loss_mrd=loss_mrd.detach())

    def forward_gen(self, wav, wav_hat):
        if wav.is_cuda:   # weight packs are shared between this turn and the discriminator turn of the same step only
            from . import native
            if not native.MEMO_IS_FRESH:     # prefetch_real of this step already emptied it (and left this step's packs in it)
                native.reset_pack_memo()
            native.MEMO_IS_FRESH = False
        with self._all_discriminators(wav):
            _, gen_mp, fr_mp, fg_mp = self.multiperioddisc(y=wav, y_hat=wav_hat)
            _, gen_mrd, fr_mrd, fg_mrd = self.multiresddisc(y=wav, y_hat=wav_hat)
        loss_gen_mp, parts_mp = self.gen_loss(disc_outputs=gen_mp)
        loss_gen_mrd, parts_mrd = self.gen_loss(disc_outputs=gen_mrd)
        loss_gen_mp = loss_gen_mp / len(parts_mp)
        loss_gen_mrd = loss_gen_mrd / len(parts_mrd)
        loss_fm_mp = self.feat_matching_loss(fmap_r=fr_mp, fmap_g=fg_mp) / len(fr_mp)
        loss_fm_mrd = self.feat_matching_loss(fmap_r=fr_mrd, fmap_g=fg_mrd) / len(fr_mrd)
        mel_loss = self._get_mel_loss(wav, wav_hat)
        mr_stft_loss = self._get_mr_stft_loss(wav, wav_hat)
        lam = self.loss_coeffs.lambda_mrd
        loss = loss_gen_mp + loss_gen_mrd * lam + loss_fm_mp + loss_fm_mrd * lam + mel_loss + mr_stft_loss
        log = dict(loss_gen_mp=loss_gen_mp.detach(), loss_gen_mrd=loss_gen_mrd.detach(), loss_fm_mp=loss_fm_mp.detach(),
                   loss_fm_mrd=loss_fm_mrd.detach(), mel_loss=mel_loss.detach(), mr_stft_loss=mr_stft_loss.detach())
        return loss, log

    def forward_val(self, wav, wav_hat):
        mel_loss = self._get_mel_loss(wav, wav_hat)
        mr_stft_loss = self._get_mr_stft_loss(wav, wav_hat)
        return mel_loss + mr_stft_loss, dict(mel_loss=mel_loss.detach(), mr_stft_loss=mr_stft_loss.detach())

    def _get_mel_loss(self, wav, wav_hat):
        return self.melspec_loss(wav_hat, wav) * self.lambda_mel

    def _get_mr_stft_loss(self, wav, wav_hat):
        sc, mag = self.mr_stft_loss(wav_hat, wav)
        return (sc + mag) * self.lambda_mr_stft
