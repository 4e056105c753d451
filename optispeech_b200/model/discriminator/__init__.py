"""Interface of vocoder discriminators (reference: optispeech/model/discriminator/__init__.py:11-23)."""
from abc import ABC, abstractmethod

from torch import nn


class BaseVocoderDiscriminator(nn.Module, ABC):
    @abstractmethod
    def forward_disc(self, wav, wav_hat):
        """Discriminator loss for a training batch -> (loss, log dict)."""

    @abstractmethod
    def forward_gen(self, wav, wav_hat):
        """Adversarial (generator-side) loss for a training batch -> (loss, log dict)."""

    @abstractmethod
    def forward_val(self, wav, wav_hat):
        """Validation loss -> (loss, log dict)."""
