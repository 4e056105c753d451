"""Reusable standalone modules (same export names as the reference's modules/__init__.py for the
classes that are on the ConvNeXt hot path)."""
from .convnext import ConvNeXtBackbone, ConvNeXtBlock, DropPath
from .core import DurationPredictor, EnergyPredictor, PitchPredictor, TextEmbedding, VariancePredictor
from .layers import LayerNorm, ScaledSinusoidalEmbedding
from .transformer import Transformer
