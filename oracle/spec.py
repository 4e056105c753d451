"""Model specification, parameter shapes and deterministic weights for the oracle.

TEST INFRASTRUCTURE (see oracle/__init__.py).  `ModelSpec` carries the numbers the reference
reads from configs/model/**.yaml + configs/data/feature_extractor/*.yaml; `generator_shapes`
lists the generator's state_dict keys and shapes (reference layout: SURVEY Appendix C, pinned by
tests/golden/state_dict_shapes.json); `deterministic_state_dict` fills them from a key-seeded
RNG so the same weights can be rebuilt on any machine without shipping a checkpoint.
"""
from __future__ import annotations

import zlib
from dataclasses import dataclass, field, asdict
from typing import Dict, Tuple

import torch


@dataclass
class PredictorSpec:
    num_layers: int
    intermediate_dim: int
    kernel_size: int
    embed_kernel_size: int = 9  # pitch / energy only


@dataclass
class ModelSpec:
    """configs/model/optispeech.yaml + generator/*/default.yaml + vocoder/wavenext.yaml."""

    dim: int = 256
    n_vocab: int = 250
    max_source_positions: int = 2000
    enc_layers: int = 4
    enc_intermediate: int = 1024
    dec_layers: int = 4
    dec_intermediate: int = 1024
    duration: PredictorSpec = field(default_factory=lambda: PredictorSpec(2, 384, 3))
    pitch: PredictorSpec = field(default_factory=lambda: PredictorSpec(5, 256, 5, 9))
    energy: PredictorSpec = field(default_factory=lambda: PredictorSpec(2, 384, 3, 9))
    voc_dim: int = 384
    voc_intermediate: int = 1152
    voc_layers: int = 8
    segment_size: int = 64
    # feature extractor (configs/data/feature_extractor/22.05khz.yaml)
    n_feats: int = 100
    n_fft: int = 1024
    hop_length: int = 256
    win_length: int = 1024
    sample_rate: int = 22050
    f_min: float = 80.0
    f_max: float = 8000.0
    # loss coefficients (generator/default.yaml, discriminator/vocos_disc.yaml)
    lambda_align: float = 5.0
    lambda_duration: float = 1.0
    lambda_pitch: float = 1.0
    lambda_energy: float = 1.0
    lambda_mrd: float = 1.0
    lambda_mel: float = 45.0
    lambda_mr_stft: float = 2.5
    num_speakers: int = 1
    num_languages: int = 1
    # encoder/decoder backbone: "convnext" (configs/model/optispeech.yaml) or "transformer" (configs/model/transformer.yaml:
    # attention_heads 2, linear_units 1024, num_blocks 4, pre-LN, conv1d k=1 position-wise layers, scaled positional encoding)
    backbone: str = "convnext"
    tf_heads: int = 2
    tf_units: int = 1024
    tf_blocks: int = 4

    def to_dict(self):
        return asdict(self)


def tiny_spec() -> ModelSpec:
    """A reduced configuration whose full state_dict fits in a small committed fixture."""
    return ModelSpec(
        dim=32, n_vocab=40, enc_layers=2, enc_intermediate=64, dec_layers=2, dec_intermediate=64,
        duration=PredictorSpec(2, 48, 3), pitch=PredictorSpec(3, 32, 5, 9), energy=PredictorSpec(2, 48, 3, 9),
        voc_dim=48, voc_intermediate=96, voc_layers=2, segment_size=64,
        n_feats=20, n_fft=126, hop_length=32, win_length=126, sample_rate=22050, f_min=80.0, f_max=8000.0,
    )


def _convnext_shapes(prefix: str, dim: int, inter: int, layers: int) -> Dict[str, Tuple[int, ...]]:
    s = {}
    for i in range(layers):
        p = f"{prefix}.convnext.{i}"
        s[f"{p}.gamma"] = (dim,)
        s[f"{p}.dwconv.weight"] = (dim, 1, 7)
        s[f"{p}.dwconv.bias"] = (dim,)
        s[f"{p}.norm.weight"] = (dim,)
        s[f"{p}.norm.bias"] = (dim,)
        s[f"{p}.pwconv1.weight"] = (inter, dim)
        s[f"{p}.pwconv1.bias"] = (inter,)
        s[f"{p}.pwconv2.weight"] = (dim, inter)
        s[f"{p}.pwconv2.bias"] = (dim,)
    s[f"{prefix}.final_layer_norm.weight"] = (dim,)
    s[f"{prefix}.final_layer_norm.bias"] = (dim,)
    return s


def _transformer_shapes(prefix: str, dim: int, units: int, blocks: int) -> Dict[str, Tuple[int, ...]]:
    """Transformer wrapper around the espnet Encoder (modules/transformer.py:9-27, _transformer/encoder.py:24-235):
    embed = Sequential(ScaledPositionalEncoding) -> `embed.0.alpha` (0-d); the `pe` table is a plain attribute, not a buffer."""
    t = f"{prefix}.transformer"
    s: Dict[str, Tuple[int, ...]] = {f"{t}.embed.0.alpha": ()}
    for i in range(blocks):
        p = f"{t}.encoders.{i}"
        for lin in ("linear_q", "linear_k", "linear_v", "linear_out"):
            s[f"{p}.self_attn.{lin}.weight"] = (dim, dim)
            s[f"{p}.self_attn.{lin}.bias"] = (dim,)
        s[f"{p}.feed_forward.w_1.weight"] = (units, dim, 1)
        s[f"{p}.feed_forward.w_1.bias"] = (units,)
        s[f"{p}.feed_forward.w_2.weight"] = (dim, units, 1)
        s[f"{p}.feed_forward.w_2.bias"] = (dim,)
        for n in ("norm1", "norm2"):
            s[f"{p}.{n}.weight"] = (dim,)
            s[f"{p}.{n}.bias"] = (dim,)
    s[f"{t}.after_norm.weight"] = (dim,)
    s[f"{t}.after_norm.bias"] = (dim,)
    return s


def _backbone_shapes(spec: "ModelSpec", prefix: str, inter: int, layers: int) -> Dict[str, Tuple[int, ...]]:
    if spec.backbone == "transformer":
        return _transformer_shapes(prefix, spec.dim, spec.tf_units, spec.tf_blocks)
    return _convnext_shapes(prefix, spec.dim, inter, layers)


def _predictor_shapes(prefix: str, dim: int, ps: PredictorSpec) -> Dict[str, Tuple[int, ...]]:
    s = {}
    for i in range(ps.num_layers):
        cin = dim if i == 0 else ps.intermediate_dim
        s[f"{prefix}.conv.{i}.0.weight"] = (ps.intermediate_dim, cin, ps.kernel_size)
        s[f"{prefix}.conv.{i}.0.bias"] = (ps.intermediate_dim,)
        s[f"{prefix}.conv.{i}.2.weight"] = (ps.intermediate_dim,)
        s[f"{prefix}.conv.{i}.2.bias"] = (ps.intermediate_dim,)
    s[f"{prefix}.linear.weight"] = (1, ps.intermediate_dim)
    s[f"{prefix}.linear.bias"] = (1,)
    return s


def generator_shapes(spec: ModelSpec) -> Dict[str, Tuple[int, ...]]:
    """state_dict keys -> shapes of OptiSpeechGenerator (optispeech/model/generator/__init__.py:51-70)."""
    d = spec.dim
    s: Dict[str, Tuple[int, ...]] = {}
    s["text_embedding.embed_tokens.weight"] = (spec.n_vocab, d)
    s["text_embedding.embed_positions.scale"] = (1,)
    s.update(_backbone_shapes(spec, "encoder", spec.enc_intermediate, spec.enc_layers))
    s.update(_predictor_shapes("duration_predictor", d, spec.duration))
    for name, cout, cin, k in (("t_conv1", d, d, 3), ("t_conv2", d, d, 1), ("f_conv1", d, spec.n_feats, 3),
                               ("f_conv2", d, d, 3), ("f_conv3", d, d, 1)):
        s[f"alignment_module.{name}.weight"] = (cout, cin, k)
        s[f"alignment_module.{name}.bias"] = (cout,)
    for nm, ps in (("pitch_predictor", spec.pitch), ("energy_predictor", spec.energy)):
        s.update(_predictor_shapes(f"{nm}.predictor", d, ps))
        s[f"{nm}.embed.0.weight"] = (d, 1, ps.embed_kernel_size)
        s[f"{nm}.embed.0.bias"] = (d,)
    s.update(_backbone_shapes(spec, "decoder", spec.dec_intermediate, spec.dec_layers))
    s["vocoder.embed.weight"] = (spec.voc_dim, d, 7)
    s["vocoder.embed.bias"] = (spec.voc_dim,)
    s["vocoder.norm.weight"] = (spec.voc_dim,)
    s["vocoder.norm.bias"] = (spec.voc_dim,)
    s.update(_convnext_shapes("vocoder.backbone", spec.voc_dim, spec.voc_intermediate, spec.voc_layers))
    s["vocoder.head.linear_1.weight"] = (spec.n_fft + 2, spec.voc_dim)
    s["vocoder.head.linear_1.bias"] = (spec.n_fft + 2,)
    s["vocoder.head.linear_2.weight"] = (spec.hop_length, spec.n_fft + 2)
    if spec.num_speakers > 1:
        s["sid_embed.weight"] = (spec.num_speakers, d)
    if spec.num_languages > 1:
        s["lid_embed.weight"] = (spec.num_languages, d)
    return s


def deterministic_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 0, frames_per_token: float = 4.5):
    """Fill `shapes` from per-key seeded normal draws (trained-model-like scales).

    Matrices/filters ~ N(0, 1/fan_in) scaled so activations stay O(1); norm weights ~ 1 ± 0.1;
    biases ~ 0.05 N(0,1); layer-scale gamma ~ 0.25 ± 10%.  The duration head bias is set to
    log(frames_per_token) so that synthesis yields LJSpeech-like lengths (≈4.5 frames/phoneme).
    """
    sd = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * seed) & 0x7FFFFFFF)
        r = torch.randn(shape, generator=g, dtype=torch.float32)
        leaf = key.rsplit(".", 1)[-1]
        if key.endswith("embed_positions.scale"):
            v = torch.full(shape, 0.08)
        elif leaf == "weight_g":        # weight_norm magnitude (discriminators): positive, O(1)
            v = 0.75 + 0.5 * r.abs()
        elif leaf == "alpha":
            v = 1.0 + 0.1 * r
        elif leaf == "gamma":
            v = 0.25 * (1.0 + 0.1 * r)
        elif ".norm." in key or "layer_norm" in key or ".norm1." in key or ".norm2." in key or ".after_norm." in key or (leaf in ("weight", "bias") and len(shape) == 1 and ".2." in key):
            v = (1.0 + 0.1 * r) if leaf == "weight" else 0.05 * r
        elif leaf == "bias":
            v = 0.05 * r
        elif key.endswith("embed_tokens.weight") or key.endswith("_embed.weight"):
            v = 0.3 * r
        else:  # conv / linear weights
            fan_in = 1
            for s_ in shape[1:]:
                fan_in *= s_
            v = r * (1.0 / max(fan_in, 1)) ** 0.5
        sd[key] = v.contiguous()
    if "duration_predictor.linear.bias" in sd:
        sd["duration_predictor.linear.bias"] = torch.full((1,), float(torch.log(torch.tensor(frames_per_token))))
        sd["duration_predictor.linear.weight"] = sd["duration_predictor.linear.weight"] * 0.3
    if "text_embedding.embed_tokens.weight" in sd:
        sd["text_embedding.embed_tokens.weight"][0].zero_()  # padding_idx row
    # keep the waveform inside (-1, 1) most of the time so that clip() is exercised but not dominant
    if "vocoder.head.linear_2.weight" in sd:
        sd["vocoder.head.linear_2.weight"] = sd["vocoder.head.linear_2.weight"] * 0.5
    return sd
