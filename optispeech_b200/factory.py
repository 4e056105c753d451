"""Programmatic construction of the model the way Hydra's `instantiate` would do it from
configs/model/optispeech.yaml (partials for every sub-module), for callers without YAML files
(tests, bench, smoke).  Field names follow the YAML keys."""
from __future__ import annotations

from functools import partial
from types import SimpleNamespace

import torch

from .model.generator import OptiSpeechGenerator
from .model.generator.modules import ConvNeXtBackbone, DurationPredictor, EnergyPredictor, PitchPredictor, TextEmbedding, Transformer
from .model.vocoder.wavenext import WaveNeXt

DEFAULT_MODEL = dict(
    dim=256,
    segment_size=64,
    text_embedding=dict(n_vocab=250, dropout=0.1, padding_idx=0, max_source_positions=2000),
    encoder=dict(intermediate_dim=1024, num_layers=4, drop_path=0.2),
    decoder=dict(intermediate_dim=1024, num_layers=4, drop_path=0.2),
    duration_predictor=dict(num_layers=2, intermediate_dim=384, kernel_size=3, dropout=0.1),
    pitch_predictor=dict(num_layers=5, intermediate_dim=256, kernel_size=5, dropout=0.5, embed_kernel_size=9, embed_dropout=0.2),
    energy_predictor=dict(num_layers=2, intermediate_dim=384, kernel_size=3, dropout=0.5, embed_kernel_size=9, embed_dropout=0.5),
    vocoder=dict(dim=384, intermediate_dim=1152, num_layers=8, drop_path=0.1),
    loss_coeffs=dict(lambda_align=5.0, lambda_duration=1.0, lambda_pitch=1.0, lambda_energy=1.0),
    feature_extractor=dict(sample_rate=22050, n_feats=100, n_fft=1024, hop_length=256, win_length=1024, f_min=80, f_max=8000),
    disc_loss_coeffs=dict(lambda_mrd=1.0, lambda_mel=45.0, lambda_mr_stft=2.5),
    num_speakers=1,
    num_languages=1,
)

# configs/model/generator/{encoder,decoder}/transformer.yaml (selected by configs/model/transformer.yaml)
TRANSFORMER_BACKBONE = dict(attention_heads=2, linear_units=1024, num_blocks=4, dropout_rate=0.2, positional_dropout_rate=0.2,
                            attention_dropout_rate=0.2, normalize_before=True, concat_after=False, positionwise_layer_type="conv1d",
                            positionwise_conv_kernel_size=1, use_scaled_pos_enc=True, init_alpha=1.0, init_type="xavier_uniform")


def transformer_model_config() -> dict:
    """DEFAULT_MODEL with `override generator/encoder: transformer` + `override generator/decoder: transformer`."""
    cfg = {k: (dict(v) if isinstance(v, dict) else v) for k, v in DEFAULT_MODEL.items()}
    cfg["encoder"] = dict(TRANSFORMER_BACKBONE, _backbone="transformer")
    cfg["decoder"] = dict(TRANSFORMER_BACKBONE, _backbone="transformer")
    return cfg


def model_config_from_spec(spec) -> dict:
    """oracle.spec.ModelSpec (or anything with the same fields) -> factory config."""
    cfg = {k: (dict(v) if isinstance(v, dict) else v) for k, v in DEFAULT_MODEL.items()}
    cfg["dim"] = spec.dim
    cfg["segment_size"] = spec.segment_size
    cfg["text_embedding"].update(n_vocab=spec.n_vocab, max_source_positions=spec.max_source_positions)
    if getattr(spec, "backbone", "convnext") == "transformer":
        tf = dict(TRANSFORMER_BACKBONE, attention_heads=spec.tf_heads, linear_units=spec.tf_units, num_blocks=spec.tf_blocks,
                  _backbone="transformer")
        cfg["encoder"], cfg["decoder"] = dict(tf), dict(tf)
    else:
        cfg["encoder"].update(intermediate_dim=spec.enc_intermediate, num_layers=spec.enc_layers)
        cfg["decoder"].update(intermediate_dim=spec.dec_intermediate, num_layers=spec.dec_layers)
    for name in ("duration", "pitch", "energy"):
        ps = getattr(spec, name)
        cfg[f"{name}_predictor"].update(num_layers=ps.num_layers, intermediate_dim=ps.intermediate_dim, kernel_size=ps.kernel_size)
        if name != "duration":
            cfg[f"{name}_predictor"].update(embed_kernel_size=ps.embed_kernel_size)
    cfg["vocoder"].update(dim=spec.voc_dim, intermediate_dim=spec.voc_intermediate, num_layers=spec.voc_layers)
    cfg["feature_extractor"].update(sample_rate=spec.sample_rate, n_feats=spec.n_feats, n_fft=spec.n_fft, hop_length=spec.hop_length,
                                    win_length=spec.win_length, f_min=spec.f_min, f_max=spec.f_max)
    cfg["loss_coeffs"].update(lambda_align=spec.lambda_align, lambda_duration=spec.lambda_duration, lambda_pitch=spec.lambda_pitch,
                              lambda_energy=spec.lambda_energy)
    cfg["disc_loss_coeffs"].update(lambda_mrd=spec.lambda_mrd, lambda_mel=spec.lambda_mel, lambda_mr_stft=spec.lambda_mr_stft)
    cfg["num_speakers"], cfg["num_languages"] = spec.num_speakers, spec.num_languages
    return cfg


def _backbone_partial(kw: dict):
    kw = dict(kw)
    kind = kw.pop("_backbone", "convnext")
    return partial(Transformer if kind == "transformer" else ConvNeXtBackbone, **kw)


def generator_partial(cfg: dict):
    """The `generator` partial of configs/model/generator/default.yaml."""
    conv = partial(torch.nn.Conv1d)
    return partial(
        OptiSpeechGenerator,
        segment_size=cfg["segment_size"],
        text_embedding=partial(TextEmbedding, **cfg["text_embedding"]),
        encoder=_backbone_partial(cfg["encoder"]),
        duration_predictor=partial(DurationPredictor, conv_layer_class=conv, **cfg["duration_predictor"]),
        pitch_predictor=partial(PitchPredictor, conv_layer_class=conv, **cfg["pitch_predictor"]),
        energy_predictor=partial(EnergyPredictor, conv_layer_class=conv, **cfg["energy_predictor"]),
        decoder=_backbone_partial(cfg["decoder"]),
        loss_coeffs=SimpleNamespace(**cfg["loss_coeffs"]),
    )


def build_model(cfg: dict | None = None, train_args: dict | None = None, text_processor=None, max_steps: int = 2_000_000):
    """OptiSpeech with the defaults of configs/model/optispeech.yaml (AdamW 2e-4 / (0.8, 0.99) / 1e-2, cosine warm-up 1000)."""
    from transformers import get_cosine_schedule_with_warmup

    from .model import OptiSpeech
    from .model.vocoder.wavenext.disc import VocosDiscriminator

    cfg = cfg or DEFAULT_MODEL
    targs = dict(cache_generator_outputs=True, gradient_clip_val=10, gradient_accumulate_batches=None, pretraining_steps=1000,
                 evaluate_periodicity=False, evaluate_utmos=False, evaluate_pesq=False)
    targs.update(train_args or {})
    fe = SimpleNamespace(**cfg["feature_extractor"])
    data_args = SimpleNamespace(name="synthetic", num_speakers=cfg["num_speakers"], text_processor=text_processor, feature_extractor=fe,
                                batch_size=32, data_statistics=None)
    model = OptiSpeech(
        dim=cfg["dim"],
        generator=generator_partial(cfg),
        vocoder=partial(WaveNeXt, **cfg["vocoder"]),
        discriminator=partial(VocosDiscriminator, loss_coeffs=SimpleNamespace(**cfg["disc_loss_coeffs"])),
        train_args=SimpleNamespace(**targs),
        data_args=data_args,
        inference_args=SimpleNamespace(d_factor=1.1, p_factor=1.6, e_factor=1.2),
        optimizer=partial(torch.optim.AdamW, lr=2e-4, betas=[0.8, 0.99], weight_decay=1e-2),
        scheduler=partial(get_cosine_schedule_with_warmup, num_warmup_steps=1000, num_training_steps=-1),
    )
    model.max_steps = max_steps
    return model


def build_generator(cfg: dict | None = None) -> OptiSpeechGenerator:
    cfg = cfg or DEFAULT_MODEL
    fe = SimpleNamespace(**cfg["feature_extractor"])
    return generator_partial(cfg)(
        dim=cfg["dim"],
        vocoder=partial(WaveNeXt, **cfg["vocoder"]),
        feature_extractor=fe,
        data_statistics=None,
        num_speakers=cfg["num_speakers"],
        num_languages=cfg["num_languages"],
    )
