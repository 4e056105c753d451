"""Harness around the REAL reference modules (mush42/optispeech @ 3bdde20).

TEST INFRASTRUCTURE (see oracle/__init__.py): only tests/, tests/golden/make_*.py, __graft_entry__ and bench.py's
reference / baseline legs may import this module; the product package never does (tests/test_abi_cpu.py enforces it).

The reference is a pure-Python package.  It is imported as the top-level package `optispeech` either from
`/root/reference` (build container) or from `oracle/_ref/` — a git-ignored copy made by `python -m oracle.make_ref`
(__graft_entry__.build() runs it), which travels to the GPU box with the snapshot like the built `.so` does.  The
packages the reference imports at module top but never uses on this path (lightning, hydra, omegaconf, matplotlib —
optispeech/utils/__init__.py:1-24, model/base_lightning_module.py:14-15) are not installed in this image and are replaced
by empty stub modules.

Name collision: the repository ships its own `optispeech/` alias package (Hydra `_target_` resolution).  A process that
wants the reference must call `import_reference()` BEFORE anything imports that alias; bench.py and the golden generators
run the reference in their own processes for that reason.
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import types
from functools import partial
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_CANDIDATES = ("/root/reference", os.path.join(HERE, "_ref"))


def reference_root():
    """Directory holding the reference's `optispeech/` package, or None."""
    for c in REF_CANDIDATES:
        if os.path.isfile(os.path.join(c, "optispeech", "model", "generator", "__init__.py")):
            return c
    return None


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []
    m.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_stubs():
    """Stub modules for the uninstalled packages the reference imports at module top."""
    _stub("matplotlib", use=lambda *a, **k: None)
    _stub("matplotlib.pyplot", Figure=object)
    _stub("omegaconf", DictConfig=dict, OmegaConf=object, open_dict=lambda *a, **k: None)
    _stub("hydra")
    _stub("hydra.core")
    _stub("hydra.core.hydra_config", HydraConfig=object)
    _stub("lightning", LightningModule=torch.nn.Module, Callback=object, LightningDataModule=object, Trainer=object)
    _stub("lightning.pytorch")
    _stub("lightning.pytorch.loggers", Logger=object)
    _stub("lightning.pytorch.utilities", rank_zero_only=lambda f: f, grad_norm=lambda *a, **k: {})


def import_reference():
    """Make `import optispeech` resolve to the reference.  Returns its root directory."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not found: neither /root/reference nor oracle/_ref (run `python -m oracle.make_ref` "
                           "in the build container)")
    if "optispeech" in sys.modules:
        f = getattr(sys.modules["optispeech"], "__file__", "") or ""
        if not f.startswith(root):
            raise RuntimeError("the repository's `optispeech` alias package is already imported in this process; "
                               "run the reference in its own process")
        return root
    install_stubs()
    sys.path.insert(0, root)
    if ROOT not in sys.path:
        sys.path.append(ROOT)  # for `oracle` only; the reference must win the name `optispeech`
    import optispeech  # noqa: F401

    assert optispeech.__file__.startswith(root), optispeech.__file__
    return root


def feature_extractor_ns(spec):
    return SimpleNamespace(n_feats=spec.n_feats, n_fft=spec.n_fft, hop_length=spec.hop_length, win_length=spec.win_length,
                           sample_rate=spec.sample_rate, f_min=spec.f_min, f_max=spec.f_max)


def build_reference_generator(spec):
    """The reference OptiSpeechGenerator for `spec` (ConvNeXt configuration), built from partials exactly as Hydra's
    `_partial_: true` would (configs/model/generator/default.yaml)."""
    from optispeech.model.generator import OptiSpeechGenerator
    from optispeech.model.generator.modules import (ConvNeXtBackbone, DurationPredictor, EnergyPredictor, PitchPredictor,
                                                    TextEmbedding)
    from optispeech.model.vocoder.wavenext import WaveNeXt

    conv = partial(torch.nn.Conv1d)
    fe = feature_extractor_ns(spec)
    gen = OptiSpeechGenerator(
        dim=spec.dim,
        segment_size=spec.segment_size,
        text_embedding=partial(TextEmbedding, n_vocab=spec.n_vocab, dropout=0.1, padding_idx=0,
                               max_source_positions=spec.max_source_positions),
        encoder=partial(ConvNeXtBackbone, intermediate_dim=spec.enc_intermediate, num_layers=spec.enc_layers, drop_path=0.2),
        duration_predictor=partial(DurationPredictor, num_layers=spec.duration.num_layers,
                                   intermediate_dim=spec.duration.intermediate_dim, kernel_size=spec.duration.kernel_size,
                                   dropout=0.1, conv_layer_class=conv),
        pitch_predictor=partial(PitchPredictor, num_layers=spec.pitch.num_layers, intermediate_dim=spec.pitch.intermediate_dim,
                                kernel_size=spec.pitch.kernel_size, dropout=0.5, embed_kernel_size=spec.pitch.embed_kernel_size,
                                embed_dropout=0.2, conv_layer_class=conv),
        energy_predictor=partial(EnergyPredictor, num_layers=spec.energy.num_layers, intermediate_dim=spec.energy.intermediate_dim,
                                 kernel_size=spec.energy.kernel_size, dropout=0.5, embed_kernel_size=spec.energy.embed_kernel_size,
                                 embed_dropout=0.5, conv_layer_class=conv),
        decoder=partial(ConvNeXtBackbone, intermediate_dim=spec.dec_intermediate, num_layers=spec.dec_layers, drop_path=0.2),
        vocoder=partial(WaveNeXt, dim=spec.voc_dim, intermediate_dim=spec.voc_intermediate, num_layers=spec.voc_layers, drop_path=0.1),
        loss_coeffs=SimpleNamespace(lambda_align=spec.lambda_align, lambda_duration=spec.lambda_duration,
                                    lambda_pitch=spec.lambda_pitch, lambda_energy=spec.lambda_energy),
        feature_extractor=fe,
        num_speakers=spec.num_speakers,
        num_languages=spec.num_languages,
        data_statistics=None,
    )
    return gen, fe


def build_reference_discriminator(spec, fe):
    from optispeech.model.vocoder.wavenext.disc import VocosDiscriminator

    return VocosDiscriminator(feature_extractor=fe, loss_coeffs=SimpleNamespace(lambda_mrd=spec.lambda_mrd, lambda_mel=spec.lambda_mel,
                                                                               lambda_mr_stft=spec.lambda_mr_stft))


def run_reference_forward(gen, batch, device=None):
    """generator.forward with the segment draw pinned to batch["seg_rand"] when present (the reference draws
    `torch.rand(B)` on the CPU generator, utils/segments.py:32)."""
    import optispeech.utils.segments as seg

    def dev(t):
        return t.to(device) if device is not None else t

    seg_rand = batch.get("seg_rand")
    real_torch = seg.torch
    if seg_rand is not None:
        seg.torch = _TorchWithRand(real_torch, seg_rand)
    try:
        out = gen(x=dev(batch["x"]), x_lengths=dev(batch["x_lengths"]), mel=dev(batch["mel"]), mel_lengths=dev(batch["mel_lengths"]),
                  pitches=dev(batch["pitches"]), energies=dev(batch["energies"]), sids=None, lids=None)
    finally:
        seg.torch = real_torch
    return out


class _TorchWithRand:
    """`torch` as seen by optispeech.utils.segments, with `rand` returning a fixed draw (nothing else is touched)."""

    def __init__(self, real, draw):
        self._real, self._draw = real, draw

    def rand(self, *a, **k):
        return self._draw.clone().cpu()

    def __getattr__(self, name):
        return getattr(self._real, name)


def reference_training_step(gen, opt, sched, batch, clip: float = 10.0, device=None):
    """One generator pre-training `training_step` restated from base_lightning_module.py:78-110 around the reference
    modules: forward, `manual_backward` = loss.backward(), `clip_gradients(norm, 10)` = clip_grad_norm_, optimizer step,
    scheduler step.  Returns the detached forward outputs."""
    out = run_reference_forward(gen, batch, device)
    opt.zero_grad()
    out["loss"].backward()
    torch.nn.utils.clip_grad_norm_([p for g in opt.param_groups for p in g["params"] if p.grad is not None], clip)
    opt.step()
    if sched is not None:
        sched.step()
    return out


def reference_optimizer(gen, lr=2e-4, betas=(0.8, 0.99), weight_decay=1e-2, warmup=1000, total=1_000_000):
    """configs/model/optimizer/adamw.yaml + scheduler/cosine_with_warmup.yaml (max_steps // 2 training steps)."""
    from transformers import get_cosine_schedule_with_warmup

    opt = torch.optim.AdamW(gen.parameters(), lr=lr, betas=betas, weight_decay=weight_decay)
    sched = get_cosine_schedule_with_warmup(opt, num_warmup_steps=warmup, num_training_steps=total)
    return opt, sched
