"""Golden vectors for the Transformer backbone configuration (configs/model/transformer.yaml, BASELINE config 4),
generated from the REAL reference modules.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_transformer.py

Outputs tests/golden/transformer.npz:
  * bb_*   : optispeech.model.generator.modules.Transformer alone (dim 256, 2 heads, 1024 units, 4 blocks), eval mode, ragged
             batch: output, d(out . w)/dx and every parameter's gradient (norms) — pins the MHA / FFN / LN semantics;
  * synth_* / train_* : OptiSpeechGenerator with Transformer encoder + decoder, same layout as generator_full.npz.
Weights come from oracle.spec.deterministic_state_dict (key-seeded), so only inputs / outputs are stored.
"""
import os
import sys
from functools import partial
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402

TF_KW = dict(attention_heads=2, linear_units=1024, num_blocks=4, dropout_rate=0.2, positional_dropout_rate=0.2,
             attention_dropout_rate=0.2, normalize_before=True, concat_after=False, positionwise_layer_type="conv1d",
             positionwise_conv_kernel_size=1, use_scaled_pos_enc=True, init_alpha=1.0, init_type="xavier_uniform")


def build_reference_generator_tf(spec):
    from optispeech.model.generator import OptiSpeechGenerator
    from optispeech.model.generator.modules import DurationPredictor, EnergyPredictor, PitchPredictor, TextEmbedding, Transformer
    from optispeech.model.vocoder.wavenext import WaveNeXt

    conv = partial(torch.nn.Conv1d)
    fe = SimpleNamespace(n_feats=spec.n_feats, n_fft=spec.n_fft, hop_length=spec.hop_length, win_length=spec.win_length,
                         sample_rate=spec.sample_rate, f_min=spec.f_min, f_max=spec.f_max)
    gen = OptiSpeechGenerator(
        dim=spec.dim, segment_size=spec.segment_size,
        text_embedding=partial(TextEmbedding, n_vocab=spec.n_vocab, dropout=0.1, padding_idx=0,
                               max_source_positions=spec.max_source_positions),
        encoder=partial(Transformer, **TF_KW),
        duration_predictor=partial(DurationPredictor, num_layers=spec.duration.num_layers, intermediate_dim=spec.duration.intermediate_dim,
                                   kernel_size=spec.duration.kernel_size, dropout=0.1, conv_layer_class=conv),
        pitch_predictor=partial(PitchPredictor, num_layers=spec.pitch.num_layers, intermediate_dim=spec.pitch.intermediate_dim,
                                kernel_size=spec.pitch.kernel_size, dropout=0.5, embed_kernel_size=spec.pitch.embed_kernel_size,
                                embed_dropout=0.2, conv_layer_class=conv),
        energy_predictor=partial(EnergyPredictor, num_layers=spec.energy.num_layers, intermediate_dim=spec.energy.intermediate_dim,
                                 kernel_size=spec.energy.kernel_size, dropout=0.5, embed_kernel_size=spec.energy.embed_kernel_size,
                                 embed_dropout=0.5, conv_layer_class=conv),
        decoder=partial(Transformer, **TF_KW),
        vocoder=partial(WaveNeXt, dim=spec.voc_dim, intermediate_dim=spec.voc_intermediate, num_layers=spec.voc_layers, drop_path=0.1),
        loss_coeffs=SimpleNamespace(lambda_align=spec.lambda_align, lambda_duration=spec.lambda_duration,
                                    lambda_pitch=spec.lambda_pitch, lambda_energy=spec.lambda_energy),
        feature_extractor=fe, num_speakers=spec.num_speakers, num_languages=spec.num_languages, data_statistics=None,
    )
    return gen


def main():
    MG.install_stubs()
    sys.path.insert(0, MG.REF)
    sys.path.append(ROOT)
    from oracle.spec import ModelSpec, deterministic_state_dict, generator_shapes

    import optispeech  # noqa: F401
    assert optispeech.__file__.startswith(MG.REF), optispeech.__file__
    from optispeech.model.generator.modules import Transformer

    torch.manual_seed(1234)
    spec = ModelSpec(backbone="transformer")
    fx = {}

    # ---- the backbone alone --------------------------------------------------------------------------------
    bb = Transformer(dim=spec.dim, **TF_KW)
    shapes = {k: tuple(v) for k, v in generator_shapes(spec).items()}
    sd_all = deterministic_state_dict(shapes, seed=0, frames_per_token=3.0)
    bb.load_state_dict({k[len("encoder."):]: v for k, v in sd_all.items() if k.startswith("encoder.")}, strict=True)
    bb.eval()
    g = torch.Generator().manual_seed(21)
    B, T = 3, 50
    lens = torch.tensor([50, 33, 1])
    x = torch.randn(B, T, spec.dim, generator=g).requires_grad_(True)
    w = torch.randn(B, T, spec.dim, generator=g)
    pad = ~(torch.arange(T)[None] < lens[:, None])
    out = bb(x, pad)
    (out * w).sum().backward()
    gkeys = sorted(k for k, _ in bb.named_parameters())
    gn = {k: float(p.grad.norm()) for k, p in bb.named_parameters()}
    fx.update(bb_x=x.detach().numpy(), bb_w=w.numpy(), bb_lens=lens.numpy(), bb_out=out.detach().numpy(), bb_dx=x.grad.numpy(),
              bb_grad_keys=np.array(gkeys), bb_grad_norms=np.array([gn[k] for k in gkeys], dtype=np.float64),
              bb_grad_q0=dict(bb.named_parameters())["transformer.encoders.0.self_attn.linear_q.weight"].grad.numpy(),
              bb_grad_alpha=float(dict(bb.named_parameters())["transformer.embed.0.alpha"].grad))

    # ---- the generator with Transformer encoder / decoder ------------------------------------------------------
    gen = build_reference_generator_tf(spec)
    ref_shapes = {k: tuple(v.shape) for k, v in gen.state_dict().items()}
    assert ref_shapes == shapes, "oracle/spec.py transformer shape table is out of date"
    gen.load_state_dict(sd_all, strict=True)
    gen.eval()
    B, Tx = 2, 37
    g = torch.Generator().manual_seed(99)
    x_lengths = torch.randint(Tx // 2, Tx + 1, (B,), generator=g)
    x_lengths[0] = Tx
    xs = torch.randint(1, 159, (B, Tx), generator=g) * (torch.arange(Tx)[None] < x_lengths[:, None])
    o = gen.synthesise(xs, x_lengths, d_factor=1.1, p_factor=1.6, e_factor=1.2)
    fx.update(synth_x=xs.numpy(), synth_x_lengths=x_lengths.numpy(), synth_wav=o["wav"].numpy(), synth_wav_lengths=o["wav_lengths"].numpy(),
              synth_durations=o["durations"].numpy(), synth_pitch=o["pitch"].numpy(), synth_energy=o["energy"].numpy())

    batch = MG.train_batch(spec, 2, 24, 110, seed=7)
    gen.zero_grad()
    o = MG.run_reference_forward(gen, batch)
    o["loss"].backward()
    fx.update({f"train_{k}": (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()})
    fx.update(train_wav_hat=o["wav_hat"].detach().numpy(), train_start_idx=o["start_idx"].numpy(), train_loss=o["loss"].item(),
              train_align_loss=o["align_loss"].item(), train_duration_loss=o["duration_loss"].item(),
              train_pitch_loss=o["pitch_loss"].item(), train_energy_loss=o["energy_loss"].item())
    gn = {k: (float(p.grad.norm()) if p.grad is not None else -1.0) for k, p in gen.named_parameters()}
    fx["train_grad_keys"] = np.array(sorted(gn))
    fx["train_grad_norms"] = np.array([gn[k] for k in sorted(gn)], dtype=np.float64)
    fx["state_dict_keys"] = np.array(sorted(ref_shapes))
    fx["state_dict_shapes"] = np.array([",".join(map(str, ref_shapes[k])) for k in sorted(ref_shapes)])
    np.savez_compressed(os.path.join(HERE, "transformer.npz"), **fx)
    print("synth frames", fx["synth_durations"].sum(1).tolist(), "loss", fx["train_loss"], "params", sum(p.numel() for p in gen.parameters()))


if __name__ == "__main__":
    main()
