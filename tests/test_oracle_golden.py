"""Pins the CPU oracle (oracle/) against golden vectors produced by the real reference modules
(tests/golden/make_golden.py).  Runs without a GPU."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import losses as L
from oracle import model as O
from oracle.spec import ModelSpec, deterministic_state_dict, generator_shapes, tiny_spec

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def _t(a):
    return torch.from_numpy(np.asarray(a))


SPECS = {"tiny": tiny_spec, "full": ModelSpec}


@pytest.mark.parametrize("name", ["tiny", "full"])
def test_state_dict_layout_matches_reference(name):
    with open(os.path.join(GOLD, "state_dict_shapes.json")) as f:
        ref = json.load(f)[name]
    mine = {k: list(v) for k, v in generator_shapes(SPECS[name]()).items()}
    assert mine == ref
    if name == "full":  # SURVEY Appendix C / README.md:168 known answers
        total = sum(int(np.prod(v)) for v in mine.values())
        align = sum(int(np.prod(v)) for k, v in mine.items() if k.startswith("alignment_module"))
        assert total == 16_493_702 and total - align == 15_891_334


@pytest.mark.parametrize("name", ["tiny", "full"])
def test_synthesise_matches_reference(name):
    fx = _load(f"generator_{name}.npz")
    spec = SPECS[name]()
    sd = deterministic_state_dict(generator_shapes(spec), seed=0, frames_per_token=3.0)
    if name == "tiny":  # the tiny fixture carries its state_dict: the formula must reproduce it bit-for-bit
        for k, v in sd.items():
            assert np.array_equal(v.numpy(), fx[f"sd/{k}"]), k
    out = O.synthesise(sd, spec, _t(fx["x"]), _t(fx["x_lengths"]), 1.1, 1.6, 1.2)
    assert np.array_equal(out["durations"].numpy(), fx["durations"])
    assert np.array_equal(out["wav_lengths"].numpy(), fx["wav_lengths"])
    assert np.abs(out["wav"].numpy() - fx["wav"]).max() <= 2e-5
    assert np.abs(out["pitch"].numpy() - fx["pitch"]).max() <= 1e-5
    assert np.abs(out["energy"].numpy() - fx["energy"]).max() <= 1e-5


@pytest.mark.parametrize("name", ["tiny", "full"])
def test_training_forward_backward_matches_reference(name):
    fx = _load(f"generator_{name}.npz")
    spec = SPECS[name]()
    sd = deterministic_state_dict(generator_shapes(spec), seed=0, frames_per_token=3.0)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = O.generator_forward(sd, spec, _t(fx["train_x"]), _t(fx["train_x_lengths"]), _t(fx["train_mel"]), _t(fx["train_mel_lengths"]),
                              _t(fx["train_pitches"]), _t(fx["train_energies"]), _t(fx["train_seg_rand"]))
    assert np.array_equal(out["start_idx"].numpy(), fx["train_start_idx"])
    for key in ("loss", "align_loss", "duration_loss", "pitch_loss", "energy_loss"):
        assert abs(out[key].item() - float(fx[f"train_{key}"])) <= 2e-5 * max(1.0, abs(float(fx[f"train_{key}"]))), key
    assert np.abs(out["wav_hat"].detach().numpy() - fx["train_wav_hat"]).max() <= 2e-5
    out["loss"].backward()
    keys = [str(k) for k in fx["train_grad_keys"]]
    for k, ref_norm in zip(keys, fx["train_grad_norms"]):
        g = sd[k].grad
        if ref_norm < 0:  # parameter that never receives a gradient in the reference (decoder, vocoder, ...)
            assert g is None or float(g.abs().max()) == 0.0, k
        else:
            assert g is not None, k
            assert abs(float(g.norm()) - ref_norm) <= 1e-3 * max(ref_norm, 1e-6) + 1e-7, k
    if name == "tiny":
        for k in keys:
            gk = f"grad/{k}"
            if gk in fx.files:
                ref = fx[gk]
                assert np.abs(sd[k].grad.numpy() - ref).max() <= 1e-4 * max(1e-3, np.abs(ref).max()), k


@pytest.mark.parametrize("name", ["tiny", "full"])
def test_spectral_losses_match_reference(name):
    fx = _load(f"generator_{name}.npz")
    spec = SPECS[name]()
    fb = L.mel_filterbank(spec.sample_rate, spec.n_fft, spec.n_feats, spec.f_min, spec.f_max)
    assert np.abs(fb.numpy() - fx["mel_fb"]).max() <= 1e-6
    wav_hat = _t(fx["train_wav_hat"]).clone().requires_grad_(True)
    wav = _t(fx["val_wav_gt"])
    ml, stft_l, sc, mag = L.forward_val_losses(wav, wav_hat, spec, fb)
    assert abs(ml.item() - float(fx["val_mel_loss"])) <= 1e-4 * float(fx["val_mel_loss"])
    assert abs(sc.item() - float(fx["val_sc_loss"])) <= 1e-5
    assert abs(mag.item() - float(fx["val_mag_loss"])) <= 1e-5
    (ml + stft_l).backward()
    ref = fx["val_dwav_hat"]
    assert np.abs(wav_hat.grad.numpy() - ref).max() <= 1e-4 * np.abs(ref).max()
    # the crop of the ground-truth waveform (base_lightning_module.py:38-43)
    crop = O.crop_wav_segments(fx["train_wav"], _t(fx["train_start_idx"]), spec.segment_size, spec.hop_length)
    assert np.array_equal(crop.numpy(), fx["val_wav_gt"])


def test_monotonic_alignment_search_bit_exact():
    fx = _load("algorithms.npz")
    for i in range(6):
        A = O.monotonic_alignment_search(fx[f"mas{i}_lp"])
        assert np.array_equal(A, fx[f"mas{i}_A"]), i


def test_viterbi_average_forwardsum():
    fx = _load("algorithms.npz")
    lp, tl, fl = _t(fx["vd_lp"]), _t(fx["vd_tl"]), _t(fx["vd_fl"])
    ds, bin_loss = O.viterbi_decode(lp, tl, fl)
    assert np.array_equal(ds.numpy(), fx["vd_ds"])
    assert abs(bin_loss.item() - float(fx["vd_bin"])) <= 1e-6
    avg = O.average_by_duration(ds, _t(fx["avg_xs"]).squeeze(-1), tl, fl)
    assert np.abs(avg.numpy() - fx["avg_out"]).max() <= 1e-6
    assert abs(O.forward_sum_loss(lp, tl, fl).item() - float(fx["fs_loss"])) <= 1e-5
    assert abs(O.forward_sum_loss_explicit(lp, tl, fl).item() - float(fx["fs_loss"])) <= 1e-4


def test_length_regulators():
    fx = _load("algorithms.npz")
    hs, d = _t(fx["gu_hs"]), _t(fx["gu_d"])
    up = O.gaussian_upsampling(hs, d, _t(fx["gu_hmask"]), _t(fx["gu_dmask"]))
    assert np.abs(up.numpy() - fx["gu_out"]).max() <= 1e-6
    upf = O.gaussian_upsampling(hs, d.float(), _t(fx["gu_hmask"]), _t(fx["gu_dmask"]))
    assert np.abs(upf.numpy() - fx["gu_out_float"]).max() <= 1e-6
    ex, exl = O.expand_by_duration(hs, d)
    assert np.array_equal(ex.numpy(), fx["ex_out"]) and np.array_equal(exl.numpy(), fx["ex_len"])
    # integer form: frame -> token index reproduces the dense one-hot expansion exactly
    Tm = int(exl.max())
    idx = O.expand_indices(d, Tm)
    gathered = torch.where(idx[..., None] >= 0, torch.gather(hs, 1, idx.clamp(min=0)[..., None].expand(-1, -1, hs.shape[-1])),
                           torch.zeros(()))
    assert np.array_equal(gathered.numpy(), fx["ex_out"])


def test_beta_binomial_prior():
    fx = _load("algorithms.npz")
    for T, N in [(5, 3), (31, 9), (110, 24)]:
        assert np.abs(O.beta_binomial_log_prior(T, N) - fx[f"prior_{T}_{N}"]).max() <= 1e-9


# ---- Transformer backbone configuration (configs/model/transformer.yaml; tests/golden/make_golden_transformer.py) ----------
def _tf_spec():
    return ModelSpec(backbone="transformer")


def test_transformer_state_dict_layout_matches_reference():
    fx = _load("transformer.npz")
    mine = generator_shapes(_tf_spec())
    ref = {str(k): tuple(int(v) for v in str(s).split(",") if v) for k, s in zip(fx["state_dict_keys"], fx["state_dict_shapes"])}
    assert {k: tuple(v) for k, v in mine.items()} == ref
    total = sum(int(np.prod(v)) for v in mine.values())
    align = sum(int(np.prod(v)) for k, v in mine.items() if k.startswith("alignment_module"))
    assert total - align == 17_982_344  # README.md:168-170 parameter table


def test_transformer_backbone_matches_reference():
    fx = _load("transformer.npz")
    spec = _tf_spec()
    sd = deterministic_state_dict(generator_shapes(spec), seed=0, frames_per_token=3.0)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith("encoder.")}
    x = _t(fx["bb_x"]).requires_grad_(True)
    lens = _t(fx["bb_lens"])
    pad = ~(torch.arange(x.shape[1])[None] < lens[:, None])
    out = O.transformer_backbone(sd, "encoder", x, pad, spec.tf_blocks, spec.tf_heads)
    assert np.abs(out.detach().numpy() - fx["bb_out"]).max() <= 2e-5
    (out * _t(fx["bb_w"])).sum().backward()
    assert np.abs(x.grad.numpy() - fx["bb_dx"]).max() <= 2e-5 * max(1.0, np.abs(fx["bb_dx"]).max())
    for k, ref_norm in zip(fx["bb_grad_keys"], fx["bb_grad_norms"]):
        g = sd["encoder." + str(k)].grad
        # linear_k.bias has a mathematically zero gradient (softmax is shift invariant): its norm is rounding noise ~1e-6
        assert g is not None and abs(float(g.norm()) - ref_norm) <= 1e-3 * max(ref_norm, 1e-6) + 1e-5, k
    assert abs(float(sd["encoder.transformer.embed.0.alpha"].grad) - float(fx["bb_grad_alpha"])) <= 1e-3 * abs(float(fx["bb_grad_alpha"])) + 1e-6


def test_transformer_generator_matches_reference():
    fx = _load("transformer.npz")
    spec = _tf_spec()
    sd = deterministic_state_dict(generator_shapes(spec), seed=0, frames_per_token=3.0)
    out = O.synthesise(sd, spec, _t(fx["synth_x"]), _t(fx["synth_x_lengths"]), 1.1, 1.6, 1.2)
    assert np.array_equal(out["durations"].numpy(), fx["synth_durations"])
    assert np.array_equal(out["wav_lengths"].numpy(), fx["synth_wav_lengths"])
    assert np.abs(out["wav"].numpy() - fx["synth_wav"]).max() <= 2e-5
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o = O.generator_forward(sdg, spec, _t(fx["train_x"]), _t(fx["train_x_lengths"]), _t(fx["train_mel"]), _t(fx["train_mel_lengths"]),
                            _t(fx["train_pitches"]), _t(fx["train_energies"]), _t(fx["train_seg_rand"]))
    for key in ("loss", "align_loss", "duration_loss", "pitch_loss", "energy_loss"):
        assert abs(o[key].item() - float(fx[f"train_{key}"])) <= 2e-5 * max(1.0, abs(float(fx[f"train_{key}"]))), key
    assert np.abs(o["wav_hat"].detach().numpy() - fx["train_wav_hat"]).max() <= 2e-5
    o["loss"].backward()
    for k, ref_norm in zip(fx["train_grad_keys"], fx["train_grad_norms"]):
        g = sdg[str(k)].grad
        if ref_norm < 0:
            assert g is None or float(g.abs().max()) == 0.0, k
        else:
            assert g is not None and abs(float(g.norm()) - ref_norm) <= 1e-3 * max(ref_norm, 1e-6) + 1e-5, k


# ---- GAN-phase discriminators and losses (tests/golden/make_golden_disc.py) ---------------------------------
def _disc_sd():
    from oracle.discriminators import discriminator_shapes

    return deterministic_state_dict(discriminator_shapes(), seed=0)


def test_discriminator_oracle_matches_reference():
    from oracle import discriminators as D

    fx = _load("discriminator.npz")
    sd = _disc_sd()
    assert sum(int(np.prod(v.shape)) for v in sd.values()) == 41_705_968
    wav, wav_hat = _t(fx["wav"]), _t(fx["wav_hat"]).requires_grad_(True)
    spec = ModelSpec()
    with torch.no_grad():
        loss_d, log_d = D.forward_disc(sd, wav, wav_hat.detach())
    assert abs(float(loss_d) - float(fx["loss_disc"])) <= 2e-5 * abs(float(fx["loss_disc"]))
    for k, v in log_d.items():
        assert abs(float(v) - float(fx[f"disc_{k}"])) <= 2e-5 * max(1.0, abs(float(fx[f"disc_{k}"]))), k
    loss_g, log_g = D.forward_gen(sd, wav, wav_hat, spec)
    for k, v in log_g.items():
        assert abs(float(v) - float(fx[f"gen_{k}"])) <= 5e-5 * max(1.0, abs(float(fx[f"gen_{k}"]))), k
    assert abs(float(loss_g) - float(fx["loss_gen"])) <= 5e-5 * abs(float(fx["loss_gen"]))
    loss_g.backward()
    assert abs(float(wav_hat.grad.norm()) - float(fx["dwav_hat_norm"])) <= 1e-3 * float(fx["dwav_hat_norm"])
    assert np.abs(wav_hat.grad[:, ::64].numpy() - fx["dwav_hat_slice"]).max() <= 1e-3 * np.abs(fx["dwav_hat_slice"]).max()
    with torch.no_grad():
        outs, _, fr, _ = D._multi(sd, wav, wav_hat.detach(), "mpd")
        assert [o.shape[1] for o in outs] == fx["mpd_out_sizes"].tolist()
        assert np.allclose([[float(f.mean()) for f in fm] for fm in fr], fx["mpd_fmap_means"], rtol=1e-4, atol=1e-6)
        outs, _, fr, _ = D._multi(sd, wav, wav_hat.detach(), "mrd")
        assert [o.shape[1] for o in outs] == fx["mrd_out_sizes"].tolist()
        assert np.allclose([[float(f.mean()) for f in fm] for fm in fr], fx["mrd_fmap_means"], rtol=1e-4, atol=1e-6)


def test_discriminator_host_modules_match_reference():
    """The product's MPD / MRD modules (stock PyTorch this round, reference module tree) and its hinge / feature-matching
    losses against the same goldens, on the CPU."""
    from optispeech_b200.model.vocoder.wavenext.disc._discriminators import MultiPeriodDiscriminator, MultiResolutionDiscriminator
    from optispeech_b200.model.vocoder.wavenext.disc.loss import DiscriminatorLoss, FeatureMatchingLoss, GeneratorLoss

    fx = _load("discriminator.npz")
    sd = _disc_sd()
    mpd, mrd = MultiPeriodDiscriminator(), MultiResolutionDiscriminator()
    mpd.load_state_dict({k[len("multiperioddisc."):]: v for k, v in sd.items() if k.startswith("multiperioddisc.")}, strict=True)
    mrd.load_state_dict({k[len("multiresddisc."):]: v for k, v in sd.items() if k.startswith("multiresddisc.")}, strict=True)
    wav, wav_hat = _t(fx["wav"]), _t(fx["wav_hat"])
    with torch.no_grad():
        r_mp, g_mp, fr_mp, fg_mp = mpd(y=wav, y_hat=wav_hat)
        r_mr, g_mr, fr_mr, fg_mr = mrd(y=wav, y_hat=wav_hat)
        l_mp, parts_mp, _ = DiscriminatorLoss()(disc_real_outputs=r_mp, disc_generated_outputs=g_mp)
        l_mr, parts_mr, _ = DiscriminatorLoss()(disc_real_outputs=r_mr, disc_generated_outputs=g_mr)
        assert abs(float(l_mp / len(parts_mp)) - float(fx["disc_loss_mp"])) <= 2e-5 * max(1.0, float(fx["disc_loss_mp"]))
        assert abs(float(l_mr / len(parts_mr)) - float(fx["disc_loss_mrd"])) <= 2e-5 * max(1.0, float(fx["disc_loss_mrd"]))
        lg_mp, pg = GeneratorLoss()(disc_outputs=g_mp)
        assert abs(float(lg_mp / len(pg)) - float(fx["gen_loss_gen_mp"])) <= 2e-5 * max(1.0, float(fx["gen_loss_gen_mp"]))
        fm_mr = FeatureMatchingLoss()(fmap_r=fr_mr, fmap_g=fg_mr) / len(fr_mr)
        assert abs(float(fm_mr) - float(fx["gen_loss_fm_mrd"])) <= 5e-5 * max(1.0, float(fx["gen_loss_fm_mrd"]))
        assert [o.shape[1] for o in r_mp] == fx["mpd_out_sizes"].tolist() and [o.shape[1] for o in r_mr] == fx["mrd_out_sizes"].tolist()
