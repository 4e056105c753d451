"""Drop-in boundary on the CPU: config composition / instantiation, dotted-path aliases, class surface, state_dict
layout, checkpoint round trip, value containers, host-side helpers.  No kernels are launched."""
import functools
import inspect
import io
import json
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def model():
    from optispeech_b200.config import build_from_config

    torch.manual_seed(0)
    return build_from_config("optispeech")


def test_compose_matches_reference_yaml_semantics():
    from optispeech_b200.config import compose_model

    cfg = compose_model("optispeech")
    assert cfg["_target_"] == "optispeech.model.OptiSpeech" and cfg["dim"] == 256
    gen = cfg["generator"]
    assert gen["_target_"] == "optispeech.model.generator.OptiSpeechGenerator" and gen["_partial_"] is True
    assert gen["encoder"]["_target_"].endswith("ConvNeXtBackbone") and gen["encoder"]["num_layers"] == 4
    assert gen["pitch_predictor"]["kernel_size"] == 5 and gen["pitch_predictor"]["conv_layer_class"]["_target_"] == "torch.nn.Conv1d"
    assert cfg["vocoder"]["dim"] == 384 and cfg["vocoder"]["intermediate_dim"] == 1152 and cfg["vocoder"]["num_layers"] == 8
    assert cfg["optimizer"]["lr"] == pytest.approx(2e-4) and cfg["optimizer"]["betas"] == [0.8, 0.99]
    assert cfg["data_args"]["feature_extractor"]["sample_rate"] == 22050      # ${data.feature_extractor} interpolation
    assert cfg["train_args"]["pretraining_steps"] == 1000 and cfg["inference_args"]["p_factor"] == 1.6
    # `override generator/encoder: transformer` of configs/model/transformer.yaml
    tcfg = compose_model("transformer")
    assert tcfg["generator"]["encoder"]["_target_"].endswith("modules.Transformer")
    assert tcfg["generator"]["decoder"]["attention_heads"] == 2
    assert tcfg["generator"]["duration_predictor"]["_target_"].endswith("DurationPredictor")


def test_instantiated_model_has_reference_layout(model):
    from optispeech_b200.model import OptiSpeech
    from optispeech_b200.model.generator import OptiSpeechGenerator

    assert isinstance(model, OptiSpeech) and isinstance(model.generator, OptiSpeechGenerator)
    with open(os.path.join(GOLD, "state_dict_shapes.json")) as f:
        ref = json.load(f)["full"]
    mine = {k: list(v.shape) for k, v in model.generator.state_dict().items()}
    assert mine == ref, "generator state_dict keys/shapes differ from the reference module's"
    assert sum(p.numel() for p in model.discriminator.parameters()) == 41_705_968
    sd = model.state_dict()
    for k in ("discriminator.melspec_loss.mel_spec.spectrogram.window", "discriminator.melspec_loss.mel_spec.mel_scale.fb",
              "discriminator.mr_stft_loss.stft_losses.2.window", "discriminator.multiperioddisc.discriminators.0.convs.0.weight_g"):
        assert k in sd, k
    assert "generator.text_embedding.embed_positions.inv_freq" not in sd  # non-persistent buffer, as in the reference
    assert model.sample_rate == 22050 and model.hop_length == 256
    assert isinstance(model.hparams.optimizer, functools.partial) and model.hparams.optimizer.func is torch.optim.AdamW


def test_dotted_paths_of_the_reference_resolve_to_this_implementation():
    import importlib

    import optispeech_b200.model.generator.modules as real

    alias = importlib.import_module("optispeech.model.generator.modules")
    assert alias is real
    from optispeech.model import OptiSpeech as A
    from optispeech_b200.model import OptiSpeech as B

    assert A is B
    from optispeech.model.vocoder.wavenext.disc import VocosDiscriminator  # noqa: F401
    from optispeech.values import InferenceInputs  # noqa: F401


def test_constructor_signatures_match_reference():
    from optispeech_b200.model import OptiSpeech
    from optispeech_b200.model.generator import OptiSpeechGenerator
    from optispeech_b200.model.generator.modules import ConvNeXtBackbone, ConvNeXtBlock, TextEmbedding, VariancePredictor
    from optispeech_b200.model.vocoder.wavenext import WaveNeXt

    def names(fn):
        return [p for p in inspect.signature(fn).parameters if p != "self"]

    assert names(OptiSpeech.__init__) == ["dim", "generator", "vocoder", "discriminator", "train_args", "data_args", "inference_args",
                                          "optimizer", "scheduler"]
    assert names(OptiSpeechGenerator.__init__) == ["dim", "segment_size", "text_embedding", "encoder", "duration_predictor",
                                                   "pitch_predictor", "energy_predictor", "decoder", "vocoder", "loss_coeffs",
                                                   "feature_extractor", "num_speakers", "num_languages", "data_statistics", "kwargs"]
    assert names(ConvNeXtBackbone.__init__) == ["dim", "intermediate_dim", "num_layers", "drop_path", "layer_scale_init_value"]
    assert names(ConvNeXtBlock.__init__) == ["dim", "intermediate_dim", "drop_path", "layer_scale_init_value"]
    assert names(WaveNeXt.__init__) == ["input_channels", "dim", "intermediate_dim", "num_layers", "n_fft", "hop_length", "sample_rate",
                                        "drop_path", "layer_scale_init_value"]
    assert names(TextEmbedding.__init__) == ["dim", "n_vocab", "dropout", "padding_idx", "max_source_positions"]
    assert names(VariancePredictor.__init__) == ["dim", "num_layers", "intermediate_dim", "kernel_size", "dropout", "conv_layer_class"]
    assert names(OptiSpeechGenerator.synthesise)[:7] == ["x", "x_lengths", "sids", "lids", "d_factor", "p_factor", "e_factor"]
    # the reference's eight arguments, plus one optional trailing extra (the host-made segment draw, BaseModule.stage_batch)
    assert names(OptiSpeechGenerator.forward) == ["x", "x_lengths", "mel", "mel_lengths", "pitches", "energies", "sids", "lids", "seg_rand"]
    assert inspect.signature(OptiSpeechGenerator.forward).parameters["seg_rand"].default is None


def test_reference_error_conventions():
    from optispeech_b200.factory import build_model

    with pytest.raises(ValueError, match="gradient_accumulate_batches"):
        build_model(train_args=dict(gradient_accumulate_batches=0))
    from optispeech_b200.factory import DEFAULT_MODEL

    bad = dict(DEFAULT_MODEL, num_speakers=0)
    with pytest.raises(ValueError, match="num_speakers"):
        build_model(bad)
    m = build_model()
    with pytest.raises(RuntimeError, match="text_processor"):
        m.prepare_input("hello")


def test_checkpoint_round_trip(model, tmp_path):
    from optispeech_b200.model import OptiSpeech

    path = tmp_path / "m.ckpt"
    model.save_checkpoint(str(path), epoch=7, global_step=123)
    loaded = OptiSpeech.load_from_checkpoint(str(path), map_location="cpu")
    assert loaded.ckpt_loaded_epoch == 7
    a, b = model.state_dict(), loaded.state_dict()
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_value_containers():
    from optispeech_b200.values import InferenceInputs, InferenceOutputs, numpy_pad_sequences, numpy_unpad_sequences

    inp = InferenceInputs.from_ids_and_lengths([[1, 2, 3], [4, 5]], [3, 2], clean_text="x", d_factor=1.1)
    assert isinstance(inp.x, np.ndarray) and inp.x.dtype == np.int64 and inp.x.tolist() == [[1, 2, 3], [4, 5, 0]]
    t = inp.as_torch()
    assert isinstance(t.x, torch.Tensor) and t.x_lengths.tolist() == [3, 2] and t.d_factor == 1.1
    out = InferenceOutputs(wav=np.zeros((2, 10), dtype=np.float32), wav_lengths=np.array([10, 4]), latency=1, rtf=0.1)
    assert [len(w) for w in out] == [10, 4]
    out_t = InferenceOutputs(wav=torch.zeros(2, 10), wav_lengths=torch.tensor([10, 4]), latency=1, rtf=0.1)
    assert [len(w) for w in out_t.unbatched_wavs()] == [10, 4]
    assert numpy_pad_sequences([[1], [1, 2]]).shape == (2, 2)
    with pytest.raises(ValueError):
        numpy_unpad_sequences(np.zeros((2, 3)), np.array([4, 1]))


def test_host_helpers_match_oracle():
    from optispeech_b200.utils import get_segments, get_segments_numpy, sequence_mask
    from oracle import model as O

    lens = torch.tensor([5, 2, 7])
    assert torch.equal(sequence_mask(lens, 7), O.sequence_mask(lens, 7))
    x = torch.arange(2 * 3 * 20, dtype=torch.float32).view(2, 3, 20)
    starts = torch.tensor([4, 11])
    seg = get_segments(x, starts, 6)
    ref = torch.stack([x[0, :, 4:10], x[1, :, 11:17]])
    assert torch.equal(seg, ref)
    assert np.array_equal(get_segments_numpy(x.numpy(), starts.numpy(), 6), ref.numpy())


def test_prior_table_matches_scipy_golden():
    from optispeech_b200.model.generator.training import AlignmentModule

    fx = np.load(os.path.join(GOLD, "algorithms.npz"))
    for T, N in [(5, 3), (31, 9), (110, 24)]:
        assert np.abs(AlignmentModule._log_prior(T, N) - fx[f"prior_{T}_{N}"]).max() <= 1e-9


def test_stage_batch_cuts_the_reference_crop_on_the_host(model):
    """BaseModule.stage_batch = get_random_segments' start indices (CPU generator, utils/segments.py:29-35) + get_segments_numpy
    (utils/segments.py:63-72) of the reference's _process_batch, for batches whose waveform is still in host memory."""
    from oracle import model as O

    g = torch.Generator().manual_seed(3)
    B, Tm, hop = 4, 300, model.hop_length
    ml = torch.tensor([300, 71, 68, 150])
    wav = (torch.rand(B, Tm * hop, generator=g) * 2 - 1).numpy().astype(np.float32)
    batch = dict(x=torch.zeros(B, 10, dtype=torch.long), x_lengths=torch.full((B,), 10), mel=torch.zeros(B, 100, Tm), mel_lengths=ml,
                 pitches=torch.zeros(B, Tm), energies=torch.zeros(B, Tm), wav=wav, sids=None, lids=None)
    torch.manual_seed(11)
    staged = model.stage_batch(batch)
    torch.manual_seed(11)
    rand = torch.rand(B)   # the same CPU draw
    assert "wav" not in staged and torch.equal(staged["seg_rand"], rand)
    start = O.segment_starts((ml - 4).float(), 64, rand)
    assert start[2] == 0 and (start <= (ml - 68).clamp(min=0)).all()
    ref = O.crop_wav_segments(wav, start, 64, hop)
    assert torch.equal(staged["wav_segment"], ref)
    assert model.batch_h2d_bytes(staged) < model.batch_h2d_bytes(batch) - wav.nbytes + ref.numel() * 4 + 64
    assert model.stage_batch(staged) is staged          # idempotent
    as3d = dict(batch, wav=wav[:, None, :])             # (B, 1, Tw) layout of get_segments_numpy's input
    torch.manual_seed(11)
    assert torch.equal(model.stage_batch(as3d)["wav_segment"], ref)


def test_transformer_config_instantiates_with_reference_layout():
    from optispeech_b200.config import build_from_config
    from optispeech_b200.model.generator.modules import Transformer

    m = build_from_config("transformer")
    assert isinstance(m.generator.encoder, Transformer) and isinstance(m.generator.decoder, Transformer)
    fx = np.load(os.path.join(GOLD, "transformer.npz"))
    ref = {str(k): tuple(int(v) for v in str(s).split(",") if v) for k, s in zip(fx["state_dict_keys"], fx["state_dict_shapes"])}
    assert {k: tuple(v.shape) for k, v in m.generator.state_dict().items()} == ref
    assert float(m.generator.encoder.transformer.embed[0].alpha) == 1.0


def test_step_packer_plan_and_job_table(monkeypatch):
    """model/packing.StepPacker on the CPU: record the pack requests of one forward, serve the same sequence from the plan
    afterwards (the single osb_pack_multi launch is replaced by an interpreter of its job table), drop the plan on deviation."""
    import ctypes as C

    from optispeech_b200 import _lib
    from optispeech_b200.model.packing import StepPacker

    g = torch.Generator().manual_seed(4)
    w_a, w_b = torch.randn(6, 8, generator=g), torch.randn(10, 8, generator=g)      # two Linear weights stacked along N
    scale = torch.rand(8, generator=g) + 0.5
    w_c = torch.randn(5, 3, 7, generator=g)                                          # Conv1d (N, Cin, k)

    def nk_direct(ws, cs):
        w = torch.cat(ws, 0) * (cs if cs is not None else 1.0)
        return w.half().view(1, -1, 8)

    def conv_direct(w, kp):
        out = torch.zeros(w.shape[2], w.shape[0], kp)
        out[:, :, : w.shape[1]] = w.permute(2, 0, 1)
        return out.half()

    launches = []

    class FakeLib:
        @staticmethod
        def osb_pack_multi(table_ptr, n_jobs, total, stream):
            jobs = (_lib.PackJob * n_jobs).from_address(table_ptr)
            done = 0
            for j in jobs:
                n = j.rows * j.dst_cols * (j.k if j.kind == 1 else 1)
                assert j.first_elem == done
                src_n = j.rows * j.cols * (j.k if j.kind == 1 else 1)
                src = np.ctypeslib.as_array((C.c_float * src_n).from_address(j.src))
                dst = np.ctypeslib.as_array((C.c_uint16 * n).from_address(j.dst)).view(np.float16)
                if j.kind == 0:
                    m = src.reshape(j.rows, j.cols).copy()
                    if j.col_scale:
                        m *= np.ctypeslib.as_array((C.c_float * j.cols).from_address(j.col_scale))[None, :]
                    buf = np.zeros((j.rows, j.dst_cols), np.float32)
                    buf[:, : j.cols] = m
                else:
                    buf = np.zeros((j.k, j.rows, j.dst_cols), np.float32)
                    buf[:, :, : j.cols] = src.reshape(j.rows, j.cols, j.k).transpose(2, 0, 1)
                dst[:] = buf.astype(np.float16).reshape(-1)
                done += n
            assert done == total
            launches.append(n_jobs)
            return 0

    monkeypatch.setattr(_lib, "load", lambda: FakeLib)
    from optispeech_b200 import ops
    monkeypatch.setattr(ops, "_stream", lambda: None)

    pk = StepPacker()
    dev = torch.device("cpu")

    def forward(third_shape_change=False):
        pk.begin(dev)
        a = pk.request("nk", [w_a, w_b], scale, None, lambda: nk_direct([w_a, w_b], scale))
        b = pk.request("nk", [w_b], None, None, lambda: nk_direct([w_b], None))
        src = w_c[:4].contiguous() if third_shape_change else w_c
        c = pk.request("conv", [src], None, 8, lambda: conv_direct(src, 8))
        pk.end()
        return a, b, c

    a0, b0, c0 = forward()                                   # recording step: direct packs, plan built at end()
    assert pk.plan is not None and pk.plan["n_jobs"] == 4 and not launches
    w_a.mul_(2.0); w_c.add_(1.0)                             # "optimizer step": same storage, new values
    a1, b1, c1 = forward()                                   # served from ONE launch of the job table
    assert launches == [4]
    assert torch.equal(a1, nk_direct([w_a, w_b], scale)) and torch.equal(b1, nk_direct([w_b], None)) and torch.equal(c1, conv_direct(w_c, 8))
    assert a1.data_ptr() == pk.plan["outs"][0].data_ptr() and a1.shape == a0.shape
    forward(third_shape_change=True)                         # the step deviates: plan dropped, that request packed directly
    assert pk.plan is None
    forward()                                                # records again
    assert pk.plan is not None and launches == [4, 4]


def test_presampled_drop_paths_follow_the_reference_distribution():
    """presample_drop_paths draws Bernoulli(keep_l) / keep_l per sample and block (reference convnext.py:121-129) for all blocks
    at once; every DropPath then consumes its own row exactly once."""
    from optispeech_b200.model.generator.modules.convnext import ConvNeXtBackbone, DropPath, presample_drop_paths

    torch.manual_seed(0)
    bb = ConvNeXtBackbone(dim=256, intermediate_dim=1024, num_layers=4, drop_path=0.2).train()
    mods = [m for m in bb.modules() if isinstance(m, DropPath)]
    assert [round(m.drop_prob, 4) for m in mods] == [0.0667, 0.1333, 0.2]          # linspace(0, 0.2, 4) without the zero rate
    B = 4096
    presample_drop_paths(bb, B, torch.device("cpu"))
    for m in mods:
        keep = 1.0 - m.drop_prob
        s = m.sample_scale(B, torch.device("cpu"))
        vals = set(torch.unique(s).tolist())
        assert vals <= {0.0, float(torch.tensor(1.0) / torch.tensor(keep))} or all(abs(v) < 1e-6 or abs(v - 1 / keep) < 1e-5 for v in vals)
        assert abs(float((s > 0).float().mean()) - keep) < 0.03
        assert "_presampled" not in m.__dict__                                      # consumed
    bb.eval()
    presample_drop_paths(bb, B, torch.device("cpu"))
    assert all(m.sample_scale(B, torch.device("cpu")) is None for m in mods)          # eval: identity


def test_reference_style_checkpoint_loads_without_omegaconf(tmp_path):
    """A checkpoint laid out like the reference's (Lightning dict; `hyper_parameters` = partials of `optispeech.*` classes,
    OmegaConf DictConfig / ListConfig containers with *Node leaves, live text-processor / feature-extractor objects;
    optispeech/model/optispeech.py:26) loads through optispeech_b200.checkpoint without omegaconf / lightning / the text
    front-end being importable.  OmegaConf is not installed here, so the containers are written by stand-in classes with
    OmegaConf's module paths and pickled `__dict__` layout (`_content`, `_metadata`, `_parent`, `_flags_cache`; nodes: `_val`)."""
    import importlib.machinery
    import sys
    import types
    from functools import partial

    from optispeech_b200.checkpoint import AttrDict, Placeholder
    from optispeech_b200.factory import DEFAULT_MODEL, build_model
    from optispeech_b200.model import OptiSpeech

    fake = {}

    def mod(name):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        fake[name] = m
        return m

    base, dictconfig, listconfig, nodes = mod("omegaconf.base"), mod("omegaconf.dictconfig"), mod("omegaconf.listconfig"), mod("omegaconf.nodes")
    mod("omegaconf")
    text_mod, fe_mod = mod("optispeech.text"), mod("optispeech.dataset.feature_extractors")

    def cls(m, name):
        c = type(name, (), {})
        c.__module__ = m.__name__
        c.__qualname__ = name
        setattr(m, name, c)
        return c

    Meta, CMeta = cls(base, "Metadata"), cls(base, "ContainerMetadata")
    DictConfig, ListConfig = cls(dictconfig, "DictConfig"), cls(listconfig, "ListConfig")
    AnyNode, IntegerNode, FloatNode, BooleanNode = (cls(nodes, n) for n in ("AnyNode", "IntegerNode", "FloatNode", "BooleanNode"))
    TextProcessor, FeatureExtractor = cls(text_mod, "TextProcessor"), cls(fe_mod, "CommonFeatureExtractor")

    def wrap(v, parent=None):
        if isinstance(v, dict):
            d = DictConfig()
            d.__dict__.update(_metadata=CMeta(), _parent=parent, _flags_cache=None, _content={})
            d._metadata.__dict__.update(ref_type=object, object_type=dict, optional=True, key=None, flags={"allow_objects": True})
            for k, x in v.items():
                d._content[k] = wrap(x, d)
            return d
        if isinstance(v, (list, tuple)):
            l = ListConfig()
            l.__dict__.update(_metadata=CMeta(), _parent=parent, _flags_cache=None, _content=[])
            l._content.extend(wrap(x, l) for x in v)
            return l
        node = {bool: BooleanNode, int: IntegerNode, float: FloatNode}.get(type(v), AnyNode)()
        node.__dict__.update(_metadata=Meta(), _parent=parent, _flags_cache=None, _val=v)
        return node

    src = build_model(DEFAULT_MODEL)
    hp = vars(src.hparams)
    tp = TextProcessor()
    tp.__dict__.update(languages=["en-us"], num_languages=1, is_multi_language=False, default_language="en-us", add_blank=True)
    fe = FeatureExtractor()
    fe.__dict__.update(vars(hp["data_args"].feature_extractor), center=False, pitch_extractor=None)
    data_args = dict(vars(hp["data_args"]), text_processor=tp, feature_extractor=fe)
    gen = hp["generator"]
    gen_kw = dict(gen.keywords, loss_coeffs=wrap(vars(gen.keywords["loss_coeffs"])))          # nested DictConfig keyword
    disc_kw = dict(hp["discriminator"].keywords, loss_coeffs=wrap(vars(hp["discriminator"].keywords["loss_coeffs"])))
    upstream = dict(dim=hp["dim"], generator=partial(gen.func, **gen_kw), vocoder=hp["vocoder"],
                    discriminator=partial(hp["discriminator"].func, **disc_kw),
                    train_args=wrap(vars(hp["train_args"])), data_args=wrap(data_args), inference_args=wrap(vars(hp["inference_args"])),
                    optimizer=partial(torch.optim.AdamW, lr=2e-4, betas=wrap([0.8, 0.99]), weight_decay=1e-2), scheduler=hp["scheduler"])
    path = tmp_path / "upstream.ckpt"
    saved = {k: sys.modules.get(k) for k in fake}
    sys.modules.update(fake)
    try:
        torch.save({"state_dict": src.state_dict(), "hyper_parameters": upstream, "epoch": 11, "global_step": 5000,
                    "pytorch-lightning_version": "2.4.0", "optimizer_states": [], "lr_schedulers": []}, str(path))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert "omegaconf" not in sys.modules
    loaded = OptiSpeech.load_from_checkpoint(str(path), map_location="cpu")
    assert loaded.ckpt_loaded_epoch == 11
    assert isinstance(loaded.train_args, AttrDict) and loaded.train_args.gradient_clip_val == 10 and loaded.train_args.pretraining_steps == 1000
    assert loaded.sample_rate == 22050 and loaded.hop_length == 256 and loaded.inference_args.d_factor == 1.1
    assert isinstance(loaded.text_processor, Placeholder) and loaded.text_processor.languages == ["en-us"]
    assert loaded.generator.loss_coeffs.lambda_align == 5.0
    with pytest.raises(RuntimeError, match="attribute bag"):
        loaded.prepare_input("hello")            # the phonemiser itself is outside this package
    a, b = src.state_dict(), loaded.state_dict()
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    opts, _ = loaded.configure_optimizers()
    assert opts[0].defaults["betas"] == (0.8, 0.99)


def test_decoder_window_geometry_and_join_bookkeeping():
    """Host logic of the second half of round 2: the training forward runs upsampler + decoder on segment + 2 * halo frames
    only when the decoder is a stack of local ConvNeXt blocks (halo = 3 frames per block; a Transformer decoder keeps the
    full-length path), and the discriminators' deferred stream joins nest and clean up after themselves."""
    import torch

    from optispeech_b200.factory import DEFAULT_MODEL, build_generator, transformer_model_config
    from optispeech_b200.model.generator.training import _decoder_halo
    from optispeech_b200.model.vocoder.wavenext.disc import native

    gen = build_generator(DEFAULT_MODEL["generator"] if "generator" in DEFAULT_MODEL else None)
    assert _decoder_halo(gen.decoder) == 3 * len(gen.decoder.convnext) == 12
    tgen = build_generator(transformer_model_config()["generator"] if "generator" in transformer_model_config() else transformer_model_config())
    assert _decoder_halo(tgen.decoder) is None

    # the window mask of training.py: rows outside [0, len_b) are padding, exactly the rows the full-length mask / zero padding cover
    start, lens, halo, S = torch.tensor([0, 131, 85]), torch.tensor([200, 200, 154]), 12, 64
    t_full = start[:, None] - halo + torch.arange(S + 2 * halo)[None, :]
    win_pad = (t_full < 0) | (t_full >= lens[:, None])
    assert win_pad[0, :halo].all() and not win_pad[0, halo:].any()                 # segment at frame 0: the left halo is outside
    assert not win_pad[1, : 200 - 131 + halo].any() and win_pad[1, 200 - 131 + halo:].all()   # runs past Tm = 200
    assert int((~win_pad[2]).sum()) == 154 - (85 - halo)                            # short sample: rows up to its length

    assert native._DEFERRED_JOINS == []
    with native.deferred_join():
        assert native._DEFERRED_JOINS == [None]
        with native.deferred_join():
            assert native._DEFERRED_JOINS == [None, None]
        assert native._DEFERRED_JOINS == [None]
    assert native._DEFERRED_JOINS == []
