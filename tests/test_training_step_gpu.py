"""The public training surface on the GPU: OptiSpeech.training_step (reference base_lightning_module.py:78-126)
in the generator pre-training phase and in the GAN phase, flat-bucket AdamW against torch.optim.AdamW."""
import numpy as np
import pytest
import torch

from oracle.spec import ModelSpec, deterministic_state_dict, generator_shapes

pytestmark = pytest.mark.gpu


def _batch(spec, B, Tx, Tm, seed=3):
    g = torch.Generator().manual_seed(seed)
    xl = torch.randint(Tx // 2, Tx + 1, (B,), generator=g); xl[0] = Tx
    x = torch.randint(1, 159, (B, Tx), generator=g) * (torch.arange(Tx)[None] < xl[:, None])
    ml = torch.clamp((xl.float() * (Tm / Tx)).round().long(), max=Tm); ml[0] = Tm
    mm = torch.arange(Tm)[None] < ml[:, None]
    return dict(x=x, x_lengths=xl, mel=torch.randn(B, spec.n_feats, Tm, generator=g) * mm[:, None, :], mel_lengths=ml,
                pitches=torch.randn(B, Tm, generator=g) * mm, energies=torch.randn(B, Tm, generator=g) * mm,
                wav=(torch.rand(B, Tm * spec.hop_length, generator=g) * 2 - 1).numpy().astype(np.float32), sids=None, lids=None)


def test_flat_adamw_matches_torch_adamw(cuda_device):
    from optispeech_b200.optim import FlatAdamW

    g = torch.Generator().manual_seed(0)
    shapes = [(256, 1024), (1024,), (7,), (33, 5)]
    ps = [torch.randn(s, generator=g).to(cuda_device) for s in shapes]
    a = [torch.nn.Parameter(p.clone()) for p in ps]
    b = [torch.nn.Parameter(p.clone()) for p in ps]
    unused = torch.nn.Parameter(torch.ones(5, device=cuda_device))   # never gets a gradient: must stay untouched
    oa = FlatAdamW([{"params": a + [unused]}], lr=2e-4, betas=(0.8, 0.99), weight_decay=1e-2, max_grad_norm=10.0, loss_scale=512.0)
    ob = torch.optim.AdamW(b, lr=2e-4, betas=(0.8, 0.99), weight_decay=1e-2)
    for it in range(5):
        grads = [torch.randn(s, generator=g).to(cuda_device) * (50.0 if it == 2 else 1.0) for s in shapes]
        oa.zero_grad(); ob.zero_grad()
        for p, q, gr in zip(a, b, grads):
            if p.grad is None:
                p.grad = (gr * 512.0).clone()
            else:
                p.grad.add_(gr * 512.0)
            q.grad = gr.clone()
        torch.nn.utils.clip_grad_norm_(b, 10.0)
        oa.step(); ob.step()
        for p, q in zip(a, b):
            assert torch.allclose(p, q, rtol=1e-5, atol=1e-6), it
    assert torch.equal(unused.detach(), torch.ones(5, device=cuda_device))
    # non-finite gradient: the step is skipped
    before = [p.detach().clone() for p in a]
    oa.zero_grad()
    for p in a:
        p.grad = torch.ones_like(p)
    a[0].grad.fill_(float("inf"))
    oa.step()
    for p, q in zip(a, before):
        assert torch.equal(p.detach(), q)


def test_training_step_pretraining_and_gan_phase(cuda_device):
    from optispeech_b200.factory import build_model, model_config_from_spec

    spec = ModelSpec()
    torch.manual_seed(1234)
    model = build_model(model_config_from_spec(spec), train_args=dict(pretraining_steps=2))
    model.generator.load_state_dict(deterministic_state_dict(generator_shapes(spec), seed=0, frames_per_token=3.0))
    model = model.to(cuda_device).train()
    batch = _batch(spec, 2, 40, 170)
    dec_before = model.generator.decoder.convnext[0].pwconv1.weight.detach().clone()
    enc_before = model.generator.encoder.convnext[0].pwconv1.weight.detach().clone()
    voc_before = model.generator.vocoder.backbone.convnext[0].pwconv1.weight.detach().clone()
    losses = []
    for i in range(4):   # steps 0,1: generator pre-training; steps 2,3: GAN phase (discriminators on stock PyTorch)
        model.training_step(batch, i)
        losses.append(float(model.logged["total_loss/generator"]))
        assert np.isfinite(losses[-1])
        if i == 1:
            assert torch.equal(voc_before, model.generator.vocoder.backbone.convnext[0].pwconv1.weight.detach()), \
                "vocoder must not move during pre-training"
    print("generator losses:", losses)
    assert "total_loss/discriminator" in model.logged and np.isfinite(float(model.logged["total_loss/discriminator"]))
    assert not torch.equal(enc_before, model.generator.encoder.convnext[0].pwconv1.weight.detach())
    assert not torch.equal(voc_before, model.generator.vocoder.backbone.convnext[0].pwconv1.weight.detach())
    # the decoder never receives a gradient at the reference commit (vocoder input is detached) -> untouched, no weight decay
    assert torch.equal(dec_before, model.generator.decoder.convnext[0].pwconv1.weight.detach())
    g_opt, d_opt = model.optimizers()
    assert np.isfinite(g_opt.grad_norm()) and np.isfinite(d_opt.grad_norm())


def _no_dropout(model):
    """Switch every stochastic layer off so that eager and graphed runs are comparable step by step."""
    from optispeech_b200.model.generator.modules.convnext import ConvNeXtBlock

    for mod in model.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, ConvNeXtBlock):
            mod.drop_path = torch.nn.Identity()


def test_graphed_training_step_matches_eager(cuda_device):
    """training_step with cuda_graph=True (3 eager warm-up steps, capture, replays) walks the same trajectory as the
    eager step: parameters, Adam step counters, LR schedule and global_step all agree after 7 steps."""
    from optispeech_b200.factory import build_model, model_config_from_spec

    spec = ModelSpec()
    models = []
    for graphed in (False, True):
        torch.manual_seed(99)
        m = build_model(model_config_from_spec(spec), train_args=dict(pretraining_steps=1000))
        m.generator.load_state_dict(deterministic_state_dict(generator_shapes(spec), seed=0, frames_per_token=3.0))
        m = m.to(cuda_device).train()
        _no_dropout(m)
        m.cuda_graph = graphed
        models.append(m)
    batches = [_batch(spec, 2, 40, 170, seed=s) for s in (3, 4, 5)]
    for i in range(7):
        for m in models:
            m.training_step(batches[i % 3], i)
        le, lg = (float(m.logged["total_loss/generator"]) for m in models)
        print(f"  step {i}: eager loss {le:.5f} graphed loss {lg:.5f}")
        assert abs(le - lg) <= 2e-3 * max(1.0, abs(le)), i
    eager, graphed = models
    assert graphed._graphed is not None and graphed._graphed.replays == 4 and graphed._graphed.last_entry.launches > 50
    assert eager.global_step == graphed.global_step == 7
    oe, og = eager.optimizers()[0], graphed.optimizers()[0]
    assert oe._steps == og._steps
    assert oe.param_groups[0]["lr"] == og.param_groups[0]["lr"]
    worst = 0.0
    for (n, p), (_, q) in zip(eager.generator.named_parameters(), graphed.generator.named_parameters()):
        worst = max(worst, float((p - q).abs().max()))
    print("  max |param_eager - param_graphed| after 7 steps:", worst)
    assert worst < 2e-4   # fp32 atomics in the weight-gradient kernels reorder sums; Adam normalises the scale away


def test_graph_replay_draws_fresh_dropout_masks(cuda_device):
    """The dropout seed of a captured launch is host seed + a device counter: bumping the counter inside the graph gives
    a new mask on every replay, and the backward epilogue regenerates the same mask as the forward one."""
    from optispeech_b200 import ops

    dev = cuda_device
    g = torch.Generator().manual_seed(5)
    B, T, N, p = 2, 64, 256, 0.4
    a = torch.randn(B, T, N, generator=g).to(dev).half()
    w = (torch.randn(1, N, N, generator=g) / N ** 0.5).to(dev).half()
    ln_w, ln_b, bias = torch.ones(N, device=dev), torch.full((N,), 0.5, device=dev), torch.zeros(N, device=dev)
    y_ref, _, _ = ops.gemm(a, w, epi=ops.EPI_RELU_LN, bias=bias, ln_w=ln_w, ln_b=ln_b, ln_eps=1e-12)
    y_out = torch.empty_like(y_ref)
    ops.gemm(a, w, epi=ops.EPI_RELU_LN, bias=bias, ln_w=ln_w, ln_b=ln_b, ln_eps=1e-12, dropout_p=p, dropout_seed=17, out=y_out)  # warm
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ops.step_counter(dev).add_(1)
        ops.gemm(a, w, epi=ops.EPI_RELU_LN, bias=bias, ln_w=ln_w, ln_b=ln_b, ln_eps=1e-12, dropout_p=p, dropout_seed=17, out=y_out)
    masks = []
    for _ in range(3):
        graph.replay()
        torch.cuda.synchronize()
        masks.append((y_out.float().abs() > 0) | (y_ref.float().abs() < 1e-2))
    assert not torch.equal(masks[0], masks[1]) and not torch.equal(masks[1], masks[2])
    for mk in masks:
        assert abs(float(mk.float().mean()) - (1 - p)) < 0.05


def test_gan_step_overlapped_schedule_matches_serial_schedule(cuda_device):
    """The GAN step runs the eight discriminators on their own streams and queues the discriminator turn next to the
    generator's backward pass (disc/native.fan_out, BaseModule.OVERLAP_TURNS), and the real signals' discriminator pass starts
    before the generator forward (PREFETCH_REAL).  Scheduling must not change the numbers: the
    gradient every generator / discriminator parameter receives equals that of the fully serial schedule up to the
    run-to-run noise of one gradient evaluation (fp32 atomics in the weight-gradient kernels, fp16 operands), eagerly after
    one step and through a captured graph after five."""
    from optispeech_b200.factory import build_model, model_config_from_spec
    from optispeech_b200.model import base_module
    from optispeech_b200.model.vocoder.wavenext.disc import native

    spec = ModelSpec()
    batch = _batch(spec, 2, 40, 170)

    def run(parallel: bool, graph: bool, steps: int):
        torch.manual_seed(1234)
        model = build_model(model_config_from_spec(spec), train_args=dict(pretraining_steps=0))
        model.generator.load_state_dict(deterministic_state_dict(generator_shapes(spec), seed=0, frames_per_token=3.0))
        model = model.to(cuda_device).eval()       # eval: no dropout / DropPath draws, the schedules see the same function
        model.cuda_graph = graph
        native.PARALLEL_DISCRIMINATORS = parallel
        base_module.OVERLAP_TURNS = parallel
        base_module.PREFETCH_REAL = parallel
        try:
            for i in range(steps):                 # graph mode: 3 eager warm-ups, the capture, one replay
                model.training_step(batch, i)
            torch.cuda.synchronize()
        finally:
            native.PARALLEL_DISCRIMINATORS = True
            base_module.OVERLAP_TURNS = True
            base_module.PREFETCH_REAL = True
        grads = {n: p.grad.detach().float().clone() for n, p in model.named_parameters() if p.grad is not None}
        losses = (float(model.logged["total_loss/generator"]), float(model.logged["total_loss/discriminator"]))
        if model._graphed is not None:
            assert model._graphed.replays >= 1
            model._graphed.release()
        return grads, losses

    def group_stats(a, b, prefix):
        """Relative error of the concatenated gradient of one optimizer's parameters, and the worst single tensor among those
        that carry a non-negligible share of it (tensors whose true gradient is ~0 — a bias in front of a LayerNorm, the
        conv_post biases of the discriminator turn — are pure rounding noise and say nothing about scheduling)."""
        names = [n for n in b if n.startswith(prefix)]
        tot = float(torch.sqrt(sum((b[n] ** 2).sum() for n in names)))
        err = float(torch.sqrt(sum(((a[n] - b[n]) ** 2).sum() for n in names))) / tot
        worst = ("", 0.0)
        for n in names:
            nb = float(b[n].norm())
            if nb >= 1e-2 * tot:
                e = float((a[n] - b[n]).norm()) / nb
                if e > worst[1]:
                    worst = (n, e)
        return err, worst

    # The forward pass itself is not bit-reproducible across schedules (partial sums of the fused blocks and of the weight
    # gradients meet in L2 through fp32 reds whose order follows CTA timing); an ulp there flips fp16 roundings and ReLU gates of
    # the 5-layer pitch predictor, whose first-layer gradient then moves by 2-4 % ("Stated tolerances").  A stream race
    # (a kernel reading a buffer another stream has not finished, or has already recycled) leaves O(1) errors in whole tensors.
    for graph, steps in ((False, 1), (True, 5)):
        ref, ref_losses = run(False, False, steps)
        again, _ = run(False, False, steps)          # the serial schedule twice: the noise floor
        got, losses = run(True, graph, steps)
        assert np.allclose(losses, ref_losses, rtol=5e-3), (graph, losses, ref_losses)
        assert set(got) == set(ref) and len(ref) > 150, (len(got), len(ref))
        for prefix in ("generator.", "discriminator."):
            err, worst = group_stats(got, ref, prefix)
            noise, worst_noise = group_stats(again, ref, prefix)
            print(f"graph={graph}, {steps} step(s), {prefix}* gradient: overlapped vs serial {err:.2e} (worst large tensor {worst[1]:.2e} "
                  f"{worst[0]}); serial vs serial {noise:.2e} (worst {worst_noise[1]:.2e})")
            # The discriminator gradients are well conditioned (0.2-0.3 % run to run): a tight bound.  The generator's GAN-phase
            # gradient is not — the feature-matching term's sign(real - fake) flips wherever two feature values nearly agree, so
            # the same schedule run twice already differs by ~5 % after one step and ~20 % after five: bounded by that noise.
            floor = 2e-2 if prefix == "discriminator." else (0.15 if steps == 1 else 0.4)
            assert err <= max(floor, 3.0 * noise), (graph, prefix, err, noise)
            assert worst[1] <= max(4.0 * floor, 4.0 * worst_noise[1]), (graph, prefix, worst, worst_noise)
