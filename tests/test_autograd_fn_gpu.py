"""Unit parity of every autograd Function (forward + hand-written backward kernels) against the oracle's fp32
restatement evaluated with torch.autograd on the same inputs.  Tolerances: fp16 operands, fp32 accumulation."""
import pytest
import torch
import torch.nn.functional as F

from oracle import model as O
from oracle.spec import PredictorSpec

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))


def _check(named_pairs, tol):
    bad = [(n, rel(a, b)) for n, a, b in named_pairs if not rel(a, b) <= tol]
    for n, a, b in named_pairs:
        print(f"  {n}: rel err {rel(a, b):.3e}")
    assert not bad, bad


def test_layernorm_fn(cuda_device):
    from optispeech_b200.autograd import LayerNormFn

    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 50, 256, generator=g).to(cuda_device).requires_grad_(True)
    w = (1 + 0.1 * torch.randn(256, generator=g)).to(cuda_device).requires_grad_(True)
    b = (0.1 * torch.randn(256, generator=g)).to(cuda_device).requires_grad_(True)
    dy = torch.randn(3, 50, 256, generator=g).to(cuda_device)
    y = LayerNormFn.apply(x, w, b, 1e-6)
    gx, gw, gb = torch.autograd.grad(y, (x, w, b), dy)
    yr = F.layer_norm(x, (256,), w, b, 1e-6)
    rx, rw, rb = torch.autograd.grad(yr, (x, w, b), dy)
    _check([("y", y, yr), ("dx", gx, rx), ("dw", gw, rw), ("db", gb, rb)], 1e-5)


@pytest.mark.parametrize("C,I,T,nsplit,fused", [(256, 1024, 77, 0, True), (384, 1152, 64, 0, True), (256, 1024, 300, 1, True),
                                                 (256, 1024, 300, 3, True), (384, 1152, 200, 1, True), (256, 1024, 77, 0, False)])
def test_convnext_block_fn(cuda_device, C, I, T, nsplit, fused):
    """ConvNeXtBlockFn forward + every gradient against the oracle block under torch autograd: the fused tcgen05 forward /
    backward kernels (intermediate dimension split over CTAs automatically, not at all, or 3-way) and the three-kernel path."""
    import ctypes

    from optispeech_b200 import _lib
    from optispeech_b200.autograd import ConvNeXtBlockFn

    lib = _lib.load()
    for fn in (lib.osb_debug_set_fused_nsplit, lib.osb_debug_set_bwd_nsplit):
        fn.argtypes = [ctypes.c_int]
        fn.restype = None
        fn(nsplit)
    ConvNeXtBlockFn.FUSED = fused

    g = torch.Generator().manual_seed(1)
    B = 4 if (C == 384 and T == 64) else 3      # even batch of 64-frame segments: two samples share a 128-row tile (pair mode)
    dev = cuda_device
    sd = {
        "b.dwconv.weight": torch.randn(C, 1, 7, generator=g) * 0.3, "b.dwconv.bias": torch.randn(C, generator=g) * 0.1,
        "b.norm.weight": 1 + 0.1 * torch.randn(C, generator=g), "b.norm.bias": 0.1 * torch.randn(C, generator=g),
        "b.pwconv1.weight": torch.randn(I, C, generator=g) / C ** 0.5, "b.pwconv1.bias": 0.1 * torch.randn(I, generator=g),
        "b.pwconv2.weight": torch.randn(C, I, generator=g) / I ** 0.5, "b.pwconv2.bias": 0.1 * torch.randn(C, generator=g),
        "b.gamma": 0.25 * (1 + 0.1 * torch.randn(C, generator=g)),
    }
    sd = {k: v.to(dev).requires_grad_(True) for k, v in sd.items()}
    x = torch.randn(B, T, C, generator=g).to(dev).requires_grad_(True)
    pad = (torch.arange(T)[None] >= torch.tensor([T, T - 9, T // 2, T - 1][:B])[:, None]).to(dev)
    rs = torch.tensor([1.0, 0.0, 1.25, 0.5][:B]).to(dev)
    dy = torch.randn(B, T, C, generator=g).to(dev)
    names = ["dwconv.weight", "dwconv.bias", "norm.weight", "norm.bias", "pwconv1.weight", "pwconv1.bias", "pwconv2.weight",
             "pwconv2.bias", "gamma"]
    params = [sd[f"b.{n}"] for n in names]
    try:
        y = ConvNeXtBlockFn.apply(x, *params, pad.to(torch.uint8), rs, 1e-6)
        grads = torch.autograd.grad(y, (x, *params), dy)
        torch.cuda.synchronize()
    finally:
        ConvNeXtBlockFn.FUSED = True
        lib.osb_debug_set_fused_nsplit(0)
        lib.osb_debug_set_bwd_nsplit(0)
    yr = O.convnext_block(sd, "b", x, drop_scale=rs) * (1 - pad.float())[..., None]
    rgrads = torch.autograd.grad(yr, (x, *params), dy)
    _check([("y", y, yr)] + [(n, a, b) for n, a, b in zip(["dx"] + names, grads, rgrads)], 4e-3)


@pytest.mark.parametrize("L,Cmid,k,needs_dx", [(2, 384, 3, True), (5, 256, 5, True), (2, 384, 3, False)])
def test_variance_predictor_fn(cuda_device, L, Cmid, k, needs_dx):
    from optispeech_b200.autograd import VariancePredictorFn

    g = torch.Generator().manual_seed(2)
    B, T, C = 3, 61, 256
    dev = cuda_device
    ps = PredictorSpec(L, Cmid, k)
    sd = {}
    for i in range(L):
        cin = C if i == 0 else Cmid
        sd[f"p.conv.{i}.0.weight"] = torch.randn(Cmid, cin, k, generator=g) / (cin * k) ** 0.5
        sd[f"p.conv.{i}.0.bias"] = 0.1 * torch.randn(Cmid, generator=g)
        sd[f"p.conv.{i}.2.weight"] = 1 + 0.1 * torch.randn(Cmid, generator=g)
        sd[f"p.conv.{i}.2.bias"] = 0.1 * torch.randn(Cmid, generator=g)
    sd["p.linear.weight"] = torch.randn(1, Cmid, generator=g) / Cmid ** 0.5
    sd["p.linear.bias"] = torch.randn(1, generator=g)
    sd = {k_: v.to(dev).requires_grad_(True) for k_, v in sd.items()}
    x = torch.randn(B, T, C, generator=g).to(dev).requires_grad_(needs_dx)
    pad = (torch.arange(T)[None] >= torch.tensor([T, T - 9, T // 2])[:, None]).to(dev)
    dy = torch.randn(B, T, generator=g).to(dev)
    layer_params = []
    for i in range(L):
        layer_params += [sd[f"p.conv.{i}.0.weight"], sd[f"p.conv.{i}.0.bias"], sd[f"p.conv.{i}.2.weight"], sd[f"p.conv.{i}.2.bias"]]
    y = VariancePredictorFn.apply(x, pad.to(torch.uint8), k, 1e-12, 0.0, 0, sd["p.linear.weight"], sd["p.linear.bias"], *layer_params)
    wrt = ([x] if needs_dx else []) + [sd["p.linear.weight"], sd["p.linear.bias"]] + layer_params
    grads = torch.autograd.grad(y, wrt, dy)
    yr = O.variance_predictor(sd, "p", x, pad, ps)
    rgrads = torch.autograd.grad(yr, wrt, dy)
    names = (["dx"] if needs_dx else []) + ["lin_w", "lin_b"] + [f"l{i}.{n}" for i in range(L) for n in ("cw", "cb", "lnw", "lnb")]
    # ReLU gates that flip under the fp16-operand forward (vs the fp32 oracle) dominate the error and grow with depth
    _check([("y", y, yr)] + list(zip(names, grads, rgrads)), 6e-3 if L <= 2 else 4e-2)


def test_conv_stack_fn(cuda_device):
    from optispeech_b200.autograd import ConvStackFn

    g = torch.Generator().manual_seed(3)
    B, T, Cin, C = 2, 90, 100, 256
    dev = cuda_device
    ws = [torch.randn(C, Cin, 3, generator=g) / (Cin * 3) ** 0.5, torch.randn(C, C, 3, generator=g) / (C * 3) ** 0.5,
          torch.randn(C, C, 1, generator=g) / C ** 0.5]
    bs = [0.1 * torch.randn(C, generator=g) for _ in range(3)]
    ws = [w.to(dev).requires_grad_(True) for w in ws]
    bs = [b.to(dev).requires_grad_(True) for b in bs]
    x = torch.randn(B, T, Cin, generator=g).to(dev)   # the mel input never needs a gradient
    dy = torch.randn(B, T, C, generator=g).to(dev)
    y = ConvStackFn.apply(x, 128, ws[0], bs[0], ws[1], bs[1], ws[2], bs[2])
    grads = torch.autograd.grad(y, (*ws, *bs), dy)
    h = x.transpose(1, 2)
    h = F.relu(F.conv1d(h, ws[0], bs[0], padding=1))
    h = F.relu(F.conv1d(h, ws[1], bs[1], padding=1))
    yr = F.conv1d(h, ws[2], bs[2]).transpose(1, 2)
    rgrads = torch.autograd.grad(yr, (*ws, *bs), dy)
    names = ["w0", "w1", "w2", "b0", "b1", "b2"]
    _check([("y", y, yr)] + list(zip(names, grads, rgrads)), 1.5e-2)
    # text path: 256 -> 256 (k3) -> 256 (k1), input gradient required
    xt = torch.randn(B, T, C, generator=g).to(dev).requires_grad_(True)
    y = ConvStackFn.apply(xt, 0, ws[1], bs[1], ws[2], bs[2])
    grads = torch.autograd.grad(y, (xt, ws[1], ws[2]), dy)
    yr = F.conv1d(F.relu(F.conv1d(xt.transpose(1, 2), ws[1], bs[1], padding=1)), ws[2], bs[2]).transpose(1, 2)
    rgrads = torch.autograd.grad(yr, (xt, ws[1], ws[2]), dy)
    _check([("y", y, yr)] + list(zip(["dx", "w1", "w2"], grads, rgrads)), 1.5e-2)


def test_embed_and_variance_embed_fn(cuda_device):
    from optispeech_b200.autograd import EmbedTextFn, VarianceEmbedFn

    g = torch.Generator().manual_seed(4)
    dev = cuda_device
    B, T, C, V = 3, 40, 256, 50
    ids = torch.randint(0, V, (B, T), generator=g).to(dev)
    table = (0.3 * torch.randn(V, C, generator=g)).to(dev).requires_grad_(True)
    scale = torch.full((1,), 0.08).to(dev).requires_grad_(True)
    inv_freq = (2000.0 ** -(torch.arange(C // 2).float() / (C // 2))).to(dev)
    dy = torch.randn(B, T, C, generator=g).to(dev)
    y = EmbedTextFn.apply(ids, table, scale, inv_freq, 0)
    gt, gs = torch.autograd.grad(y, (table, scale), dy)
    ang = torch.arange(T, device=dev).float()[:, None] * inv_freq[None]
    yr = C ** 0.5 * F.embedding(ids, table, padding_idx=0) + torch.cat((ang.sin(), ang.cos()), -1) * scale
    rt, rs_ = torch.autograd.grad(yr, (table, scale), dy)
    _check([("y", y, yr), ("dtable", gt, rt), ("dscale", gs, rs_)], 1e-5)

    x = torch.randn(B, T, C, generator=g).to(dev).requires_grad_(True)
    val = torch.randn(B, T, generator=g).to(dev)
    w = (0.3 * torch.randn(C, 1, 9, generator=g)).to(dev).requires_grad_(True)
    b = (0.1 * torch.randn(C, generator=g)).to(dev).requires_grad_(True)
    pad = (torch.arange(T)[None] >= torch.tensor([T, T - 9, T // 2])[:, None]).to(dev)
    y = VarianceEmbedFn.apply(x, val, w, b, pad.to(torch.uint8))
    grads = torch.autograd.grad(y, (x, w, b), dy)
    yr = (x + F.conv1d(val.unsqueeze(1), w, b, padding=4).transpose(1, 2)) * (1 - pad.float())[..., None]
    rgrads = torch.autograd.grad(yr, (x, w, b), dy)
    _check([("y", y, yr)] + list(zip(["dx", "dw", "db"], grads, rgrads)), 1e-5)


def test_wavenext_head_fn(cuda_device):
    from optispeech_b200.autograd import WaveNeXtHeadFn

    g = torch.Generator().manual_seed(5)
    dev = cuda_device
    B, T, C = 2, 64, 384
    x = torch.randn(B, T, C, generator=g).to(dev).requires_grad_(True)
    w1 = (torch.randn(1026, C, generator=g) / C ** 0.5).to(dev).requires_grad_(True)
    b1 = (0.1 * torch.randn(1026, generator=g)).to(dev).requires_grad_(True)
    w2 = (torch.randn(256, 1026, generator=g) / 1026 ** 0.5 * 0.5).to(dev).requires_grad_(True)
    dy = torch.randn(B, T * 256, generator=g).to(dev)
    y = WaveNeXtHeadFn.apply(x, w1, b1, w2)
    grads = torch.autograd.grad(y, (x, w1, b1, w2), dy)
    yr = torch.clip(F.linear(F.linear(x, w1, b1), w2).reshape(B, -1), -1.0, 1.0)
    rgrads = torch.autograd.grad(yr, (x, w1, b1, w2), dy)
    # elements whose pre-clip value sits within fp16 operand error of +-1 may flip the clip gate; exclude nothing, use a norm bound
    _check([("y", y, yr)] + list(zip(["dx", "dw1", "db1", "dw2"], grads, rgrads)), 2e-2)


def test_predictor_dropout_mask_is_consistent_between_forward_and_backward(cuda_device):
    """Train-mode dropout is counter-based (no stored mask): the backward pass must regenerate exactly the forward mask.
    The predictor output is linear in the final Linear's weight and affine-linear (to first order) in the LayerNorm biases,
    so directional finite differences of the forward must match <grad, direction> from the hand-written backward."""
    from optispeech_b200.autograd import VariancePredictorFn

    g = torch.Generator().manual_seed(9)
    B, T, C, Cmid, k, L, p, seed = 2, 50, 256, 256, 3, 3, 0.5, 12345
    dev = cuda_device
    params = []
    for i in range(L):
        cin = C if i == 0 else Cmid
        params += [torch.randn(Cmid, cin, k, generator=g) / (cin * k) ** 0.5, 0.1 * torch.randn(Cmid, generator=g),
                   1 + 0.1 * torch.randn(Cmid, generator=g), 0.1 * torch.randn(Cmid, generator=g)]
    params = [t.to(dev).requires_grad_(True) for t in params]
    lin_w = (torch.randn(1, Cmid, generator=g) / Cmid ** 0.5).to(dev).requires_grad_(True)
    lin_b = torch.zeros(1, device=dev, requires_grad=True)
    x = torch.randn(B, T, C, generator=g).to(dev)
    pad = torch.zeros(B, T, dtype=torch.uint8, device=dev)
    dy = torch.randn(B, T, generator=g).to(dev)

    def f(lw=lin_w, ps=params):
        return VariancePredictorFn.apply(x, pad, k, 1e-12, p, seed, lw, lin_b, *ps)

    y = f()
    y_eval = VariancePredictorFn.apply(x, pad, k, 1e-12, 0.0, 0, lin_w, lin_b, *params)
    assert rel(y, y_eval) > 0.1, "dropout had no effect"
    assert torch.equal(y, f()), "same seed must give the same mask"
    grads = torch.autograd.grad(y, [lin_w] + params, dy)
    # exact linearity in lin_w.  The direction comes from the seeded CPU generator (the device generator's state depends on
    # the tests that ran before); the backward recomputes the LayerNorm statistics from the fp16-saved activations, so the
    # two sides agree to the fp16 rounding level (2^-11 = 4.9e-4 per element), not to fp32.
    v = torch.randn(lin_w.shape, generator=g).to(dev)
    fd = ((f(lw=lin_w + v) - y) * dy).sum()
    an = (grads[0] * v).sum()
    print(f"  lin_w directional: fd {float(fd):.5f} analytic {float(an):.5f}")
    assert abs(float(fd - an)) <= 5e-3 * max(1.0, abs(float(an)))


def test_inner_dropout_mask_matches_between_relu_ln_forward_and_backward_epilogues(cuda_device):
    """Read the masks back: forward mask = dropped / undropped LayerNorm output; backward mask = aux / acc of the
    RELU_LN_BWD epilogue driven through an identity weight."""
    from optispeech_b200 import ops

    g = torch.Generator().manual_seed(10)
    dev = cuda_device
    B, T, N, p, seed = 2, 77, 256, 0.3, 987654321
    a = torch.randn(B, T, N, generator=g).to(dev).half()
    w = (torch.randn(1, N, N, generator=g) / N ** 0.5).to(dev).half()
    ln_w, ln_b = torch.ones(N, device=dev), torch.full((N,), 0.5, device=dev)
    bias = torch.zeros(N, device=dev)
    y_d, r, _ = ops.gemm(a, w, epi=ops.EPI_RELU_LN, flags=ops.FLAG_SAVE_PRE, bias=bias, ln_w=ln_w, ln_b=ln_b, ln_eps=1e-12, dropout_p=p,
                         dropout_seed=seed)
    y, _, _ = ops.gemm(a, w, epi=ops.EPI_RELU_LN, bias=bias, ln_w=ln_w, ln_b=ln_b, ln_eps=1e-12)
    ok = y.float().abs() > 1e-2
    mask_f = torch.where(ok, y_d.float() / y.float(), torch.zeros((), device=dev))
    keep_frac = float((mask_f[ok] > 0.5).float().mean())
    print(f"  forward keep fraction {keep_frac:.4f} (expected {1 - p:.4f})")
    assert abs(keep_frac - (1 - p)) < 0.02
    assert torch.allclose(mask_f[ok][mask_f[ok] > 0.5], torch.full((), 1 / (1 - p), device=dev), rtol=5e-3)
    gacc = (torch.randn(B, T, N, generator=g).to(dev) + 3.0).half()        # bounded away from zero
    eye = torch.eye(N, device=dev).half().view(1, N, N).contiguous()
    _, gy, _ = ops.gemm(gacc, eye, epi=ops.EPI_RELU_LN_BWD, flags=ops.FLAG_OUT_H16, aux_in=r, ln_w=ln_w, ln_eps=1e-12, dropout_p=p,
                        dropout_seed=seed)
    mask_b = gy.float() / gacc.float()
    assert torch.equal(mask_b[ok] > 0.5, mask_f[ok] > 0.5), "backward regenerated a different dropout mask"


@pytest.mark.parametrize("B,Tm,Tx", [(3, 150, 64), (2, 333, 41), (2, 400, 192)])
def test_attention_logprob_fn(cuda_device, B, Tm, Tx):
    """Fused pairwise-distance + masked log-softmax + prior (tcgen05 split precision) vs the reference formulation."""
    import numpy as np
    from optispeech_b200 import ops
    from optispeech_b200.autograd import AttnLogProbFn

    g = torch.Generator().manual_seed(6)
    dev = cuda_device
    C = 256
    fe = torch.randn(B, Tm, C, generator=g).to(dev).requires_grad_(True)
    te = torch.randn(B, Tx, C, generator=g).to(dev).requires_grad_(True)
    xl = torch.randint(max(1, Tx // 2), Tx + 1, (B,), generator=g); xl[0] = Tx
    ml = torch.randint(max(xl.max().item(), Tm // 2), Tm + 1, (B,), generator=g); ml[0] = Tm
    lf = torch.from_numpy(np.concatenate([[0.0], np.cumsum(np.log(np.arange(1, Tm + Tx + 4, dtype=np.float64)))])).to(dev)
    prior = ops.beta_binomial_prior(lf, xl.to(dev), ml.to(dev), Tm, Tx)
    ref_prior = torch.full((B, Tm, Tx), -float("inf"))
    for b in range(B):
        ref_prior[b, : ml[b], : xl[b]] = torch.from_numpy(O.beta_binomial_log_prior(int(ml[b]), int(xl[b]))).float()
    assert torch.equal(torch.isinf(prior.cpu()), torch.isinf(ref_prior))
    fin = torch.isfinite(ref_prior)
    assert (prior.cpu()[fin] - ref_prior[fin]).abs().max().item() <= 1e-5
    lp = AttnLogProbFn.apply(fe, te, prior, xl.to(dev), ml.to(dev))
    # reference formulation (alignments.py:66-81)
    dist = torch.norm(fe.unsqueeze(2) - te.unsqueeze(1), p=2, dim=3)
    x_mask = (torch.arange(Tx)[None] >= xl[:, None]).to(dev)
    ref = F.log_softmax((-dist).masked_fill(x_mask.unsqueeze(-2), -float("inf")), dim=-1) + prior
    fin = torch.isfinite(ref)
    assert torch.equal(torch.isfinite(lp), fin)
    err = (lp[fin] - ref[fin]).abs().max().item()
    print(f"  log_p_attn max-abs err {err:.3e}")
    assert err <= 2e-4
    G = torch.randn(B, Tm, Tx, generator=g).to(dev) * fin
    gf, ge = torch.autograd.grad(lp, (fe, te), G)
    ref0 = torch.where(fin, ref, torch.zeros((), device=dev))
    rf, re_ = torch.autograd.grad(ref0, (fe, te), G)
    _check([("dF", gf, rf), ("dE", ge, re_)], 3e-3)


def test_forward_sum_loss_fn(cuda_device):
    """CTC forward-sum kernel (loss + gradient in one launch) vs the oracle's per-sample F.ctc_loss formulation."""
    from optispeech_b200.autograd import ForwardSumLossFn

    g = torch.Generator().manual_seed(8)
    dev = cuda_device
    B, Tm, Tx = 4, 120, 37
    tl = torch.tensor([37, 20, 1, 30])
    fl = torch.tensor([120, 64, 5, 30])
    lp = torch.log_softmax(torch.randn(B, Tm, Tx, generator=g) * 2, dim=-1)
    for b in range(B):
        lp[b, fl[b]:, :] = -float("inf")
        lp[b, :, tl[b]:] = -float("inf")
    ref_in = lp.clone().requires_grad_(True)
    ref = O.forward_sum_loss(ref_in, tl, fl)
    (rg,) = torch.autograd.grad(ref, ref_in)
    x = lp.to(dev).requires_grad_(True)
    loss = ForwardSumLossFn.apply(x, tl.to(dev), fl.to(dev), -1.0)
    (gg,) = torch.autograd.grad(loss, x)
    print(f"  loss cuda {float(loss):.6f} oracle {float(ref):.6f}")
    assert abs(float(loss) - float(ref)) <= 1e-4 * abs(float(ref))
    rg = torch.nan_to_num(rg, nan=0.0)
    _check([("dlogp", gg.cpu(), rg)], 1e-3)
    # infeasible alignment (more tokens than frames): zero_infinity semantics
    tl2, fl2 = torch.tensor([10]), torch.tensor([4])
    lp2 = torch.log_softmax(torch.randn(1, 12, 10, generator=g), dim=-1)
    l2 = ForwardSumLossFn.apply(lp2.to(dev).requires_grad_(True), tl2.to(dev), fl2.to(dev), -1.0)
    assert float(l2) == 0.0


def test_forward_sum_long_text_takes_the_sequential_fallback(cuda_device):
    """More than 511 tokens: the alpha and beta chains no longer fit one CTA side by side and are walked one after the other."""
    from optispeech_b200.autograd import ForwardSumLossFn

    g = torch.Generator().manual_seed(12)
    dev = cuda_device
    B, Tm, Tx = 2, 640, 530
    tl, fl = torch.tensor([530, 300]), torch.tensor([640, 333])
    lp = torch.log_softmax(torch.randn(B, Tm, Tx, generator=g) * 2, dim=-1)
    for b in range(B):
        lp[b, fl[b]:, :] = -float("inf")
        lp[b, :, tl[b]:] = -float("inf")
    ref_in = lp.clone().requires_grad_(True)
    ref = O.forward_sum_loss(ref_in, tl, fl)
    (rg,) = torch.autograd.grad(ref, ref_in)
    x = lp.to(dev).requires_grad_(True)
    loss = ForwardSumLossFn.apply(x, tl.to(dev), fl.to(dev), -1.0)
    (gg,) = torch.autograd.grad(loss, x)
    assert abs(float(loss) - float(ref)) <= 1e-4 * abs(float(ref))
    _check([("dlogp", gg.cpu(), torch.nan_to_num(rg, nan=0.0))], 1e-3)


@pytest.mark.parametrize("Tx,Tm,sharp", [(192, 864, 2.0), (192, 864, 12.0), (300, 700, 3.0), (100, 333, 6.0)])
def test_forward_sum_warp_recursion_matches_oracle_and_legacy(cuda_device, Tx, Tm, sharp):
    """The warp-synchronous linear-domain recursion (registers + shuffles, exact power-of-two rescaling per frame) against the
    oracle's F.ctc_loss formulation and against the shared-memory log-domain kernel it replaces, at the training shape and with
    sharply peaked rows (per-frame probabilities down to e^-60: the rescaling has to carry the range)."""
    import ctypes

    from optispeech_b200 import _lib, ops

    g = torch.Generator().manual_seed(21)
    dev = cuda_device
    B = 3
    tl = torch.tensor([Tx, Tx // 2 + 3, 1])
    fl = torch.tensor([Tm, Tm // 2 + 11, 7])
    lp = torch.log_softmax(torch.randn(B, Tm, Tx, generator=g) * sharp, dim=-1)
    for b in range(B):
        lp[b, fl[b]:, :] = -float("inf")
        lp[b, :, tl[b]:] = -float("inf")
    ref_in = lp.clone().requires_grad_(True)
    ref = O.forward_sum_loss(ref_in, tl, fl)
    (rg,) = torch.autograd.grad(ref, ref_in)
    rg = torch.nan_to_num(rg, nan=0.0)
    x = lp.to(dev)
    loss_w, grad_w = ops.forward_sum(x, tl.to(dev), fl.to(dev), -1.0)
    lib = _lib.load()
    lib.osb_debug_forward_sum_legacy.argtypes = [ctypes.c_int]
    lib.osb_debug_forward_sum_legacy(1)
    try:
        loss_l, grad_l = ops.forward_sum(x, tl.to(dev), fl.to(dev), -1.0)
    finally:
        lib.osb_debug_forward_sum_legacy(0)
    total = float(loss_w.sum()) / B
    print(f"  Tx={Tx} Tm={Tm} sharp={sharp}: warp {total:.6f} legacy {float(loss_l.sum()) / B:.6f} oracle {float(ref):.6f}")
    assert abs(total - float(ref)) <= 2e-5 * abs(float(ref))
    assert torch.allclose(loss_w, loss_l, rtol=2e-5, atol=1e-6)
    # sharply peaked rows put the posteriors at exp() of differences of ~1e3-magnitude logs: 3e-3 there, 1e-3 otherwise
    _check([("dlogp vs oracle", grad_w.cpu(), rg), ("dlogp vs legacy", grad_w.cpu(), grad_l.cpu())], 3e-3 if sharp > 8 else 1e-3)


@pytest.mark.parametrize("B,T_in,Cin,Cout,stride", [(3, 200, 128, 128, 3), (2, 1366, 32, 128, 3), (2, 97, 64, 64, 2)])
def test_strided_conv_gemm_with_leaky_relu(cuda_device, B, T_in, Cin, Cout, stride):
    """Conv1d(k=5, stride, padding=2) + LeakyReLU(0.1) as an implicit GEMM whose row stride is a TMA traversal stride (the shape
    of the period discriminators' (5,1)/stride-3 convolutions, reference disc/_discriminators.py:52-60), against F.conv1d."""
    from optispeech_b200 import ops

    g = torch.Generator().manual_seed(B * T_in)
    dev = cuda_device
    k, pad = 5, 2
    x = torch.randn(B, T_in, Cin, generator=g).to(dev).half()
    w = (torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5)
    bias = 0.1 * torch.randn(Cout, generator=g)
    wp = ops.pack_conv_h16(w.to(dev))                                              # (k, Cout, Cin)
    out, _, _ = ops.gemm(x, wp, epi=ops.EPI_BIAS, pad=pad, bias=bias.to(dev), row_stride=stride, lrelu=0.1)
    ref = F.leaky_relu(F.conv1d(x.float().cpu().transpose(1, 2), w.half().float(), bias, stride=stride, padding=pad), 0.1).transpose(1, 2)
    assert out.shape == ref.shape, (out.shape, ref.shape)
    err = float((out.cpu() - ref).abs().max())
    print(f"  stride {stride} T_in {T_in} -> {out.shape[1]} rows: max-abs err {err:.3e}")
    assert err <= 2e-3
