"""Developer probe: run one tcgen05 GEMM case on the GPU and compare with torch fp32 matmul.

Usage: python tools/probe_gemm.py <case> ; each case runs in its own process (see tools/run_probes.sh)
so that a device trap in one case cannot poison the others.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optispeech_b200 import _lib  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def ref_conv(a, w, taps, pad):
    # a: (B,T,K) fp32 ; w: (taps,N,K) fp32 -> (B,T,N)
    B, T, K = a.shape
    acc = torch.zeros(B, T, w.shape[1], device=a.device, dtype=torch.float32)
    ap = torch.nn.functional.pad(a, (0, 0, pad, taps - 1 - pad))
    for tap in range(taps):
        acc += ap[:, tap:tap + T, :] @ w[tap].t()
    return acc


def run_nt(name, B, T, K, N, taps, pad, epi, flags=0, timing=True):
    lib = _lib.load()
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(1234)
    a = (torch.randn(B, T, K, generator=g) * 1.0).to(dev).half()
    w = (torch.randn(taps, N, K, generator=g) * 0.05).to(dev).half()
    bias = torch.randn(N, generator=g).to(dev)
    resid = torch.randn(B * T, N, generator=g).to(dev)
    gamma = torch.randn(N, generator=g).to(dev)
    ln_w = torch.randn(N, generator=g).to(dev)
    ln_b = torch.randn(N, generator=g).to(dev)
    dot_w = torch.randn(N, generator=g).to(dev)
    pad_mask = (torch.rand(B * T, generator=g) < 0.2).to(torch.uint8).to(dev)
    out_f32 = torch.full((B * T, N), float("nan"), device=dev)
    out_h16 = torch.full((B * T, N), float("nan"), device=dev, dtype=torch.half)
    aux = torch.full((B * T, N), float("nan"), device=dev, dtype=torch.half)
    out_dot = torch.full((B * T,), float("nan"), device=dev)

    d = _lib.GemmDesc()
    d.a, d.w = a.data_ptr(), w.data_ptr()
    d.lda, d.ldw = K, K
    d.B, d.T, d.N, d.K, d.taps, d.pad = B, T, N, K, taps, pad
    d.epi, d.flags = epi, flags
    f32_out = epi in (_lib.EPI_BIAS, _lib.EPI_RESID, _lib.EPI_BIAS_LN)
    d.out = (out_f32 if f32_out else out_h16).data_ptr()
    d.aux_h16 = aux.data_ptr()
    d.ldo = N
    d.bias = bias.data_ptr()
    d.resid, d.gamma = resid.data_ptr(), gamma.data_ptr()
    d.row_scale = None
    d.pad_mask = pad_mask.data_ptr()
    d.ln_w, d.ln_b, d.ln_eps = ln_w.data_ptr(), ln_b.data_ptr(), 1e-6
    dot_b = torch.full((1,), 0.25, device=dev)
    d.dot_w, d.dot_b, d.out_dot = dot_w.data_ptr(), dot_b.data_ptr(), out_dot.data_ptr()
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.osb_gemm(C.byref(d), C.c_void_p(st))
    torch.cuda.synchronize()
    print(f"[{name}] rc={rc}")
    if rc != 0:
        return False

    acc = ref_conv(a.float(), w.float(), taps, pad).reshape(B * T, N) + bias
    keep = (1 - pad_mask.float())[:, None]
    if epi == _lib.EPI_BIAS:
        ref = acc
        if flags & _lib.FLAG_CLIP:
            ref = ref.clamp(-1, 1)
        if flags & _lib.FLAG_KEEPMASK:
            ref = ref * keep
        got = out_f32
    elif epi == _lib.EPI_GELU:
        ref = torch.nn.functional.gelu(acc)
        got = out_h16.float()
    elif epi == _lib.EPI_RELU:
        ref = acc.relu()
        got = out_h16.float()
    elif epi == _lib.EPI_RESID:
        ref = resid + gamma * acc
        if flags & _lib.FLAG_KEEPMASK:
            ref = ref * keep
        got = out_f32
    elif epi == _lib.EPI_RELU_LN:
        ref = torch.nn.functional.layer_norm(acc.relu(), (N,), ln_w, ln_b, 1e-6)
        got = out_h16.float()
    elif epi == _lib.EPI_BIAS_LN:
        ref = torch.nn.functional.layer_norm(acc, (N,), ln_w, ln_b, 1e-6)
        got = out_f32
    err = (got - ref).abs()
    tol = 2e-2 if not f32_out else 2e-3
    ok = bool(torch.isfinite(got).all()) and float(err.max()) < tol * max(1.0, float(ref.abs().max()))
    print(f"[{name}] max_abs_err={float(err.max()):.3e} ref_absmax={float(ref.abs().max()):.3e} nan={int(torch.isnan(got).sum())} ok={ok}")
    if flags & _lib.FLAG_DOT:
        dref = (ref * dot_w).sum(-1) + 0.25
        dref = torch.where(pad_mask.bool(), torch.zeros_like(dref), dref)
        derr = (out_dot - dref).abs().max()
        print(f"[{name}] dot max_abs_err={float(derr):.3e}")
        ok = ok and float(derr) < 5e-2
    if not ok:
        np.savez(os.path.join(OUT, f"probe_{name}.npz"), got=got[:256].cpu().numpy(), ref=ref[:256].cpu().numpy())
    if timing and ok:
        for _ in range(3):
            lib.osb_gemm(C.byref(d), C.c_void_p(st))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 20
        for _ in range(n):
            lib.osb_gemm(C.byref(d), C.c_void_p(st))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 2.0 * B * T * N * K * taps
        print(f"[{name}] {ms*1e3:.1f} us/launch  {fl/ms/1e9:.1f} TFLOP/s")
    return ok


def run_wgrad(name, B, T, N, K, taps, pad):
    lib = _lib.load()
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(4321)
    dy = torch.randn(B, T, N, generator=g).to(dev).half()
    a = torch.randn(B, T, K, generator=g).to(dev).half()
    dw = torch.zeros(taps, N, K, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.osb_gemm_wgrad(dy.data_ptr(), N, a.data_ptr(), K, dw.data_ptr(), B, T, N, K, taps, pad, C.c_void_p(st))
    torch.cuda.synchronize()
    print(f"[{name}] rc={rc}")
    if rc != 0:
        return False
    ap = torch.nn.functional.pad(a.float(), (0, 0, pad, taps - 1 - pad))
    ref = torch.stack([torch.einsum("btn,btk->nk", dy.float(), ap[:, tap:tap + T]) for tap in range(taps)])
    err = (dw - ref).abs().max()
    ok = float(err) < 2e-3 * float(ref.abs().max())
    print(f"[{name}] max_abs_err={float(err):.3e} ref_absmax={float(ref.abs().max()):.3e} ok={ok}")
    if not ok:
        np.savez(os.path.join(OUT, f"probe_{name}.npz"), got=dw[0].cpu().numpy(), ref=ref[0].cpu().numpy())
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            lib.osb_gemm_wgrad(dy.data_ptr(), N, a.data_ptr(), K, dw.data_ptr(), B, T, N, K, taps, pad, C.c_void_p(st))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"[{name}] {ms*1e3:.1f} us/launch  {2.0*B*T*N*K*taps/ms/1e9:.1f} TFLOP/s")
    return ok


L = _lib
CASES = {
    "nt_tiny": lambda: run_nt("nt_tiny", 1, 128, 64, 64, 1, 0, L.EPI_BIAS),
    "nt_k256": lambda: run_nt("nt_k256", 1, 128, 256, 256, 1, 0, L.EPI_BIAS),
    "nt_gelu": lambda: run_nt("nt_gelu", 2, 200, 256, 1024, 1, 0, L.EPI_GELU, L.FLAG_SAVE_PRE),
    "nt_resid": lambda: run_nt("nt_resid", 2, 200, 1024, 256, 1, 0, L.EPI_RESID, L.FLAG_KEEPMASK),
    "nt_voc1": lambda: run_nt("nt_voc1", 4, 300, 384, 1152, 1, 0, L.EPI_GELU),
    "nt_voc2": lambda: run_nt("nt_voc2", 4, 300, 1152, 384, 1, 0, L.EPI_RESID, L.FLAG_KEEPMASK | L.FLAG_OUT_H16),
    "nt_conv7_ln": lambda: run_nt("nt_conv7_ln", 3, 211, 256, 384, 7, 3, L.EPI_BIAS_LN, L.FLAG_OUT_H16),
    "nt_conv3_reluln": lambda: run_nt("nt_conv3_reluln", 3, 192, 256, 384, 3, 1, L.EPI_RELU_LN, L.FLAG_DOT | L.FLAG_SAVE_PRE),
    "nt_conv5_reluln": lambda: run_nt("nt_conv5_reluln", 3, 192, 256, 256, 5, 2, L.EPI_RELU_LN),
    "nt_big": lambda: run_nt("nt_big", 32, 864, 256, 1024, 1, 0, L.EPI_GELU),
    "nt_big2": lambda: run_nt("nt_big2", 32, 864, 1024, 256, 1, 0, L.EPI_RESID),
    "nt_clip": lambda: run_nt("nt_clip", 2, 100, 384, 256, 1, 0, L.EPI_BIAS, L.FLAG_CLIP),
    "wg_small": lambda: run_wgrad("wg_small", 1, 64, 128, 128, 1, 0),
    "wg_mid": lambda: run_wgrad("wg_mid", 2, 200, 256, 128, 1, 0),
    "wg_conv3": lambda: run_wgrad("wg_conv3", 2, 200, 384, 256, 3, 1),
    "wg_big": lambda: run_wgrad("wg_big", 32, 864, 1024, 256, 1, 0),
}

if __name__ == "__main__":
    name = sys.argv[1]
    ok = CASES[name]()
    sys.exit(0 if ok else 1)
