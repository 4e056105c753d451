"""Developer probe: device time per launch of the step's typical small GEMMs inside a CUDA graph (20 dependent launches on one
stream, replayed) — free of host launch cost — next to the in-kernel clock64 span of CTA (0,0)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optispeech_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
lib.osb_debug_set_gemm_trace.argtypes = [C.c_void_p]
B, T = 32, 192
g = torch.Generator().manual_seed(0)


def run(name, K, N, taps, fn, n=20):
    a = torch.randn(B, T, K, generator=g).to(dev).half()
    w = (torch.randn(taps, N, K, generator=g) / (K * taps) ** 0.5).to(dev).half()
    for _ in range(3):
        fn(a, w)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(n):
            fn(a, w)
    for _ in range(3):
        graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (reps * n) * 1e3
    tr = torch.zeros(16, dtype=torch.int64, device=dev)
    lib.osb_debug_set_gemm_trace(C.c_void_p(tr.data_ptr()))
    fn(a, w)
    torch.cuda.synchronize()
    lib.osb_debug_set_gemm_trace(None)
    t = tr.cpu().tolist()
    flops = 2.0 * B * T * N * K * taps
    print(f"{name:34s} {us:7.2f} us/launch in a graph ({flops / us / 1e6:6.1f} TFLOP/s) | CTA(0,0) in-kernel {(t[8] - t[0]) / 1.965e3:6.2f} us")


bias256 = torch.zeros(256, device=dev)
bias384 = torch.zeros(384, device=dev)
bias1024 = torch.zeros(1024, device=dev)
ones = torch.ones(256, device=dev)
resid = torch.randn(B, T, 256, generator=g).to(dev)
lnw, lnb = torch.ones(256, device=dev), torch.zeros(256, device=dev)
lnw3, lnb3 = torch.ones(384, device=dev), torch.zeros(384, device=dev)
run("BIAS   N=256 K=256 taps=1", 256, 256, 1, lambda a, w: ops.gemm(a, w, epi=ops.EPI_BIAS, bias=bias256))
run("BIAS   N=256 K=256 taps=5", 256, 256, 5, lambda a, w: ops.gemm(a, w, epi=ops.EPI_BIAS, bias=bias256, pad=2))
run("GELU   N=1024 K=256 taps=1", 256, 1024, 1, lambda a, w: ops.gemm(a, w, epi=ops.EPI_GELU, bias=bias1024))
run("RESID  N=256 K=1024 taps=1", 1024, 256, 1, lambda a, w: ops.gemm(a, w, epi=ops.EPI_RESID, bias=bias256, resid=resid, gamma=ones))
run("RELU_LN N=256 K=256 taps=5", 256, 256, 5, lambda a, w: ops.gemm(a, w, epi=ops.EPI_RELU_LN, bias=bias256, pad=2, ln_w=lnw, ln_b=lnb, ln_eps=1e-12))
run("RELU_LN N=384 K=384 taps=3", 384, 384, 3, lambda a, w: ops.gemm(a, w, epi=ops.EPI_RELU_LN, bias=bias384, pad=1, ln_w=lnw3, ln_b=lnb3, ln_eps=1e-12))
