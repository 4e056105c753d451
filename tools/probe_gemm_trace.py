"""Developer probe: where does a small gemm_nt launch spend its time?  Back-to-back timing of the step's typical shapes
plus a clock64 timeline of CTA (0,0): entry, set-up done, first TMA issued, last TMA issued, first stage landed, MMAs
issued, accumulator ready, epilogue done, exit."""
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


def run(name, K, N, taps, fn):
    a = torch.randn(B, T, K, generator=g).to(dev).half()
    w = (torch.randn(taps, N, K, generator=g) / (K * taps) ** 0.5).to(dev).half()
    for _ in range(3):
        fn(a, w)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        fn(a, w)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    tr = torch.zeros(16, dtype=torch.int64, device=dev)
    lib.osb_debug_set_gemm_trace(C.c_void_p(tr.data_ptr()))
    fn(a, w)
    torch.cuda.synchronize()
    lib.osb_debug_set_gemm_trace(None)
    t = tr.cpu().tolist()
    rel = [(v - t[0]) if v else None for v in t[:9]]
    names = ["entry", "setup", "tma0", "tmaN", "stage0", "mma_done_issue", "epi_vectors_staged", "epi_done", "exit"]
    print(f"{name:34s} {us:7.1f} us back-to-back | cycles: " + " ".join(f"{n}={v}" for n, v in zip(names, rel)))


bias256 = torch.zeros(256, device=dev)
bias1024 = torch.zeros(1024, device=dev)
ones = torch.ones(256, device=dev)
resid = torch.randn(B, T, 256, generator=g).to(dev)
lnw, lnb = torch.ones(256, device=dev), torch.zeros(256, device=dev)
run("BIAS   N=256 K=256 taps=1", 256, 256, 1, lambda a, w: ops.gemm(a, w, epi=ops.EPI_BIAS, bias=bias256))
run("BIAS   N=256 K=256 taps=5", 256, 256, 5, lambda a, w: ops.gemm(a, w, epi=ops.EPI_BIAS, bias=bias256, pad=2))
run("RELU   N=256 K=256 taps=3", 256, 256, 3, lambda a, w: ops.gemm(a, w, epi=ops.EPI_RELU, bias=bias256, pad=1))
run("GELU   N=1024 K=256 taps=1", 256, 1024, 1, lambda a, w: ops.gemm(a, w, epi=ops.EPI_GELU, bias=bias1024))
run("RESID  N=256 K=1024 taps=1", 1024, 256, 1, lambda a, w: ops.gemm(a, w, epi=ops.EPI_RESID, bias=bias256, resid=resid, gamma=ones))
run("RELU_LN N=256 K=256 taps=5", 256, 256, 5, lambda a, w: ops.gemm(a, w, epi=ops.EPI_RELU_LN, bias=bias256, pad=2, ln_w=lnw, ln_b=lnb, ln_eps=1e-12))
# the same output with an empty kernel-equivalent: a plain fp32 copy of the output size, for the launch floor
x = torch.randn(B, T, 256, device=dev)
y = torch.empty_like(x)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    y.copy_(x)
e1.record()
torch.cuda.synchronize()
print(f"torch copy of the (6144, 256) fp32 output: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us back-to-back")
