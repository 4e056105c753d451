"""Developer probe (round 2): isolated timings of the training-path kernels at the benchmark shapes."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optispeech_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()
for fn in (lib.osb_debug_set_fused_nsplit, lib.osb_debug_set_bwd_nsplit, lib.osb_debug_forward_sum_legacy):
    fn.argtypes = [C.c_int]
    fn.restype = None


def timeit(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


which = set(sys.argv[1:]) or {"block", "fs", "timeline"}
torch.manual_seed(0)
if "block" in which:
    for (Cc, I, B, T) in [(256, 1024, 32, 192), (256, 1024, 32, 864), (384, 1152, 32, 64)]:
        x = torch.randn(B, T, Cc, device=dev)
        dw_w, dw_b = torch.randn(Cc, 7, device=dev) * 0.3, torch.randn(Cc, device=dev) * 0.1
        w1 = (torch.randn(I, Cc, device=dev) / Cc ** 0.5).half().contiguous()
        w2 = (torch.randn(Cc, I, device=dev) / I ** 0.5).half().contiguous()
        b1, b2 = torch.randn(I, device=dev) * 0.1, torch.randn(Cc, device=dev) * 0.1
        gamma = torch.full((Cc,), 0.25, device=dev)
        fl = 2.0 * B * T * 2 * Cc * I
        for ns in (0, 1, 2, 3, 4):
            lib.osb_debug_set_fused_nsplit(ns)
            lib.osb_debug_set_bwd_nsplit(ns)
            us_i = timeit(lambda: ops.convnext_block_fwd(x, dw_w, dw_b, w1, b1, w2, b2, gamma))
            us_f = timeit(lambda: ops.convnext_block_fwd_train(x, dw_w, dw_b, w1, b1, w2, b2, gamma))
            out, xhat, rstd, pre, h = ops.convnext_block_fwd_train(x, dw_w, dw_b, w1, b1, w2, b2, gamma)
            dout = torch.randn_like(x)
            us_b = timeit(lambda: ops.convnext_block_bwd(dout, gamma, None, None, pre, w2, w1))
            print(f"C={Cc} rows={B*T} nsplit={ns}: fwd(infer) {us_i:7.1f} us {fl/us_i/1e6:6.1f} TF | fwd_train {us_f:7.1f} us {fl/us_f/1e6:6.1f} TF | "
                  f"bwd {us_b:7.1f} us {fl/us_b/1e6:6.1f} TF")
        lib.osb_debug_set_fused_nsplit(0)
        lib.osb_debug_set_bwd_nsplit(0)
        dyg, dh, dxh = ops.convnext_block_bwd(dout, gamma, None, None, pre, w2, w1)
        us = timeit(lambda: ops.ln_dwconv_bwd(dxh, xhat, rstd, dout, x, dw_w, None))
        us2 = timeit(lambda: ops.resid_param_grad(dout, out, x, gamma, None, None))
        dw2 = torch.zeros(1, Cc, I, device=dev)
        us3 = timeit(lambda: ops.gemm_wgrad(dyg, h, dw2))
        us4 = timeit(lambda: ops.colsum_h16(dh))
        print(f"   ln_dwconv_bwd {us:6.1f} us | resid_param_grad {us2:6.1f} us | wgrad(dW2) {us3:6.1f} us {2.0*B*T*Cc*I/us3/1e6:6.1f} TF | colsum {us4:6.1f} us")
if "fs" in which:
    B, Tm, Tx = 32, 864, 192
    lp = torch.log_softmax(torch.randn(B, Tm, Tx, device=dev) * 2, dim=-1)
    xl = torch.randint(96, 193, (B,), device=dev); xl[0] = 192
    ml = torch.clamp((4.5 * xl.float()).round().long(), max=864); ml[0] = 864
    for legacy in (0, 1):
        lib.osb_debug_forward_sum_legacy(legacy)
        us = timeit(lambda: ops.forward_sum(lp, xl, ml, -1.0), n=10, warm=2)
        print(f"forward_sum legacy={legacy}: {us:7.1f} us (lse + recursion + grad)")
    lib.osb_debug_forward_sum_legacy(0)
    us = timeit(lambda: ops.mas(lp, xl, ml), n=10, warm=2)
    print(f"mas: {us:7.1f} us")
if "timeline" in which:
    from optispeech_b200.model.generator.modules import ConvNeXtBlock

    lib.osb_debug_set_fused_trace.argtypes = [C.c_void_p]
    for (Cc, I, B, T, ns) in [(256, 1024, 32, 864, 0), (256, 1024, 32, 192, 3)]:
        torch.manual_seed(0)
        blk = ConvNeXtBlock(Cc, I, 0.0, 0.25).to(dev).eval()
        x = torch.randn(B, T, Cc, device=dev)
        tr = torch.zeros(3 * 256, dtype=torch.int64, device=dev)
        lib.osb_debug_set_fused_nsplit(ns)
        with torch.no_grad():
            blk.forward_cl(x, None, split=False)
            lib.osb_debug_set_fused_trace(C.c_void_p(tr.data_ptr()))
            blk.forward_cl(x, None, split=False)
            torch.cuda.synchronize()
            lib.osb_debug_set_fused_trace(None)
        lib.osb_debug_set_fused_nsplit(0)
        t = tr.cpu().view(3, 256)
        base = int(t[1, 0])
        rel = lambda v: int(v) - base if int(v) else None
        print(f"--- timeline C={Cc} I={I} rows={B*T} nsplit={ns} (cycles since MMA thread start)")
        print("mma: a_ready", rel(t[1, 1]), "| worker0: prologue done", rel(t[2, 0]), " acc2_full", rel(t[2, 200]), " end", rel(t[2, 201]))
        for j in range(0, min(6, I // 64)):
            print(f"chunk {j}: prod w1_empty {rel(t[0, 2*j])} w2_empty {rel(t[0, 2*j+1])} | mma w1_full {rel(t[1, 2+6*j])} acc1_empty {rel(t[1, 3+6*j])} "
                  f"g1_issued {rel(t[1, 4+6*j])} w2_full {rel(t[1, 5+6*j])} h_full {rel(t[1, 6+6*j])} g2_issued {rel(t[1, 7+6*j])} | "
                  f"wk acc1_full {rel(t[2, 1+5*j])} ld_done {rel(t[2, 2+5*j])} gelu_done {rel(t[2, 3+5*j])} h_empty {rel(t[2, 4+5*j])} "
                  f"stores_done {rel(t[2, 210+j])} fence_done {rel(t[2, 220+j])} h_written {rel(t[2, 5+5*j])}")
