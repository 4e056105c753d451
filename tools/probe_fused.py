"""Developer probe: time the fused ConvNeXt block kernel on the decoder / vocoder shapes."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optispeech_b200 import ops
from optispeech_b200.model.generator.modules import ConvNeXtBlock

dev = torch.device("cuda:0")
for (C, I, B, T) in [(256, 1024, 32, 864), (384, 1152, 32, 64), (384, 1152, 8, 1024)]:
    torch.manual_seed(0)
    blk = ConvNeXtBlock(C, I, 0.0, 0.25).to(dev).eval()
    x = torch.randn(B, T, C, device=dev)
    with torch.no_grad():
        for split in (False, True):
            for _ in range(3):
                blk.forward_cl(x, None, split=split)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 20
            for _ in range(n):
                blk.forward_cl(x, None, split=split)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / n * 1e3
            fl = 2.0 * B * T * (2 * C * I)
            print(f"C={C} I={I} rows={B*T} {'3-kernel fp16x3' if split else 'fused fp16   '}: {us:8.1f} us  {fl/us/1e6:7.1f} TFLOP/s (algorithmic)")

# ---- clock64 timeline of CTA 0 (developer hook) ----
import ctypes as C
from optispeech_b200 import _lib
lib = _lib.load()
lib.osb_debug_set_fused_trace.argtypes = [C.c_void_p]
for (Cc, I, B, T) in [(384, 1152, 32, 64), (256, 1024, 32, 864)]:
    torch.manual_seed(0)
    blk = ConvNeXtBlock(Cc, I, 0.0, 0.25).to(dev).eval()
    x = torch.randn(B, T, Cc, device=dev)
    tr = torch.zeros(3 * 256, dtype=torch.int64, device=dev)
    with torch.no_grad():
        blk.forward_cl(x, None, split=False)
        lib.osb_debug_set_fused_trace(C.c_void_p(tr.data_ptr()))
        blk.forward_cl(x, None, split=False)
        torch.cuda.synchronize()
        lib.osb_debug_set_fused_trace(None)
    t = tr.cpu().view(3, 256)
    base = int(t[1, 0])
    rel = lambda v: int(v) - base if int(v) else None
    print(f"--- timeline C={Cc} I={I} (cycles since MMA thread start)")
    print("mma: a_ready", rel(t[1, 1]))
    print("worker0: prologue done", rel(t[2, 0]), " acc2_full", rel(t[2, 200]), " end", rel(t[2, 201]))
    for j in range(0, min(6, I // 64)):
        print(f"chunk {j}: prod w1_empty {rel(t[0, 2*j])} w2_empty {rel(t[0, 2*j+1])} | mma w1_full {rel(t[1, 2+6*j])} acc1_empty {rel(t[1, 3+6*j])} "
              f"g1_issued {rel(t[1, 4+6*j])} w2_full {rel(t[1, 5+6*j])} h_full {rel(t[1, 6+6*j])} g2_issued {rel(t[1, 7+6*j])} | "
              f"wk acc1_full {rel(t[2, 1+5*j])} ld_done {rel(t[2, 2+5*j])} gelu_done {rel(t[2, 3+5*j])} h_empty {rel(t[2, 4+5*j])} "
              f"stores_done {rel(t[2, 210+j])} fence_done {rel(t[2, 220+j])} h_written {rel(t[2, 5+5*j])}")
