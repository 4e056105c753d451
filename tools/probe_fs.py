"""Developer probe: launch shapes of the forward-sum warp recursion (osb_debug_forward_sum_variant) at the benchmark shape —
device time per call inside a CUDA graph and equality of the results with the default shape."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from optispeech_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()
lib.osb_debug_forward_sum_variant.argtypes = [C.c_int]
hb = bench.make_batch(32, 1234)
xl, ml = hb["x_lengths"].to(dev), hb["mel_lengths"].to(dev)
g = torch.Generator().manual_seed(0)
lp = torch.log_softmax(torch.randn(32, 864, 192, generator=g) * 2.0, dim=-1).to(dev)
ref = None
for v in (0, 1, 0, 1):
    lib.osb_debug_forward_sum_variant(v)
    for _ in range(2):
        loss, grad = ops.forward_sum(lp, xl, ml, -1.0)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(5):
            loss, grad = ops.forward_sum(lp, xl, ml, -1.0)
    graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    if ref is None:
        ref = (loss.clone(), grad.clone())
    dl, dg = (loss - ref[0]).abs().max().item(), (grad - ref[1]).abs().max().item()
    print(f"variant {v}: {us:7.1f} us per call (lse + recursion + gradient kernel); max diff vs variant 0: loss {dl:.2e}, grad {dg:.2e}; "
          f"loss[0] {loss[0].item():.6f}")
lib.osb_debug_forward_sum_variant(1)
