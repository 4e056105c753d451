"""Developer tool: time osb_mas / osb_forward_sum alone on the bench shape."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optispeech_b200 import ops

dev = torch.device("cuda:0")
B, Tm, Tx = 32, 864, 192
g = torch.Generator().manual_seed(0)
lp = torch.log_softmax(torch.randn(B, Tm, Tx, generator=g), dim=-1).to(dev)
xl = torch.randint(Tx // 2, Tx + 1, (B,), generator=g).to(dev); ml = (xl * 4.5).long().clamp(max=Tm)
xl[0], ml[0] = Tx, Tm


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


print("osb_mas          %.1f us" % timeit(lambda: ops.mas(lp, xl, ml)))
print("osb_forward_sum  %.1f us" % timeit(lambda: ops.forward_sum(lp, xl, ml, -1.0)))
for tx in (40, 96, 512):
    lp2 = torch.log_softmax(torch.randn(8, 600, tx, generator=g), dim=-1).to(dev)
    x2 = torch.full((8,), tx, dtype=torch.int64, device=dev); m2 = torch.full((8,), 600, dtype=torch.int64, device=dev)
    print("osb_mas Tx=%d Tm=600  %.1f us" % (tx, timeit(lambda: ops.mas(lp2, x2, m2))))
