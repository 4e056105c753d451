"""Developer probe: per-kernel device times of one Transformer-configuration training step (launches queued behind a spin,
CUDA events around every library call) and of the attention kernels alone at the decoder / encoder shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from optispeech_b200 import _lib, ops  # noqa: E402
from optispeech_b200.factory import build_model, transformer_model_config  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = build_model(transformer_model_config(), train_args=dict(pretraining_steps=10 ** 9)).to(dev).train()
batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in bench.make_batch(32, 1234).items()}
for i in range(3):
    model.training_step(batch, i)
torch.cuda.synchronize()
torch.cuda._sleep(60_000_000)
with _lib.LaunchProfiler() as prof:
    model.training_step(batch, 3)
summ = prof.summary()
total = sum(a["total_ms"] for a in summ)
print(f"library kernels: {sum(a['launches'] for a in summ)} launches, {total:.3f} ms")
for a in summ[:28]:
    print(f"  {a['total_ms']:.4f} ms {a['launches']:3d}x {a['avg_us']:8.2f} us  {a['key']}")

for (B, T) in [(32, 864), (32, 192)]:
    qkv = torch.randn(B, T, 768, device=dev).half()
    lens = torch.full((B,), T, device=dev, dtype=torch.int64)
    for _ in range(3):
        ctx, rmax, rinv = ops.mha_fwd(qkv, 2, lens, save_stats=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(20_000_000)
    e0.record()
    for _ in range(20):
        ops.mha_fwd(qkv, 2, lens, save_stats=True)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 4.0 * B * 2 * T * T * 128 * 1.5     # QK^T twice + PV
    print(f"mha_fwd B={B} T={T}: {us:.1f} us, {fl / us / 1e6:.1f} TFLOP/s executed ({fl / 1.5 / us / 1e6:.1f} algorithmic)")
    d = torch.randn(B, T, 256, device=dev).half()
    for _ in range(2):
        ops.mha_bwd(qkv, 2, lens, ctx, d, rmax, rinv)
    torch.cuda._sleep(20_000_000)
    e0.record()
    for _ in range(10):
        ops.mha_bwd(qkv, 2, lens, ctx, d, rmax, rinv)
    e1.record()
    torch.cuda.synchronize()
    print(f"mha_bwd (+dK/dV contractions, pack) B={B} T={T}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")
