"""Developer tool: torch.profiler breakdown of one training step (which kernels / host ops dominate)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from optispeech_b200.factory import DEFAULT_MODEL, build_model

dev = torch.device("cuda:0")
torch.manual_seed(1234)
model = build_model(DEFAULT_MODEL, train_args=dict(pretraining_steps=10 ** 9)).to(dev).train()
hb = bench.make_batch(32, 1234)
db = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in hb.items()}
for i in range(3):
    model.training_step(db, i)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(2):
        model.training_step(db, i)
    torch.cuda.synchronize()
evs = [e for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in evs)
n = sum(e.count for e in evs)
print(f"device kernels: {n // 2} per step, {tot / 2e3:.3f} ms per step")
ours = sum(e.self_device_time_total for e in evs if "osb::" in e.key or "osb_" in e.key or "_kernel<" in e.key and "at::" not in e.key)
print(f"share of library kernels (name match, approximate): {ours / tot:.3f}")
for e in evs[:70]:
    print(f"{e.self_device_time_total / 2e3:9.4f} ms  {e.count // 2:4d}x  {e.self_device_time_total / e.count:8.2f} us  {e.key[:110]}")
