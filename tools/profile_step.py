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
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=70))
