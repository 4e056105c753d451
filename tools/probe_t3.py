"""GAN-phase training step (T3): CUDA-graph replay time with the native period discriminators vs cuDNN, and a per-kernel profile."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from optispeech_b200.factory import DEFAULT_MODEL, build_model  # noqa: E402
from optispeech_b200.model.vocoder.wavenext.disc import _discriminators as D  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
batch = bench.make_batch(bench.B_PER_GPU, seed=1)
which = sys.argv[1:] or ["native", "cudnn"]
for mode in which:
    D.NATIVE_MPD = mode == "native"
    torch.manual_seed(0)
    model = build_model(DEFAULT_MODEL, train_args=dict(pretraining_steps=0)).to(dev).train()
    dev_batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
    model.cuda_graph = True
    t3 = bench._event_time(lambda i: model.training_step(dev_batch, i), n=10, warm=7)
    print(f"[{mode}] T3 graph replay: {t3:.3f} ms/step")
    if model._graphed is not None:
        model._graphed.release()
    if int(os.environ.get('T3_ROWS', '28')) == 0:
        continue
    model.cuda_graph = False
    model._graphed = None
    for i in range(2):
        model.training_step(dev_batch, i)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model.training_step(dev_batch, 3)
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    total = sum(e.device_time_total for e in rows)
    print(f"[{mode}] eager step: {total / 1e3:.2f} ms of kernel time, {sum(e.count for e in rows)} launches")
    for e in rows[:int(os.environ.get('T3_ROWS', '28'))]:
        print(f"   {e.device_time_total / 1e3:8.3f} ms  x{e.count:<4d} {e.key[:110]}")
    del model
    torch.cuda.empty_cache()
