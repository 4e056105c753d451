"""Developer tool: per-kernel timeline (start, duration, stream) of ONE CUDA-graph replay of the training step."""
import json
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from optispeech_b200.factory import DEFAULT_MODEL, build_model

dev = torch.device("cuda:0")
torch.manual_seed(1234)
model = build_model(DEFAULT_MODEL, train_args=dict(pretraining_steps=10 ** 9)).to(dev).train()
model.cuda_graph = "--eager" not in sys.argv
hb = bench.make_batch(32, 1234)
db = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in hb.items()}
for i in range(8):
    model.training_step(db, i)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model.training_step(db, 0)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
tr = json.load(open(path))
ks = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
ks.sort(key=lambda e: e["ts"])
t0 = ks[0]["ts"]
end = max(e["ts"] + e["dur"] for e in ks)
streams = {}
for e in ks:
    streams.setdefault(e["args"].get("stream"), len(streams))
print(f"kernels {len(ks)}  span {end - t0:.1f} us  streams {len(streams)}")
busy = {}
for e in ks:
    s = streams[e["args"].get("stream")]
    busy[s] = busy.get(s, 0.0) + e["dur"]
print("busy us per stream:", {k: round(v, 1) for k, v in sorted(busy.items())})
for e in ks:
    name = e["name"]
    for pre in ("void osb::(anonymous namespace)::", "osb::(anonymous namespace)::", "void at::native::", "at::native::"):
        name = name.replace(pre, "")
    print(f"{e['ts'] - t0:9.1f} {e['dur']:7.1f}  s{streams[e['args'].get('stream')]}  {name[:90]}")
