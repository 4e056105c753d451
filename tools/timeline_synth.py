"""Developer tool: per-kernel timeline (start, duration, gap to the previous activity's end) of ONE graph-mode synthesis call."""
import json
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from optispeech_b200.factory import DEFAULT_MODEL, build_model
from torch.profiler import ProfilerActivity, profile

dev = torch.device("cuda:0")
torch.manual_seed(1234)
model = build_model(DEFAULT_MODEL, train_args=dict(pretraining_steps=10 ** 9)).to(dev).eval()
for name in (sys.argv[1:] or ["single_B1_Tx120"]):
    ids, lens, durs = bench.synth_inputs(name)
    ids_pin = ids.pin_memory()
    for _ in range(5):
        out = model.generator.synthesise(ids_pin.to(dev, non_blocking=True), lens, durations=durs)
    torch.cuda.synchronize()
    print(f"--- {name}: latency {out['latency']:.3f} ms")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        out = model.generator.synthesise(ids_pin.to(dev, non_blocking=True), lens, durations=durs)
        torch.cuda.synchronize()
    path = os.path.join(tempfile.mkdtemp(), "trace.json")
    prof.export_chrome_trace(path)
    tr = json.load(open(path))
    ks = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
    ks.sort(key=lambda e: e["ts"])
    t0 = ks[0]["ts"]
    prev_end = t0
    busy = 0.0
    for e in ks:
        nm = e["name"]
        for pre in ("void osb::(anonymous namespace)::", "osb::(anonymous namespace)::", "void at::native::", "at::native::"):
            nm = nm.replace(pre, "")
        print(f"{e['ts'] - t0:9.1f} {e['dur']:7.1f} gap {e['ts'] - prev_end:7.1f}  {nm[:80]}")
        prev_end = max(prev_end, e["ts"] + e["dur"])
        busy += e["dur"]
    print(f"span {prev_end - t0:.1f} us, busy {busy:.1f} us, {len(ks)} activities")
