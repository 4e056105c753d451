"""Per-kernel profile of one B=1 x 120-phoneme synthesis (eager), sorted by total device time."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from optispeech_b200.factory import DEFAULT_MODEL, build_model
from torch.profiler import ProfilerActivity, profile
dev = torch.device('cuda:0')
torch.manual_seed(1234)
model = build_model(DEFAULT_MODEL, train_args=dict(pretraining_steps=10**9)).to(dev).eval()
model.generator.synthesis_graphs = False
for name in (sys.argv[1:] or ["single_B1_Tx120"]):
    ids, lens, durs = bench.synth_inputs(name)
    for _ in range(3):
        model.generator.synthesise(ids.to(dev), lens, durations=durs)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model.generator.synthesise(ids.to(dev), lens, durations=durs)
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_time_total > 0]
    ev.sort(key=lambda e: e.time_range.start)
    t0 = ev[0].time_range.start
    print(f"--- {name}: {len(ev)} device activities, span {(ev[-1].time_range.end - t0):.0f} us, busy {sum(e.device_time_total for e in ev):.0f} us")
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    for e in rows[:22]:
        print(f"   {e.device_time_total:8.1f} us  x{e.count:<4d} avg {e.device_time_total / e.count:6.1f}  {e.key[:100]}")
