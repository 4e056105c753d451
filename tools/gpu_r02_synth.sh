#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_synthesis_gpu.py tests/test_public_surface_gpu.py tests/test_fullsize_gpu.py tests/test_transformer_gpu.py -m gpu -q -x -k "synth or surface or long or utterance" > gpurun_out/synth_test.log 2>&1; echo "rc=$?"
tail -5 gpurun_out/synth_test.log
python - <<'PY'
import sys, time, torch
sys.path.insert(0, '.')
import bench
from optispeech_b200.factory import DEFAULT_MODEL, build_model
dev = torch.device('cuda:0')
torch.manual_seed(1234)
model = build_model(DEFAULT_MODEL, train_args=dict(pretraining_steps=10**9)).to(dev).eval()
for graphs in (False, True, False):
    model.generator.synthesis_graphs = graphs
    for name in ("long_B8_Tx512", "single_B1_Tx120"):
        ids, lens, durs = bench.synth_inputs(name)
        ids_pin = ids.pin_memory()
        for _ in range(4):
            out = model.generator.synthesise(ids_pin.to(dev, non_blocking=True), lens, durations=durs)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 10
        for _ in range(n):
            out = model.generator.synthesise(ids_pin.to(dev, non_blocking=True), lens, durations=durs)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / n * 1e3
        print(f"graphs={graphs} {name}: e2e {ms:.3f} ms, device latency {out['latency']:.3f} ms (am {out['am_rtf']:.2e} v {out['v_rtf']:.2e}), wav {tuple(out['wav'].shape)}")
PY
