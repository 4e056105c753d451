#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_full.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
tail -2 gpurun_out/bench_ref.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_full.json'))
print("ms/step", d['ms_per_step'], "e2e", d['e2e'], "launches", d['gpu_launches_per_step'])
print("clocks", d['clocks'])
print("roofline", {k:v for k,v in d['roofline'].items() if k not in ('per_shape','all_tensor_kernels')})
print("cpu", d['cpu_baseline']); print("gpu lib", d['gpu_library_baseline']); print("vs_torch_gpu", d['vs_torch_gpu'])
for k,v in d['variants'].items(): print(k, json.dumps(v)[:600])
print("synth", json.dumps(d['synthesis'])[:1500])
r=json.load(open('gpurun_out/bench_ref.json')); print("ref line", json.dumps(r)[:700])
PY
