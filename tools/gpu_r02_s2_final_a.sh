#!/bin/bash
# final evidence of round 2, part A: GPU test log, smoke, bench lines (ours with variants + baselines, reference arm)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_final.log; tail -3 gpurun_out/r02_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -2
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_full.json') if l.startswith('{')][-1])
print("ms/step", d['ms_per_step'], "value", d['value'], "e2e", d['e2e'])
print("roofline", {k:v for k,v in d['roofline'].items() if k not in ('per_shape','all_tensor_kernels','timing')})
print("cpu", d['cpu_baseline'].get('value'), "gpu lib", d['gpu_library_baseline'].get('value'), "vs_torch_gpu", d.get('vs_torch_gpu'))
for k,v in d['variants'].items(): print(k, v.get('ms_per_step'), v.get('error'))
print("synth", {k: (v['ms'], v['device_ms']) for k, v in d['synthesis'].items()})
PY
