#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python tools/probe_t3.py > gpurun_out/probe_t3.txt 2>&1; grep -v Warn gpurun_out/probe_t3.txt | head -80
timeout 600 python bench.py --steps 20 --warmup 5 --no-variants --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print("ms/step", d['ms_per_step'], "e2e", d['e2e']['ms_per_step'], "launches", d['gpu_launches_per_step'])
PY
python tools/timeline_step.py > gpurun_out/timeline_graph.txt 2>&1; head -3 gpurun_out/timeline_graph.txt
