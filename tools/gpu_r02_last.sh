#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_final.log; tail -4 gpurun_out/r02_pytest_gpu_final.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-variants --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['ms_per_step'], 4), 'synth', {k: (round(v['ms'],3), round(v['device_ms'],3)) for k, v in d['synthesis'].items()})"
