#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 --no-variants --no-cpu-baseline 2>gpurun_out/s2_bench_quick.err > gpurun_out/s2_bench_quick.json
tail -3 gpurun_out/s2_bench_quick.err
python -c "
import sys, json
d = json.loads([l for l in open('gpurun_out/s2_bench_quick.json') if l.startswith('{')][-1])
print('ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['ms_per_step'], 4), 'synth', {k: (round(v['ms'],3), round(v['device_ms'],3)) for k, v in d['synthesis'].items()})
r = d['roofline']; print(r['kernel'], r['achieved'], r['frac'], r['avg_us'], r['share_of_lib_time']); print(r['all_tensor_kernels'])
for t in d['top_kernels']: print(t)
"
