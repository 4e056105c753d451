#!/bin/bash
mkdir -p gpurun_out
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 --no-variants > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err; echo "bench$n rc=$?"
python - <<PY
import json
for ln in open('gpurun_out/bench_${n}gpu.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print("N=$n ms/step", round(d['ms_per_step'],4), "value", round(d['value']), "e2e", round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), "allreduce", d.get('allreduce'), "clk samples", d['clocks'].get('samples_in_timed_region'))
    else: print("non-json:", ln[:80])
PY
done
