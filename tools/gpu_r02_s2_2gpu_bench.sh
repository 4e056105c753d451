#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-variants > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_2gpu.json') if l.startswith('{')][-1])
print({k: d.get(k) for k in ('n_gpus','ms_per_step','value','allreduce')}, d['e2e'])
PY
