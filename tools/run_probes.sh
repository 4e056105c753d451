#!/bin/bash
# Runs every GEMM probe case in its own process under a timeout; logs to gpurun_out/probe_gemm.log
mkdir -p gpurun_out
LOG=gpurun_out/probe_gemm.log
: > $LOG
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv >> $LOG 2>&1
python - >> $LOG 2>&1 <<'PY'
import os, torch
p = torch.cuda.get_device_properties(0)
print("SMs", p.multi_processor_count, "L2", p.L2_cache_size, "cc", p.major, p.minor, "cpus", os.cpu_count())
print("ref exists:", os.path.exists("/root/reference"))
PY
lscpu | grep -E "Model name|^CPU\(s\)" >> $LOG 2>&1
for c in "$@"; do
  timeout 120 python tools/probe_gemm.py $c >> $LOG 2>&1
  echo "[$c] exit=$?" >> $LOG
done
tail -n 80 $LOG
