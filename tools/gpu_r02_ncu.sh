#!/bin/bash
# ncu evidence for round 2: launch list of one (eager) training step + one --set full capture of the tensor kernels
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv \
   python bench.py --eager --quick --steps 1 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:convnext_fused|convnext_bwd_fused|gemm_nt_kernel|gemm_wgrad' \
   --launch-skip 300 -c 40 -o gpurun_out/r02_full -f python bench.py --eager --quick --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out | head -30; du -sh gpurun_out
