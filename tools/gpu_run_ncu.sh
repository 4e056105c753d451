#!/bin/bash
# ncu evidence for the current build: launch list of one training step + one --set full capture of the tensor kernels
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --eager --quick --steps 1 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:convnext_fused|gemm_nt_kernel|gemm_wgrad' \
   --launch-skip 250 -c 24 -o gpurun_out/r01_full -f python bench.py --eager --quick --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out; du -sh gpurun_out
