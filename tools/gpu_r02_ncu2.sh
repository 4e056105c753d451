#!/bin/bash
set -x
mkdir -p gpurun_out /tmp/ncu
timeout 1200 ncu --set full --clock-control none -k 'regex:convnext_fused|convnext_bwd_fused|gemm_nt_kernel|gemm_wgrad' \
   --launch-skip 219 -c 73 -o /tmp/ncu/r02_full -f python bench.py --eager --quick --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la /tmp/ncu
ncu -i /tmp/ncu/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2> gpurun_out/ncu_export.err
# one launch of the decoder-shape fused kernel with source-level counters, small enough to travel
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:convnext_fused_kernel' \
   --launch-skip 60 -c 2 -o gpurun_out/r02_fused_src -f python bench.py --eager --quick --steps 1 --warmup 3 > gpurun_out/ncu_src.log 2>&1
ls -la gpurun_out; du -sh gpurun_out
sz=$(stat -c %s /tmp/ncu/r02_full.ncu-rep); if [ "$sz" -lt 40000000 ]; then cp /tmp/ncu/r02_full.ncu-rep gpurun_out/; fi
du -sh gpurun_out
