#!/bin/bash
# final evidence of round 2, part B: ncu launch list of one eager step + --set full capture of the step's contraction kernels
mkdir -p gpurun_out /tmp/ncu
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv \
   python bench.py --eager --quick --steps 1 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
timeout 1200 ncu --set full --clock-control none -k 'regex:convnext_fused|convnext_bwd_fused|gemm_nt_kernel|gemm_wgrad' \
   --launch-skip 216 -c 72 -o /tmp/ncu/r02_full -f python bench.py --eager --quick --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i /tmp/ncu/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2> gpurun_out/ncu_export.err
wc -l gpurun_out/r02_launches.csv gpurun_out/r02_full_raw.csv
du -sh gpurun_out
