#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:forward_sum_warp -c 1 -o gpurun_out/fs_warp -f python tools/probe_r02.py fs > gpurun_out/ncu_fs.log 2>&1
ncu -i gpurun_out/fs_warp.ncu-rep --page source --csv > gpurun_out/fs_warp_source.csv 2>/dev/null
ncu -i gpurun_out/fs_warp.ncu-rep --page raw --csv | python -c "
import sys,csv
r=list(csv.reader(sys.stdin))
h=r[0]
for row in r[2:]:
    for k,v in zip(h,row):
        if any(s in k for s in ('gpu__time_duration','sm__warps_active','smsp__issue_active','smsp__inst_executed.sum','launch__registers','smsp__cycles_active.avg','sm__inst_executed_pipe_xu','smsp__average_warp','stall')): print(k,v)
" > gpurun_out/fs_warp_raw.txt
tail -5 gpurun_out/ncu_fs.log; head -50 gpurun_out/fs_warp_raw.txt
