#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/probe_fs.py 2>&1 | grep -v Warn | tee gpurun_out/s2_probe_fs.txt
bash tools/gpu_r02_s2_full.sh
