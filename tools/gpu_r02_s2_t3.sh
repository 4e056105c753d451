#!/bin/bash
mkdir -p gpurun_out
T3_ROWS=200 timeout 400 python tools/probe_t3.py native 2>&1 | grep -v "Warn\|WeightNorm\|_warn_once" > gpurun_out/s2_probe_t3.txt
head -5 gpurun_out/s2_probe_t3.txt
