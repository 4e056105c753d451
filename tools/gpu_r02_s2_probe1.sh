#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probe_gemm_trace.py > gpurun_out/s2_gemm_trace.txt 2>&1
timeout 300 python tools/probe_synth.py single_B1_Tx120 long_B8_Tx512 2>&1 | grep -v "Warn\|WeightNorm" > gpurun_out/s2_probe_synth.txt
cat gpurun_out/s2_gemm_trace.txt | tail -12
head -30 gpurun_out/s2_probe_synth.txt
