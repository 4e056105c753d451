#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
grep -n "waveform max-abs\|max-abs diff" gpurun_out/pytest_gpu.log | head
timeout 300 python tools/probe_synth.py single_B1_Tx120 long_B8_Tx512 2>&1 | grep -v "Warn\|WeightNorm" | grep "^---\|gemm_nt" | head -24
