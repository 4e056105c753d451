#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest_gpu.log; tail -4 gpurun_out/s2_pytest_gpu.log
bash tools/gpu_r02_s2_bq.sh
