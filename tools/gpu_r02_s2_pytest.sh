#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_final.log; tail -3 gpurun_out/r02_pytest_gpu_final.log
