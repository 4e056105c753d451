#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_two_rank_nccl_gpu.py -m gpu -q -s > gpurun_out/s2_nccl_2gpu_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_nccl_2gpu_test.log
grep -n "graph=\|passed\|failed\|Error\|diverged" gpurun_out/s2_nccl_2gpu_test.log | head -20
