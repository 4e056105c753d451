#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_training_step_gpu.py -m gpu -q -x -s -k "overlapped" > gpurun_out/s2_pytest_sched.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest_sched.log
grep -n "serial vs\|passed\|failed\|Error\|assert" gpurun_out/s2_pytest_sched.log | head -20
