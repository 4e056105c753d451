#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_training_gpu.py -m gpu -q -s -k windowed > gpurun_out/s2_pytest_one.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest_one.log
grep -n "sample \|wav_hat:\|passed\|failed\|^E " gpurun_out/s2_pytest_one.log | head -20
