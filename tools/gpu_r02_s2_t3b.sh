#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_disc_native_gpu.py tests/test_public_surface_gpu.py tests/test_training_step_gpu.py -m gpu -q -x > gpurun_out/s2_pytest_disc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest_disc.log
tail -5 gpurun_out/s2_pytest_disc.log
T3_ROWS=0 timeout 300 python tools/probe_t3.py native 2>&1 | grep "T3 graph"
