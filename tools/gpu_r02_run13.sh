#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_disc_native_gpu.py tests/test_features_gpu.py -m gpu -q -s > gpurun_out/mrd_test.log 2>&1; echo "rc=$?"
grep -n "resolution \|n_fft\|passed\|failed\|^E  " gpurun_out/mrd_test.log | cut -c1-400 | head -40
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python tools/probe_t3.py native > gpurun_out/probe_t3.txt 2>&1; grep -v Warn gpurun_out/probe_t3.txt | head -40
