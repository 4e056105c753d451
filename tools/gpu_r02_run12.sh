#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_autograd_fn_gpu.py tests/test_synthesis_gpu.py tests/test_mpd_native_gpu.py -m gpu -q -x -k "convnext or synthes or period" > gpurun_out/pytest_sub.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_sub.log
python tools/probe_r02.py block timeline 2>&1 | grep -v Warn | tail -24
timeout 600 python tools/probe_t3.py native 2>&1 | grep -v Warn | head -14
