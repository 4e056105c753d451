#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_disc_native_gpu.py tests/test_public_surface_gpu.py -m gpu -q -s -k "period or vocos" > gpurun_out/mpd_test.log 2>&1
grep -n "period \|forward_gen\|d loss\|passed\|failed\|^E  " gpurun_out/mpd_test.log | head -80
