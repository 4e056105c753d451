#!/bin/bash
# round 2, run 1: new parity tests + the whole GPU suite, bench (all arms), box facts
set -x
mkdir -p gpurun_out
nproc > gpurun_out/box.txt; lscpu | grep -i "model name" >> gpurun_out/box.txt; nvidia-smi -L >> gpurun_out/box.txt
ls oracle/_ref/optispeech | head -3 >> gpurun_out/box.txt
timeout 900 python -m pytest tests/test_fullsize_golden_gpu.py tests/test_public_surface_gpu.py -m gpu -q -x -s > gpurun_out/pytest_new.log 2>&1; echo "pytest-new rc=$?" >> gpurun_out/pytest_new.log
tail -25 gpurun_out/pytest_new.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json
tail -5 gpurun_out/bench.err
