#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/probe_gemm_trace.py > gpurun_out/probe_gemm_trace.log 2>&1; cat gpurun_out/probe_gemm_trace.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/bench.json
