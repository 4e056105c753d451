#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probe_gemm_graph.py 2>&1 | grep -v Warn > gpurun_out/s2_gemm_graph.txt; cat gpurun_out/s2_gemm_graph.txt
