#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/timeline_step.py 2>&1 | grep -v "Warn\|WeightNorm\|_warn_once" > gpurun_out/s2_timeline_graph.txt
head -3 gpurun_out/s2_timeline_graph.txt
