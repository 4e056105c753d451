#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/timeline_synth.py single_B1_Tx120 2>&1 | grep -v "Warn\|WeightNorm" > gpurun_out/s2_timeline_synth.txt
tail -3 gpurun_out/s2_timeline_synth.txt
