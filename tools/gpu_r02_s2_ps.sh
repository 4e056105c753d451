#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probe_synth.py single_B1_Tx120 long_B8_Tx512 2>&1 | grep -v "Warn\|WeightNorm" > gpurun_out/s2_probe_synth2.txt
head -48 gpurun_out/s2_probe_synth2.txt | cut -c1-150
