#!/bin/bash
# end-of-round evidence run: parity suite, bench (both arms), ncu launch list + full capture, Transformer-step probe
set -x
mkdir -p gpurun_out
timeout 300 python tools/probe_transformer.py > gpurun_out/probe_transformer.log 2>&1; tail -45 gpurun_out/probe_transformer.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --eager --quick --steps 1 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:convnext_fused|gemm_nt_kernel|gemm_wgrad' \
   --launch-skip 243 -c 24 -o gpurun_out/r01_full -f python bench.py --eager --quick --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
du -sh gpurun_out
