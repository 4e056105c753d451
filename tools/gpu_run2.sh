#!/bin/bash
# One GPU session: new Transformer-path tests first, then the whole parity suite, bench (both arms), ncu launch list and
# one --set full capture of the top tensor kernels (kept small: gpurun_out/ must stay under 64 MiB).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
timeout 420 python -m pytest tests/test_transformer_gpu.py -q -s > gpurun_out/pytest_tf.log 2>&1; echo "pytest_tf rc=$?" >> gpurun_out/pytest_tf.log
tail -15 gpurun_out/pytest_tf.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_transformer_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --eager --quick --steps 1 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:convnext_fused|gemm_nt_kernel|gemm_wgrad' \
   --launch-skip 250 -c 22 -o gpurun_out/r01_full -f python bench.py --eager --quick --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out; du -sh gpurun_out
