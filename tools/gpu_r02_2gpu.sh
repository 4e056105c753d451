#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_two_rank_nccl_gpu.py -m gpu -q -s > gpurun_out/nccl_test.log 2>&1; echo "nccl test rc=$?"
grep -n "graph=\|passed\|failed\|skipped\|Error" gpurun_out/nccl_test.log | cut -c1-400 | head
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-variants > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"
head -c 300 gpurun_out/bench_2gpu.json; echo
