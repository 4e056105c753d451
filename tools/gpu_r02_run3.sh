#!/bin/bash
# round 2, run 3: log2 warp forward-sum, fused ConvNeXt train forward / backward kernels
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_autograd_fn_gpu.py -m gpu -q -k "forward_sum or convnext_block" -s > gpurun_out/pytest_k.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_k.log
grep -n "rel err\|warp \|PASS\|FAIL\|passed\|failed\|Error\|rc=" gpurun_out/pytest_k.log | head -80
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
grep -n "per family\|MAS durations\|eager vs eager\|step losses\|wav_hat max\|durations differing\|utterance\|forward_gen\|d loss_gen" gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-variants --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches_per_step'])
for t in d['top_kernels']: print(t)
print(d['roofline']['all_tensor_kernels'])
PY
