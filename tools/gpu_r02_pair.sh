#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('ms/step', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['ms_per_step'], 4))
for k in ('T2_generator_step_with_spectral_losses','T3_full_gan_training_step','transformer_backbone'): print(k, d['variants'][k]['ms_per_step'])
print([ (s['shape'], s['avg_us']) for s in d['roofline']['per_shape'][:3]])
print({k: v for k, v in d['roofline']['all_tensor_kernels'].items()})"
