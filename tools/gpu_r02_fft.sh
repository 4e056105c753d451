#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_spectral_gpu.py tests/test_features_gpu.py tests/test_public_surface_gpu.py -m gpu -q -k "spectral or mel or stft or features or energy or vocos" > gpurun_out/fft_test.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/fft_test.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_var.json'))
print("ms/step", d['ms_per_step'], "e2e", d['e2e']['ms_per_step'])
for k,v in d['variants'].items(): print(k, json.dumps(v)[:420])
print("synth", {k: (v['ms'], v['device_ms']) for k, v in d['synthesis'].items()})
PY
