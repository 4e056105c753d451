#!/bin/bash
# final evidence of round 2: GPU test log, bench lines (ours + reference arm), ncu launch list + --set full summary (exported on the box)
mkdir -p gpurun_out /tmp/ncu
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_final.log; tail -3 gpurun_out/r02_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -2
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv \
   python bench.py --eager --quick --steps 1 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
timeout 1200 ncu --set full --clock-control none -k 'regex:convnext_fused|convnext_bwd_fused|gemm_nt_kernel|gemm_wgrad' \
   --launch-skip 219 -c 73 -o /tmp/ncu/r02_full -f python bench.py --eager --quick --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/ncu/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2> gpurun_out/ncu_export.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_full.json'))
print("ms/step", d['ms_per_step'], "value", d['value'], "e2e", d['e2e'])
print("roofline", {k:v for k,v in d['roofline'].items() if k not in ('per_shape','all_tensor_kernels')})
print("cpu", d['cpu_baseline']['value'], "gpu lib", d['gpu_library_baseline']['value'], "vs_torch_gpu", d['vs_torch_gpu'])
for k,v in d['variants'].items(): print(k, v.get('ms_per_step'))
print("synth", {k: (v['ms'], v['device_ms']) for k, v in d['synthesis'].items()})
PY
du -sh gpurun_out
