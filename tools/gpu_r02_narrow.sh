#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for mode in 1 0 1; do
OSB_NARROW=$mode timeout 600 python - <<'PY'
import os, sys, json, subprocess, ctypes
sys.path.insert(0, '.')
from optispeech_b200 import _lib
lib = _lib.load()
lib.osb_debug_set_gemm_narrow_tiles.argtypes = [ctypes.c_int]
lib.osb_debug_set_gemm_narrow_tiles(int(os.environ["OSB_NARROW"]))
import bench
sys.argv = ["bench.py", "--steps", "30", "--warmup", "5", "--no-variants", "--no-cpu-baseline"]
import io, contextlib
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    bench.main()
d = json.loads([l for l in buf.getvalue().splitlines() if l.startswith("{")][-1])
print("narrow", os.environ["OSB_NARROW"], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), "gemm family tflops", round(d["roofline"]["achieved"], 1), "avg_us", round(d["roofline"]["avg_us"], 2))
PY
done
