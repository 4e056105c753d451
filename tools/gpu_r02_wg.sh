#!/bin/bash
for it in 48 96 192 1000; do
OSB_WG=$it timeout 300 python - <<'PY' 2>/dev/null
import os, sys, json, ctypes, io, contextlib
sys.path.insert(0, '.')
from optispeech_b200 import _lib
lib = _lib.load()
lib.osb_debug_set_wgrad_min_iters.argtypes = [ctypes.c_int]
lib.osb_debug_set_wgrad_min_iters(int(os.environ["OSB_WG"]))
import bench
sys.argv = ["bench.py", "--steps", "30", "--warmup", "5", "--no-variants", "--no-cpu-baseline"]
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    bench.main()
d = json.loads([l for l in buf.getvalue().splitlines() if l.startswith("{")][-1])
print("wgrad min iters", os.environ["OSB_WG"], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), "wgrad", d["roofline"]["all_tensor_kernels"]["osb_gemm_wgrad"])
PY
done
