#!/bin/bash
# A/B of two launch-shape hooks inside the final multi-stream step (the windowed decoder freed SM-time): 128-wide GEMM tiles for
# small problems, and the minimum pipeline iterations per weight-gradient CTA
mkdir -p gpurun_out
for cfg in "0 48" "1 48" "0 24" "0 96" "1 24" "0 48"; do
set -- $cfg
OSB_NARROW=$1 OSB_WG=$2 timeout 300 python - <<'PY' 2>/dev/null
import os, sys, json, ctypes, io, contextlib
sys.path.insert(0, '.')
from optispeech_b200 import _lib
lib = _lib.load()
lib.osb_debug_set_gemm_narrow_tiles.argtypes = [ctypes.c_int]
lib.osb_debug_set_wgrad_min_iters.argtypes = [ctypes.c_int]
lib.osb_debug_set_gemm_narrow_tiles(int(os.environ["OSB_NARROW"]))
lib.osb_debug_set_wgrad_min_iters(int(os.environ["OSB_WG"]))
import bench
sys.argv = ["bench.py", "--steps", "40", "--warmup", "5", "--quick"]
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    bench.main()
d = json.loads([l for l in buf.getvalue().splitlines() if l.startswith("{")][-1])
print("narrow", os.environ["OSB_NARROW"], "wgrad min iters", os.environ["OSB_WG"], "ms/step", round(d["ms_per_step"], 4))
PY
done | tee gpurun_out/s2_tune.txt
