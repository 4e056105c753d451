#!/bin/bash
for it in 6 9; do
OSB_NS=$it timeout 300 python - <<'PY' 2>/dev/null
import os, sys, json, ctypes, io, contextlib
sys.path.insert(0, '.')
from optispeech_b200 import _lib
lib = _lib.load()
lib.osb_debug_set_fused_nsplit_cap.argtypes = [ctypes.c_int]
lib.osb_debug_set_fused_nsplit_cap(int(os.environ["OSB_NS"]))
import bench
sys.argv = ["bench.py", "--steps", "30", "--warmup", "5", "--no-variants", "--no-cpu-baseline"]
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    bench.main()
d = json.loads([l for l in buf.getvalue().splitlines() if l.startswith("{")][-1])
print("fused nsplit cap", os.environ["OSB_NS"], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4))
PY
done
