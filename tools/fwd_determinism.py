"""Where does run-to-run variation enter?  Same model, same batch: forward + backward three times, compare every aux output and gradient."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_public_surface_gpu import ModelSpec, _fresh_model, _small_batch  # noqa: E402

from optispeech_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
spec = ModelSpec()
A = _small_batch(spec, 3, 48, 200, seed=7, dev=dev)
ops.SIDE_STREAMS_ENABLED = bool(int(os.environ.get("SIDE", "0")))
model = _fresh_model(spec, dev)
model.generator.vocoder_needs_grad = False
runs = []
for it in range(3):
    for p in model.generator.parameters():
        p.grad = None
    out = model._process_batch(A)
    (out["loss"] * 1024.0).backward()
    ops.join_grad_streams()
    torch.cuda.synchronize()
    rec = {k: v.detach().float().clone() for k, v in out["_aux"].items() if isinstance(v, torch.Tensor)}
    for k in ("loss", "align_loss", "duration_loss", "pitch_loss", "energy_loss"):
        rec[k] = out[k].detach().float().clone()
    for n, p in model.generator.named_parameters():
        if p.grad is not None:
            rec["grad/" + n] = p.grad.detach().float().clone()
    runs.append(rec)
for a, b, tag in ((runs[0], runs[1], "run0 vs run1"), (runs[1], runs[2], "run1 vs run2")):
    rows = []
    for k in a:
        x, y = a[k], b[k]
        fin = torch.isfinite(x) & torch.isfinite(y)
        d = float((x[fin] - y[fin]).norm() / (x[fin].norm() + 1e-20))
        rows.append((d, k))
    rows.sort(reverse=True)
    print(tag)
    for d, k in rows[:14]:
        print(f"   {d:.3e}  {k}")
    print("   non-grad entries:", [(k, f"{d:.2e}") for d, k in rows if not k.startswith("grad/")])
