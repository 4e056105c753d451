"""Run-to-run determinism of the gradient bucket (eager steps, same batch and weights, lr = 0), with and without side streams."""
import os
import sys
from functools import partial

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_public_surface_gpu import ModelSpec, _fresh_model, _small_batch  # noqa: E402

from optispeech_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
spec = ModelSpec()
A = _small_batch(spec, 3, 48, 200, seed=7, dev=dev)


def bucket(side: bool, steps: int = 4):
    ops.SIDE_STREAMS_ENABLED = side
    model = _fresh_model(spec, dev)
    model.hparams.optimizer = partial(torch.optim.AdamW, lr=0.0, betas=[0.8, 0.99], weight_decay=0.0)
    model.cuda_graph = False
    out = []
    for i in range(steps):
        model.training_step(A, i)
        torch.cuda.synchronize()
        b = model.optimizers()[0].buckets()[0]
        out.append(b.flat_g.clone())
    names = {id(p): n for n, p in model.generator.named_parameters()}
    return out, [(names[id(p)], tuple(p.shape)) for p in b.params], list(b.offsets)


def report(tag, x, y, shapes, offs):
    rows = []
    for (name, shp), o in zip(shapes, offs):
        n = int(np.prod(shp))
        a, b = x[o:o + n], y[o:o + n]
        rows.append((float((a - b).norm() / (a.norm() + 1e-20)), name, shp, float(a.norm())))
    rows.sort(reverse=True)
    print(tag)
    for r in rows[:6]:
        print(f"   {r[0]:.3e}  {r[1]} {r[2]}  |g| = {r[3]:.3e}")


for side in (False, True):
    g1, shapes, offs = bucket(side)
    g2, _, _ = bucket(side)
    report(f"side={side}: step 3 vs step 2 of the same model", g1[3], g1[2], shapes, offs)
    report(f"side={side}: model A step 3 vs model B step 3", g1[3], g2[3], shapes, offs)
