#!/usr/bin/env python
"""A/B of the batched v5 launch: clusters drawing systems from a counter vs the fixed stride (GBD_PCG_STATIC_BATCH=1).
Prints ms per batched solve and a digest of (lambda, iters, flags) -- the two modes must print the same digest."""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpcgpu_b200 as m  # noqa: E402
from mpcgpu_b200 import synth  # noqa: E402

n, N, B = 14, int(os.environ.get("KNOTS", "128")), int(os.environ.get("BATCH", "1024"))
d = synth.make_systems(n, N, batch=B, seed=5000)
if os.environ.get("SPREAD"):      # right-hand sides over six decades: iteration counts from 1 to the cap
    d["gamma"] = (d["gamma"] * (10.0 ** np.random.default_rng(5).uniform(-4.0, 2.0, size=B)).astype(np.float32)[:, None]).astype(np.float32)
S, P, g = (torch.from_numpy(d[k]).cuda() for k in ("S", "Pinv", "gamma"))
it = torch.zeros(B, dtype=torch.int32, device="cuda")
fl = torch.zeros(B, dtype=torch.uint8, device="cuda")
reps, warm = 6, 3
lam = torch.zeros(reps + warm, B, n * N, device="cuda")
for i in range(warm):
    m.solve_batched(n, N, B, S, P, g, lam[i], it, fl, 167, 1e-4)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(warm, warm + reps):
    m.solve_batched(n, N, B, S, P, g, lam[i], it, fl, 167, 1e-4)
e1.record()
torch.cuda.synchronize()
h = hashlib.sha256()
h.update(lam[-1].cpu().numpy().tobytes()); h.update(it.cpu().numpy().tobytes()); h.update(fl.cpu().numpy().tobytes())
itn = it.cpu().numpy()
print({"mode": "static" if os.environ.get("GBD_PCG_STATIC_BATCH") else "draw", "N": N, "batch": B, "spread": bool(os.environ.get("SPREAD")),
       "ms": e0.elapsed_time(e1) / reps, "traj_per_sec": B / (e0.elapsed_time(e1) / reps * 1e-3),
       "iters_mean": float(itn.mean()), "iters_min": int(itn.min()), "iters_max": int(itn.max()),
       "digest": h.hexdigest()[:16]})
