#!/usr/bin/env python
"""Per-kernel durations of one batched SQP linear-system step (gbd_step_run_f32, 1024 trajectories); run under
ncu --metrics gpu__time_duration.sum (tools/gpu_validate.sh style).  Diagnostic tool."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpcgpu_b200 as mp  # noqa: E402
from mpcgpu_b200 import synth  # noqa: E402

n, m, N, B = 14, 7, 128, int(os.environ.get("BATCH", "1024"))
G, C, g, c = (torch.from_numpy(x).cuda().reshape(-1) for x in synth.make_kkt_batch(n, m, N, B, seed=1))
lam = torch.zeros(B * n * N, device="cuda")
dz = torch.zeros(B * ((n + m) * (N - 1) + n), device="cuda")
plan = mp.StepPlan(n, m, N, B)
for _ in range(2):
    Gw = G.clone()
    lam.zero_()
    plan.run(Gw, C, g, c, 1e-3, lam, dz, 167, 1e-4)
it, fl = plan.results()
print("mean iters", it.mean(), "flags", fl.mean())
