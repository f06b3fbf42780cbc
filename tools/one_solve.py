#!/usr/bin/env python
"""Tiny driver for ncu: a few single solves (n=14, N=128) then one batched solve (256 systems)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpcgpu_b200 as m  # noqa: E402
from mpcgpu_b200 import synth  # noqa: E402

n, N = int(os.environ.get("STATE", "14")), int(os.environ.get("KNOTS", "128"))
B = int(os.environ.get("BATCH", "256"))
d = synth.make_systems(n, N, batch=max(B, 8), seed=3)
S, P, g = (torch.from_numpy(d[k]).cuda() for k in ("S", "Pinv", "gamma"))
if os.environ.get("TUNE_MODE"):                            # pin a kernel variant: TUNE_C (cluster size), TUNE_MODE (gbd_variants.h)
    from mpcgpu_b200 import _capi
    assert _capi.lib().gbd_pcg_set_tuning(n, N, 0, int(os.environ.get("TUNE_C", "0")), int(os.environ["TUNE_MODE"])) == 0
it = torch.zeros(B, dtype=torch.int32, device="cuda")
fl = torch.zeros(B, dtype=torch.uint8, device="cuda")
for i in range(int(os.environ.get("SINGLES", "24"))):
    lam = torch.zeros(n * N, device="cuda")
    m.pcg_launch(n, N, S[i % 8], P[i % 8], g[i % 8], lam, None, None, None, None, it, fl, int(os.environ.get("CAP", "167")),
                 float(os.environ.get("TOL", "1e-4")))
torch.cuda.synchronize()
if B > 1:
    for _ in range(2):
        lam = torch.zeros(B, n * N, device="cuda")
        m.solve_batched(n, N, B, S[:B], P[:B], g[:B], lam, it, fl, 167, 1e-4)
    torch.cuda.synchronize()
print("done", int(it[0].item()))
