#!/usr/bin/env python
"""Repeatability stress: the same solves many times, every result compared bit-for-bit with the first (the packet kernels
synchronise through polled shared memory; an ordering bug would show up as a rare mismatch or a hang).  Bounded run time."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpcgpu_b200 as mp  # noqa: E402
from mpcgpu_b200 import synth  # noqa: E402

n = 14
t_end = time.time() + float(os.environ.get("STRESS_SECONDS", "60"))
report = {}
for N, cap in ((32, 173), (64, 167), (128, 167)):
    d = synth.make_systems(n, N, batch=8, seed=N)
    S, P, g = (torch.from_numpy(d[k]).cuda() for k in ("S", "Pinv", "gamma"))
    it = torch.zeros(1, dtype=torch.int32, device="cuda")
    fl = torch.zeros(1, dtype=torch.uint8, device="cuda")
    want = []
    for i in range(8):
        lam = torch.zeros(n * N, device="cuda")
        mp.pcg_launch(n, N, S[i], P[i], g[i], lam, None, None, None, None, it, fl, cap, 1e-5)
        torch.cuda.synchronize()
        want.append((lam.clone(), int(it.item())))
    reps = bad = 0
    while reps < 3000 and time.time() < t_end:
        i = reps % 8
        lam = torch.zeros(n * N, device="cuda")
        mp.pcg_launch(n, N, S[i], P[i], g[i], lam, None, None, None, None, it, fl, cap, 1e-5)
        torch.cuda.synchronize()
        if not torch.equal(lam, want[i][0]) or int(it.item()) != want[i][1]:
            bad += 1
        reps += 1
    report[f"single N={N}"] = (reps, bad)
# batched: 256 systems, repeated
N, B = 128, 256
d = synth.make_systems(n, N, batch=B, seed=9)
S, P, g = (torch.from_numpy(d[k]).cuda() for k in ("S", "Pinv", "gamma"))
itb = torch.zeros(B, dtype=torch.int32, device="cuda")
flb = torch.zeros(B, dtype=torch.uint8, device="cuda")
lam0 = torch.zeros(B, n * N, device="cuda")
mp.solve_batched(n, N, B, S, P, g, lam0, itb, flb, 167, 1e-4)
torch.cuda.synchronize()
want_l, want_i = lam0.clone(), itb.clone()
reps = bad = 0
while reps < 100 and time.time() < t_end + 20:
    lam = torch.zeros(B, n * N, device="cuda")
    mp.solve_batched(n, N, B, S, P, g, lam, itb, flb, 167, 1e-4)
    torch.cuda.synchronize()
    if not torch.equal(lam, want_l) or not torch.equal(itb, want_i):
        bad += 1
    reps += 1
report["batched 256 x N=128"] = (reps, bad)
# direct solver
lamd = torch.zeros(n * N, device="cuda")
mp.solve_direct(n, N, S[0], g[0], lamd)
torch.cuda.synchronize()
want_d = lamd.clone()
reps = bad = 0
while reps < 1000 and time.time() < t_end + 40:
    lamd.zero_()
    mp.solve_direct(n, N, S[0], g[0], lamd)
    torch.cuda.synchronize()
    bad += int(not torch.equal(lamd, want_d))
    reps += 1
report["direct N=128"] = (reps, bad)
# round 2: batches whose iteration counts span 1 .. cap (right-hand sides over six decades), launched back to back without host
# synchronisation in between -- clusters draw systems from the work counter in a data-dependent order, epochs run across launches
import numpy as np  # noqa: E402
N, B = 32, 600
d = synth.make_systems(n, N, batch=B, seed=21)
scale = (10.0 ** np.random.default_rng(5).uniform(-4.0, 2.0, size=B)).astype(np.float32)
S, P = (torch.from_numpy(d[k]).cuda() for k in ("S", "Pinv"))
g = torch.from_numpy((d["gamma"] * scale[:, None]).astype(np.float32)).cuda()
itb = torch.zeros(B, dtype=torch.int32, device="cuda")
flb = torch.zeros(B, dtype=torch.uint8, device="cuda")
lam0 = torch.zeros(B, n * N, device="cuda")
mp.solve_batched(n, N, B, S, P, g, lam0, itb, flb, 173, 1e-4)
torch.cuda.synchronize()
want_l, want_i = lam0.clone(), itb.clone()
reps = bad = 0
lams = [torch.zeros(B, n * N, device="cuda") for _ in range(8)]
its = [torch.zeros(B, dtype=torch.int32, device="cuda") for _ in range(8)]
while reps < 400 and time.time() < t_end + 60:
    for k in range(8):
        lams[k].zero_()
        mp.solve_batched(n, N, B, S, P, g, lams[k], its[k], flb, 173, 1e-4)
    torch.cuda.synchronize()
    for k in range(8):
        bad += int(not torch.equal(lams[k], want_l) or not torch.equal(its[k], want_i))
    reps += 8
report[f"batched 600 x N=32, iterations {int(want_i.min())}..{int(want_i.max())}, 8 launches in flight"] = (reps, bad)
# config 5 on the whole GPU (cooperative grid kernel, packets through L2)
n5, N5 = 64, 256
d5 = synth.make_systems(n5, N5, batch=2, seed=4242)
S5, P5, g5 = (torch.from_numpy(d5[k]).cuda() for k in ("S", "Pinv", "gamma"))
lam5 = torch.zeros(n5 * N5, device="cuda")
mp.pcg_launch(n5, N5, S5[0], P5[0], g5[0], lam5, None, None, None, None, it, fl, 200, 1e-6)
torch.cuda.synchronize()
want5, want5_it = lam5.clone(), int(it.item())
reps = bad = 0
while reps < 300 and time.time() < t_end + 80:
    lam5.zero_()
    mp.pcg_launch(n5, N5, S5[0], P5[0], g5[0], lam5, None, None, None, None, it, fl, 200, 1e-6)
    torch.cuda.synchronize()
    bad += int(not torch.equal(lam5, want5) or int(it.item()) != want5_it)
    reps += 1
report["config 5 (n=64, N=256) grid kernel"] = (reps, bad)
print({k: f"{v[0]} runs, {v[1]} mismatches" for k, v in report.items()})
assert all(v[1] == 0 for v in report.values())
