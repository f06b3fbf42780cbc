#!/usr/bin/env python
"""Tiny driver for ncu: a few direct solves (n=14, N=128) and one Schur assembly."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpcgpu_b200 as mp  # noqa: E402
from mpcgpu_b200 import synth  # noqa: E402

n, m, N = 14, 7, 128
d = synth.make_systems(n, N, batch=4, seed=3)
S, g = torch.from_numpy(d["S"]).cuda(), torch.from_numpy(d["gamma"]).cuda()
lam = torch.zeros(n * N, device="cuda")
for i in range(6):
    mp.solve_direct(n, N, S[i % 4], g[i % 4], lam)
G, C, gg, c = (torch.from_numpy(x).cuda().reshape(-1) for x in synth.make_kkt_batch(n, m, N, 1, seed=1))
dS, dP, dgam = torch.zeros(3 * n * n * N, device="cuda"), torch.zeros(3 * n * n * N, device="cuda"), torch.zeros(n * N, device="cuda")
for i in range(4):
    mp.form_schur_system(n, m, N, G.clone(), C, gg, c, dS, dP, dgam, 1e-3)
torch.cuda.synchronize()
print("done")
