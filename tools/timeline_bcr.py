#!/usr/bin/env python
"""Timeline of the block-cyclic-reduction direct solver (timeline build of bcr_cluster_kernel): %clock stamps per warp in
program order -> cycles per phase and level.  Diagnostic tool."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpcgpu_b200 import _capi, synth  # noqa: E402

L = _capi.lib()
n = 14
for N, C in ((128, 16), (32, 4)):
    levels = int(np.log2(N))
    nst = 1 + 4 * levels + 2 + 2 * levels + 1
    d = synth.make_systems(n, N, batch=1, seed=5)
    S, g = torch.from_numpy(d["S"][0]).cuda(), torch.from_numpy(d["gamma"][0]).cuda()
    lam = torch.zeros(n * N, device="cuda")
    dbg = torch.zeros(64 * C * 4, dtype=torch.int32, device="cuda")
    L.gbd_pcg_set_debug_buffer(dbg.data_ptr())
    for _ in range(3):
        assert L.gbd_bcr_solve_f32(n, N, S.data_ptr(), g.data_ptr(), lam.data_ptr(), 0) == 0
    torch.cuda.synchronize()
    L.gbd_pcg_set_debug_buffer(None)
    a = dbg.cpu().numpy().astype(np.int64).reshape(64, C, 4)
    print(f"--- BCR n={n} N={N} C={C}: CTA 1, cycles between consecutive stamps (max over its 4 warps of the arrival)")
    names = ["load"]
    for l in range(levels):
        names += [f"L{l} phase1 (GJ + W)", f"L{l} cluster_sync", f"L{l} phase2 (absorb)", f"L{l} cta sync"]
    names += ["root", "cluster_sync"]
    for l in range(levels):
        names += [f"back {l}", f"back {l} cluster_sync"]
    cta = 1 if C > 1 else 0
    for i in range(1, min(nst, 63)):
        dt = (a[i, cta] - a[i - 1, cta]) & 0xFFFFFFFF
        print(f"{names[i]:28s} per-warp {dt}")
    tot = (a[nst - 1, cta, 0] - a[0, cta, 0]) & 0xFFFFFFFF
    print("total cycles from after-load to before-output:", tot)
