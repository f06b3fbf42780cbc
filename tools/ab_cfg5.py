#!/usr/bin/env python
"""Config 5 (n=64, N=256, tol 1e-6, cap 200): every compiled grid variant against the C oracle (bit-exact) and timed."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpcgpu_b200 as m  # noqa: E402
from mpcgpu_b200 import _capi, synth  # noqa: E402
from oracle import pcg as oracle_pcg  # noqa: E402  (checker only)

n, N, cap, tol = 64, 256, 200, 1e-6
L = _capi.lib()
d = synth.make_systems(n, N, batch=2, seed=4242)
want = [oracle_pcg.pcg(d["S"][i], d["Pinv"][i], d["gamma"][i], d["lambda0"][i], n, N, cap, tol) for i in range(2)]
S, P, g = (torch.from_numpy(d[k]).cuda() for k in ("S", "Pinv", "gamma"))
it = torch.zeros(1, dtype=torch.int32, device="cuda")
fl = torch.zeros(1, dtype=torch.uint8, device="cuda")
for v in [v for v in _capi.variants() if v["n"] == n and v["N"] == N and not v["f64"]]:
    assert L.gbd_pcg_set_tuning(n, N, 0, v["cluster"], v["mode"]) == 0
    ok = True
    for i in range(2):
        lam = torch.zeros(n * N, device="cuda")
        r, p = torch.zeros(n * N, device="cuda"), torch.zeros(n * N, device="cuda")
        m.pcg_launch(n, N, S[i], P[i], g[i], lam, r, p, None, None, it, fl, cap, tol)
        torch.cuda.synchronize()
        ok = ok and int(it.item()) == want[i]["iters"] and np.array_equal(lam.cpu().numpy(), want[i]["lam"]) \
            and np.array_equal(r.cpu().numpy(), want[i]["r"]) and np.array_equal(p.cpu().numpy(), want[i]["p"])
    lams = torch.zeros(20, n * N, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for q in range(4):
        m.pcg_launch(n, N, S[q % 2], P[q % 2], g[q % 2], lams[q], None, None, None, None, it, fl, cap, tol)
    torch.cuda.synchronize()
    e0.record()
    for q in range(4, 20):
        m.pcg_launch(n, N, S[q % 2], P[q % 2], g[q % 2], lams[q], None, None, None, None, it, fl, cap, tol)
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / 16
    mit = 0.5 * (want[0]["iters"] + want[1]["iters"])
    print({"ctas": v["cluster"], "mode": v["mode"], "bit_exact_vs_oracle": bool(ok), "kernel_us": us, "mean_iters": mit,
           "us_per_iter": us / mit}, flush=True)
    L.gbd_pcg_set_tuning(n, N, 0, 0, -1)
