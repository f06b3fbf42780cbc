#!/usr/bin/env python
"""Host-buffer path (gbd_pcg_plan_solve_host_f32) under the three GBD_PCG_ZEROCOPY settings, one process each: the e2e leg of bench.py
in isolation (pinned inputs, one solve per call, IIWA ring when oracle/_ref is present).  Prints microseconds per solve."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import time
    import numpy as np
    import torch
    import mpcgpu_b200 as m
    from bench import load_ring
    n, N = 14, 128
    host, data = load_ring(N, 64)
    hS, hP, hg = (torch.from_numpy(host[k][:64]).pin_memory() for k in ("S", "Pinv", "gamma"))
    hl = torch.zeros(64, n * N).pin_memory()
    plan = m.HostPlan(n, N, 1)
    hl_np = hl.numpy()
    pS, pP, pg, pl = ([int(t[i].data_ptr()) for i in range(64)] for t in (hS, hP, hg, hl))
    def step(s):
        i = s % 64
        hl_np[i].fill(0.0)
        return plan.solve_raw(pS[i], pP[i], pg[i], pl[i], 167, 1e-4)
    for s in range(50):
        step(s)
    t0 = time.perf_counter()
    K = 1500
    its = 0
    for s in range(K):
        its += step(s)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"GBD_PCG_ZEROCOPY={os.environ.get('GBD_PCG_ZEROCOPY', '(default 1)')} data={data}: {1e6 * dt / K:.1f} us per solve, mean iters {its / K:.1f}")
else:
    for mode in ("1", "0", "2"):
        env = dict(os.environ, GBD_PCG_ZEROCOPY=mode)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
