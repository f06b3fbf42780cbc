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
    if os.environ.get("GBD_PCG_ZEROCOPY") == "1":
        # where the time goes: the same kernel launched through the device-pointer entry (CUDA events, back to back) with its tiles
        # (a) resident in HBM, (b) in pinned host memory (read by the kernel's own TMA loads over PCIe; lambda in HBM)
        from mpcgpu_b200 import _capi
        L = _capi.lib()
        dS, dP, dg = (t.cuda() for t in (hS, hP, hg))
        lam = torch.zeros(64, n * N, device="cuda")
        it = torch.zeros(64, dtype=torch.int32, device="cuda")
        fl = torch.zeros(64, dtype=torch.uint8, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for name, (xS, xP, xg) in (("tiles in HBM", (dS, dP, dg)), ("tiles in pinned host memory", (hS, hP, hg))):
            def go(i):
                rc = L.gbd_pcg_solve_f32(n, N, xS[i].data_ptr(), xP[i].data_ptr(), xg[i].data_ptr(), lam[i].data_ptr(), 0, 0, 0, 0,
                                         it[i:].data_ptr(), fl[i:].data_ptr(), 167, 1e-4, st)
                assert rc == 0
            for i in range(16):
                go(i)
            lam.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(64):
                go(i)
            e1.record()
            torch.cuda.synchronize()
            print(f"   kernel, {name}: {1e3 * e0.elapsed_time(e1) / 64:.1f} us per solve (CUDA events, 64 back-to-back launches, mean iters {it.float().mean().item():.1f})")
        # one synchronous launch + wait, tiles in HBM: launch + completion latency on top of the kernel
        lam.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(64):
            L.gbd_pcg_solve_f32(n, N, dS[i].data_ptr(), dP[i].data_ptr(), dg[i].data_ptr(), lam[i].data_ptr(), 0, 0, 0, 0,
                                it[i:].data_ptr(), fl[i:].data_ptr(), 167, 1e-4, st)
            torch.cuda.synchronize()
        print(f"   launch + cudaStreamSynchronize per solve, tiles in HBM: {1e6 * (time.perf_counter() - t0) / 64:.1f} us")
else:
    for mode in ("1", "0", "2"):
        env = dict(os.environ, GBD_PCG_ZEROCOPY=mode)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
