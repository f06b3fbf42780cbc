#!/usr/bin/env python
"""A/B timing on one GPU of rows f1 / f2: our form_schur_system + compute_dz against the reference's own kernels
(oracle/_ref/libref_schur.so), same inputs, CUDA events.  Test-side tool (uses oracle/); writes gpurun_out/ab_schur.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpcgpu_b200 as mp  # noqa: E402
from oracle import refgpu, schur  # noqa: E402


def timed(fn, reps=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


def main():
    out = []
    n, m = 14, 7
    for N in (32, 128, 512):
        G, C, g, c = schur.make_kkt(n, m, N, seed=N)
        G0 = torch.from_numpy(G).cuda()
        dG, dC, dg, dc = (torch.from_numpy(x.copy()).cuda() for x in (G, C, g, c))
        dS = torch.zeros(3 * n * n * N, device="cuda")
        dP = torch.zeros_like(dS)
        dgam = torch.zeros(n * N, device="cuda")
        lam = torch.randn(n * N, device="cuda")
        dz = torch.zeros((n + m) * (N - 1) + n, device="cuda")
        t_copy = timed(lambda: dG.copy_(G0))

        from mpcgpu_b200 import _capi
        L = _capi.lib()
        st = torch.cuda.current_stream().cuda_stream
        pG, pC, pg, pc, pS, pP, pgam, plam, pdz = (int(x.data_ptr()) for x in (dG, dC, dg, dc, dS, dP, dgam, lam, dz))

        def ours():                                   # raw C-ABI calls: no Python-side argument checking in the timed loop
            dG.copy_(G0)
            L.gbd_form_schur_system_f32(n, m, N, pG, pC, pg, pc, pS, pP, pgam, 1e-3, st)

        def ref():
            dG.copy_(G0)
            refgpu.form_schur_system(n, m, N, dG, dC, dg, dc, dS, dP, dgam, 1e-3)

        t_ours = timed(ours) - t_copy
        t_dz = timed(lambda: L.gbd_compute_dz_f32(n, m, N, pG, pC, pg, plam, pdz, st))
        rec = dict(n=n, m=m, N=N, ours_form_schur_us=t_ours, ours_compute_dz_us=t_dz)
        if refgpu.schur_available():
            rec["ref_form_schur_us"] = timed(ref, reps=100) - t_copy
            rec["ref_compute_dz_us"] = timed(lambda: refgpu.compute_dz(n, m, N, dG, dC, dg, lam, dz), reps=100)
        byt = 4 * ((n * n + m * m) * N * 2 + (n * n + n * m) * N + (2 * n + m) * N + 2 * 3 * n * n * N + n * N)
        rec["algorithmic_bytes"] = byt
        rec["ours_form_schur_gbs"] = byt / (t_ours * 1e-6) / 1e9
        out.append(rec)
        print(rec, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ab_schur.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
