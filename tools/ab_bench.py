#!/usr/bin/env python
"""A/B timing on one GPU: every compiled variant of our kernel and (when oracle/_ref is present) the
UNMODIFIED reference pcg<> kernel, same inputs, CUDA-event kernel time + the reference's stopwatch
window.  Test-side tool (it may use oracle/); writes gpurun_out/ab_bench.json.  Not part of bench.py."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpcgpu_b200 as m  # noqa: E402
from mpcgpu_b200 import _capi, synth  # noqa: E402
from oracle import refgpu  # noqa: E402

CAPS = {32: 173, 64: 167, 128: 167, 256: 118, 512: 67}
SHAPES = [(14, 32), (14, 64), (14, 128), (14, 256), (14, 512), (64, 256)]
if os.environ.get('AB_SHAPES'):
    SHAPES = [tuple(int(x) for x in t.split('x')) for t in os.environ['AB_SHAPES'].split(',')]
MODES = [int(x) for x in os.environ['AB_MODES'].split(',')] if os.environ.get('AB_MODES') else None


def time_fn(fn, reps, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps      # us


def main():
    L = _capi.lib()
    out = []
    ring = 64
    for (n, N) in SHAPES:
        for tol in ((1e-4,) if os.environ.get('AB_QUICK') else (1e-4, 1e-6)):
            cap = CAPS[N] if n == 14 else 200          # config 5 (n=64, N=256): cap 200
            ring = 64 if n == 14 else 4
            d = synth.make_systems(n, N, batch=ring, seed=77)
            S, P, g = (torch.from_numpy(d[k]).cuda() for k in ("S", "Pinv", "gamma"))
            lam = torch.zeros(ring, n * N, device="cuda")
            it = torch.zeros(ring, dtype=torch.int32, device="cuda")
            fl = torch.zeros(ring, dtype=torch.uint8, device="cuda")
            st = torch.cuda.current_stream().cuda_stream
            state = {"i": 0}

            def ours():
                i = state["i"] = (state["i"] + 1) % ring
                lam[i].zero_()
                rc = L.gbd_pcg_solve_f32(n, N, S[i].data_ptr(), P[i].data_ptr(), g[i].data_ptr(), lam[i].data_ptr(), 0, 0,
                                         0, 0, it[i:].data_ptr(), fl[i:].data_ptr(), cap, tol, st)
                assert rc == 0, rc

            def zero_only():
                i = state["i"] = (state["i"] + 1) % ring
                lam[i].zero_()

            t_zero = time_fn(zero_only, 200)
            for v in [v for v in _capi.variants() if v["n"] == n and v["N"] == N and not v["f64"] and (MODES is None or v["mode"] in MODES)]:
                assert L.gbd_pcg_set_tuning(n, N, 0, v["cluster"], v["mode"]) == 0
                us = time_fn(ours, 200) - t_zero
                torch.cuda.synchronize()
                mean_it = float(it.float().mean().item())
                out.append(dict(impl="ours", n=n, N=N, tol=tol, cap=cap, cluster=v["cluster"], mode=v["mode"],
                                kernel_us=us, mean_iters=mean_it, us_per_iter=us / mean_it))
                print(out[-1], flush=True)
                L.gbd_pcg_set_tuning(n, N, 0, 0, -1)
            if refgpu.available() and not os.environ.get('AB_NOREF'):
                ws = refgpu.RefWorkspace(n, N)

                def ref():
                    i = state["i"] = (state["i"] + 1) % ring
                    lam[i].zero_()
                    refgpu.launch(n, N, S[i], P[i], g[i], lam[i], ws, cap, tol, 128, st)

                us = time_fn(ref, 100) - t_zero
                # its iteration counts (same as ours by bit-parity; read back the last one)
                out.append(dict(impl="reference_gbdpcg", n=n, N=N, tol=tol, cap=cap, kernel_us=us,
                                last_iters=int(ws.iters.item()), us_per_iter=us / max(1, int(ws.iters.item()))))
                print(out[-1], flush=True)
                # stopwatch windows (sqp.cuh:224-241)
                w_ref, w_ours = [], []
                r = torch.zeros(n * N, device="cuda")
                p = torch.zeros(n * N, device="cuda")
                for k in range(40):
                    i = k % ring
                    lam[i].zero_()
                    w_ref.append(refgpu.linsys_window(n, N, S[i], P[i], g[i], lam[i], ws, cap, tol)[2])
                    lam[i].zero_()
                    w_ours.append(m.linsys_window(n, N, S[i], P[i], g[i], lam[i], r, p, it[:1], fl[:1], cap, tol)[2])
                out.append(dict(impl="windows", n=n, N=N, tol=tol, ref_window_us_median=float(np.median(w_ref[5:])),
                                ours_window_us_median=float(np.median(w_ours[5:]))))
                print(out[-1], flush=True)
    # batched kernels: every 2-CTA/SM build (modes 3 and 6) and the 1-CTA/SM builds at the same cluster size
    for (n, N, B) in [(14, 128, 1024), (14, 32, 1024)] + ([(14, 64, 1024), (14, 256, 512), (14, 512, 256)] if os.environ.get('AB_ALLBATCH') else []):
        cap, tol = CAPS[N], 1e-4
        d = synth.make_systems(n, N, batch=B, seed=1234)
        S, P, g = (torch.from_numpy(d[k]).cuda() for k in ("S", "Pinv", "gamma"))
        lam = torch.zeros(B, n * N, device="cuda")
        it = torch.zeros(B, dtype=torch.int32, device="cuda")
        fl = torch.zeros(B, dtype=torch.uint8, device="cuda")

        def batched():
            lam.zero_()
            m.solve_batched(n, N, B, S, P, g, lam, it, fl, cap, tol)

        for v in [v for v in _capi.variants() if v["n"] == n and v["N"] == N and not v["f64"] and v["mode"] in (MODES or (2, 3, 5, 6, 7, 8, 11, 12, 13))]:
            assert L.gbd_pcg_set_tuning(n, N, 0, v["cluster"], v["mode"]) == 0
            us = time_fn(batched, 5, warm=2)
            out.append(dict(impl="ours_batched", n=n, N=N, batch=B, cluster=v["cluster"], mode=v["mode"], ms=us / 1e3,
                            systems_per_s=B / (us * 1e-6), mean_iters=float(it.float().mean().item())))
            print(out[-1], flush=True)
            L.gbd_pcg_set_tuning(n, N, 0, 0, -1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ab_bench.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
