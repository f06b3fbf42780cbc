#!/usr/bin/env python
"""Per-phase timeline of the v4 kernel from its timeline build (mode 10): %clock stamps of iterations 8..11,
reduced to mean / min / max cycles per interval over all threads.  Diagnostic tool; prints a table and writes
gpurun_out/timeline_<N>_<C>.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpcgpu_b200 as m  # noqa: E402
from mpcgpu_b200 import _capi, synth  # noqa: E402

NAMES_V2 = ["top sync -> chain S.p done", "-> per-knot tree, partial staged", "-> ship (named barrier, sender st.async)", "-> mbarrier wait over",
            "-> N-tree + alpha", "-> r update + mid sync", "-> chain Pinv.r done", "-> tree + ship", "-> mbarrier wait over",
            "-> N-tree (eta')", "-> (exit test, beta, p update, next top sync)"]
NAMES = ["top sync -> chain S.p done", "-> partial+edge sent (14-tree)", "-> all packets seen (poll exit)",
         "-> N-tree + alpha", "-> r update + mid sync", "-> chain Pinv.r done", "-> sent", "-> poll exit",
         "-> N-tree (eta')", "-> beta + p update", "-> next top sync"]
PTS = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10]


def main():
    L = _capi.lib()
    n = 14
    for (N, C, mode) in [(32, 4, 10), (128, 8, 10), (128, 16, 14), (128, 8, 14)]:
        v = [v for v in _capi.variants() if v["n"] == n and v["N"] == N and v["cluster"] == C and v["mode"] == mode]
        if not v:
            continue
        names = NAMES if mode == 10 else NAMES_V2
        nt = v[0]["threads"]
        d = synth.make_systems(n, N, batch=1, seed=5)
        S, P, g = (torch.from_numpy(d[k][0]).cuda() for k in ("S", "Pinv", "gamma"))
        dbg = torch.zeros(4 * 12 * C * nt, dtype=torch.int32, device="cuda")
        it = torch.zeros(1, dtype=torch.int32, device="cuda")
        fl = torch.zeros(1, dtype=torch.uint8, device="cuda")
        assert L.gbd_pcg_set_tuning(n, N, 0, C, mode) == 0
        L.gbd_pcg_set_debug_buffer(dbg.data_ptr())
        for _ in range(3):
            lam = torch.zeros(n * N, device="cuda")
            m.pcg_launch(n, N, S, P, g, lam, None, None, None, None, it, fl, 60, 1e-30)
        torch.cuda.synchronize()
        L.gbd_pcg_set_debug_buffer(None)
        L.gbd_pcg_set_tuning(n, N, 0, 0, -1)
        a = dbg.cpu().numpy().astype(np.int64).reshape(4, 12, C * nt)
        rows = []
        print(f"--- {'v4' if mode == 10 else 'v2'} n={n} N={N} C={C} threads={nt} iters={int(it.item())}")
        for i, name in enumerate(names):
            if i < 10:
                dt = (a[:, PTS[i + 1]] - a[:, PTS[i]]) & 0xFFFFFFFF
            else:
                dt = (a[1:, 0] - a[:-1, 10]) & 0xFFFFFFFF
            rows.append(dict(interval=name, mean=float(dt.mean()), min=int(dt.min()), max=int(dt.max())))
            print(f"{name:36s} mean {dt.mean():7.1f}  min {dt.min():5d}  max {dt.max():5d}")
        per_iter = ((a[1:, 0] - a[:-1, 0]) & 0xFFFFFFFF)
        print(f"iteration (top sync to top sync)     mean {per_iter.mean():7.1f}  min {per_iter.min()}  max {per_iter.max()}")
        # per-warp means of every interval, CTA 1 (an interior CTA), averaged over the 4 iterations
        cta = 1 if C > 1 else 0
        print("per-warp interval means, CTA %d (columns = warps):" % cta)
        for i in range(10):
            dt = ((a[:, PTS[i + 1]] - a[:, PTS[i]]) & 0xFFFFFFFF).reshape(4, C, nt)[:, cta, :].reshape(4, nt // 32, 32)
            print(f"  {names[i]:34s}", np.round(dt.mean(axis=(0, 2))).astype(int), " lane spread", int((dt.max(axis=2) - dt.min(axis=2)).max()))
        # who is last?  per-warp poll wait of phase A in iteration 9, CTA 0
        w = ((a[1, 3] - a[1, 2]) & 0xFFFFFFFF).reshape(C, nt)[:, ::32]
        print("phase-A poll wait per warp (rows = CTA):")
        print(w)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"timeline_{'v4' if mode == 10 else 'v2'}_{N}_{C}.json"), "w") as f:
            json.dump(dict(n=n, N=N, C=C, threads=nt, intervals=rows, iteration_cycles=float(per_iter.mean())), f, indent=1)


if __name__ == "__main__":
    main()
