#!/usr/bin/env python
"""Per-phase timeline of the fast cluster kernel from its timeline build (mode 22): %clock stamps of iterations 8..11,
reduced to mean / min / max cycles per interval over all threads.  Diagnostic tool; prints a table and writes
gpurun_out/timeline_fast_<N>_<C>.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpcgpu_b200 as m  # noqa: E402
from mpcgpu_b200 import _capi, synth  # noqa: E402

NAMES = ["updates done -> r stored, CTA barrier 1 passed", "-> u = Pinv r chain done, u stored", "-> CTA barrier 2 passed",
         "-> w = S u chain done", "-> halo sent + warp butterfly done", "(warp 0) -> warp partials polled, CTA pair sent",
         "butterfly done -> all packets seen (poll exit)", "-> totals", "-> exit test, beta, alpha", "-> p, s, lambda, r updates (next top)"]
NS = 10


def main():
    L = _capi.lib()
    n = 14
    for (N, C) in [(128, 16), (128, 8), (32, 4)]:
        v = [v for v in _capi.variants() if v["n"] == n and v["N"] == N and v["cluster"] == C and v["mode"] == 22]
        if not v:
            continue
        nt = v[0]["threads"]
        d = synth.make_systems(n, N, batch=1, seed=5)
        S, P, g = (torch.from_numpy(d[k][0]).cuda() for k in ("S", "Pinv", "gamma"))
        dbg = torch.zeros(4 * NS * C * nt, dtype=torch.int32, device="cuda")
        it = torch.zeros(1, dtype=torch.int32, device="cuda")
        fl = torch.zeros(1, dtype=torch.uint8, device="cuda")
        assert L.gbd_pcg_set_tuning(n, N, 0, C, 22) == 0
        L.gbd_pcg_set_debug_buffer(dbg.data_ptr())
        for _ in range(3):
            lam = torch.zeros(n * N, device="cuda")
            m.pcg_launch(n, N, S, P, g, lam, None, None, None, None, it, fl, 60, 1e-30)
        torch.cuda.synchronize()
        L.gbd_pcg_set_debug_buffer(None)
        L.gbd_pcg_set_tuning(n, N, 0, 0, -1)
        a = dbg.cpu().numpy().astype(np.int64).reshape(4, NS, C * nt)
        w0 = (np.arange(C * nt) % nt) < 32           # threads of warp 0 (the only ones that stamp point 6)

        def diff(x, y):
            return (x - y) & 0xFFFFFFFF

        ivals = [diff(a[:, 1], a[:, 0]), diff(a[:, 2], a[:, 1]), diff(a[:, 3], a[:, 2]), diff(a[:, 4], a[:, 3]), diff(a[:, 5], a[:, 4]),
                 diff(a[:, 6], a[:, 5])[:, w0], diff(a[:, 7], a[:, 5]), diff(a[:, 8], a[:, 7]), diff(a[:, 9], a[:, 8]),
                 diff(a[1:, 0], a[:-1, 9])]
        rows = []
        print(f"--- fast n={n} N={N} C={C} threads={nt} iters={int(it.item())}")
        for name, dt in zip(NAMES, ivals):
            rows.append(dict(interval=name, mean=float(dt.mean()), min=int(dt.min()), max=int(dt.max())))
            print(f"{name:58s} mean {dt.mean():7.1f}  min {dt.min():5d}  max {dt.max():5d}")
        per_iter = diff(a[1:, 0], a[:-1, 0])
        print(f"iteration (top to top)                                     mean {per_iter.mean():7.1f}  min {per_iter.min()}  max {per_iter.max()}")
        cta = 1 if C > 1 else 0
        print("per-warp interval means, CTA %d (columns = warps):" % cta)
        for i, name in enumerate(NAMES):
            if i in (5, 9):
                continue
            dt = ivals[i].reshape(4, C, nt)[:, cta, :].reshape(4, nt // 32, 32)
            print(f"  {name:56s}", np.round(dt.mean(axis=(0, 2))).astype(int))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"timeline_fast_{N}_{C}.json"), "w") as f:
            json.dump(dict(n=n, N=N, C=C, threads=nt, intervals=rows, iteration_cycles=float(per_iter.mean())), f, indent=1)


if __name__ == "__main__":
    main()
