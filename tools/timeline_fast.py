#!/usr/bin/env python
"""Per-phase timeline of the fast cluster kernel from its timeline build (mode 22): every thread keeps %clock stamps of ONE
iteration (iteration 9) in registers and writes them after the solve, so the stamped iteration runs the same instruction
stream as the others.  Prints mean / min / max cycles per interval over all threads and writes
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

NS = 12
# stamp points (own-row warps): 0 top | 1 r stored, barrier 1 arrival | 2 u chain done, u + r.u stored | 3 barrier 2 arrival | 4 w chain done
#   5 w.u parked, arrived at named barrier 1, boundary rows sent | 9 woken by named barrier 2 (scalars published) | 10 scalars read
# halo warp: 4 r.u reduced | 5 named barrier 1 passed (w.u parked) | 6 w.u reduced, CTA pair sent | 7 poll exit | 8 totals | 9 scalars published
SEGS = [("top -> updates, r stored, CTA barrier 1", 0, 1, None), ("-> (barrier wait +) u = Pinv r chain, u stored", 1, 2, None),
        ("-> (barrier wait +) w = S u chain", 3, 4, "own"), ("-> w.u parked, barrier arrive, boundary rows sent", 4, 5, "own"),
        ("own warps asleep until the scalars are published", 5, 9, "own"), ("-> scalars read (loop top of the next iteration)", 9, 10, "own"),
        ("halo warp: barrier 2 arrival -> r.u reduced", 3, 4, "halo"), ("halo warp: -> named barrier passed (w.u parked)", 4, 5, "halo"),
        ("halo warp: -> w.u reduced, CTA pair sent", 5, 6, "halo"), ("halo warp: pair sent -> all C pairs seen", 6, 11, "halo"),
        ("halo warp: pair sent -> pairs AND boundary rows seen (poll exit)", 6, 7, "halo"),
        ("halo warp: -> totals", 7, 8, "halo"), ("halo warp: -> exit test, beta, alpha, published", 8, 9, "halo"),
        ("whole iteration: top -> scalars known", 0, 10, None)]


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
        dbg = torch.zeros(NS * C * nt, dtype=torch.int32, device="cuda")
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
        a = dbg.cpu().numpy().astype(np.int64).reshape(NS, C, nt)
        halo = np.zeros((C, nt), bool)
        halo[:, nt - 32:] = True
        rows = []
        print(f"--- fast n={n} N={N} C={C} threads={nt} iters={int(it.item())}")
        for name, p0, p1, sel in SEGS:
            dt = (a[p1] - a[p0]) & 0xFFFFFFFF
            dt = dt[halo] if sel == "halo" else (dt[~halo] if sel == "own" else dt.reshape(-1))
            rows.append(dict(interval=name, mean=float(dt.mean()), min=int(dt.min()), max=int(dt.max())))
            print(f"{name:52s} mean {dt.mean():7.1f}  min {dt.min():5d}  max {dt.max():5d}")
        cta = 1 if C > 1 else 0
        print("per-warp means, CTA %d (columns = warps):" % cta)
        for name, p0, p1, sel in SEGS:
            if sel:
                continue
            dt = ((a[p1] - a[p0]) & 0xFFFFFFFF)[cta].reshape(nt // 32, 32)
            print(f"  {name:50s}", np.round(dt.mean(axis=1)).astype(int))
        # skew between CTAs: when does each CTA's warp 0 send, when does each CTA leave the poll (relative to the earliest top)
        t0 = a[0].min()
        print("per-CTA (halo warp): top / pair sent / poll exit / published, cycles after the earliest top")
        for nm, pt in (("top", 0), ("pair sent", 6), ("pairs seen", 7), ("published", 9), ("halo seen", 11)):
            print(f"  {nm:10s}", (a[pt][:, nt - 1] - t0))
        print("  halo sent (warp 0 / last own warp)", (a[5][:, 0] - t0), (a[5][:, nt - 33] - t0))
        print("  own top   ", (a[0][:, 0] - t0))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"timeline_fast_{N}_{C}.json"), "w") as f:
            json.dump(dict(n=n, N=N, C=C, threads=nt, intervals=rows), f, indent=1)


if __name__ == "__main__":
    main()
