#!/usr/bin/env python
"""Per-phase timeline of the fast BATCH kernel (gbd_cluster_pcg_fastb.cuh) from its timeline build (mode 29): %clock stamps of
iteration 9 held in registers.  Prints mean / min / max cycles per interval per thread role and writes
gpurun_out/timeline_fastb_<N>_<C>.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpcgpu_b200 as m  # noqa: E402
from mpcgpu_b200 import _capi, synth  # noqa: E402

NS = 16
SEGS = [("top -> updates, r stored, CTA barrier A passed", 0, 1, "all"),
        ("P-threads: u = Pinv r (4 rows), u + r.u stored", 1, 2, "P"),
        ("   P: barrier A passed -> chain entered", 1, 12, "P"), ("   P: the chain proper", 12, 13, "P"), ("   P: u stored, r.u parked", 13, 2, "P"),
        ("   S: barrier B passed -> chain entered", 3, 14, "S"), ("   S: the chain proper", 14, 15, "S"), ("   S: w.u parked", 15, 4, "S"),
        ("halo lanes: near-halo u from staged tiles", 1, 2, "halo"),
        ("S-threads (non-halo): idle until barrier B passed", 1, 3, "S"),
        ("P-threads: wait at barrier B", 2, 3, "P"),
        ("S-threads: w = S u (4 rows), w.u parked", 3, 4, "S"),
        ("S-threads: arrive, boundary rows sent", 4, 5, "S"),
        ("S-threads: asleep until scalars published", 5, 9, "S"),
        ("exchange warp: r.u summed", 3, 4, "X"),
        ("exchange warp: wait for the S-threads (named barrier 1)", 4, 5, "X"),
        ("exchange warp: w.u summed, pair sent", 5, 6, "X"),
        ("exchange warp: pair sent -> all C pairs seen", 6, 7, "X"),
        ("exchange warp: totals, exit test, beta, alpha, stored", 7, 8, "X"),
        ("exchange warp: barrier 2 arrive", 8, 9, "X"),
        ("halo lanes: woken -> both w halo packets seen", 9, 10, "halo"),
        ("woken -> end of step", 9, 11, "all"),
        ("whole iteration: top -> end of step", 0, 11, "all")]


def main():
    L = _capi.lib()
    n = 14
    for (N, C) in [(128, 4), (32, 1)]:
        v = [v for v in _capi.variants() if v["n"] == n and v["N"] == N and v["cluster"] == C and v["mode"] == 29]
        if not v:
            continue
        nt = v[0]["threads"]
        NP = nt // 2
        d = synth.make_systems(n, N, batch=1, seed=5)
        S, P, g = (torch.from_numpy(d[k][0]).cuda() for k in ("S", "Pinv", "gamma"))
        dbg = torch.zeros(NS * C * nt, dtype=torch.int32, device="cuda")
        it = torch.zeros(1, dtype=torch.int32, device="cuda")
        fl = torch.zeros(1, dtype=torch.uint8, device="cuda")
        assert L.gbd_pcg_set_tuning(n, N, 0, C, 29) == 0
        L.gbd_pcg_set_debug_buffer(dbg.data_ptr())
        for _ in range(3):
            lam = torch.zeros(n * N, device="cuda")
            m.pcg_launch(n, N, S, P, g, lam, None, None, None, None, it, fl, 60, 1e-30)
        torch.cuda.synchronize()
        L.gbd_pcg_set_debug_buffer(None)
        L.gbd_pcg_set_tuning(n, N, 0, 0, -1)
        a = dbg.cpu().numpy().astype(np.int64).reshape(NS, C, nt)
        tid = np.arange(nt)
        sel = {"all": np.ones(nt, bool), "P": tid < NP, "S": (tid >= NP) & ~((tid >= NP) & (tid < NP + 2 * n)),
               "X": tid < 32, "halo": (tid >= NP) & (tid < NP + 2 * n)}
        rows = []
        print(f"--- fastb n={n} N={N} C={C} threads={nt} iters={int(it.item())}")
        for name, p0, p1, who in SEGS:
            dt = ((a[p1] - a[p0]) & 0xFFFFFFFF)[:, sel[who]]
            rows.append(dict(interval=name, mean=float(dt.mean()), min=int(dt.min()), max=int(dt.max())))
            print(f"{name:62s} mean {dt.mean():7.1f}  min {dt.min():6d}  max {dt.max():6d}")
        cta = 1 if C > 1 else 0
        print("per-warp means, CTA %d (columns = warps):" % cta)
        for name, p0, p1 in (("0->1", 0, 1), ("1->2", 1, 2), ("2->3", 2, 3), ("3->4", 3, 4), ("4->5", 4, 5), ("5->9", 5, 9), ("9->11", 9, 11), ("0->11", 0, 11)):
            dt = ((a[p1] - a[p0]) & 0xFFFFFFFF)[cta].reshape(nt // 32, 32)
            print(f"  {name:8s}", np.round(dt.mean(axis=1)).astype(int))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"timeline_fastb_{N}_{C}.json"), "w") as f:
            json.dump(dict(n=n, N=N, C=C, threads=nt, intervals=rows), f, indent=1)


if __name__ == "__main__":
    main()
