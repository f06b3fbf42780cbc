#!/usr/bin/env python
"""Summarise the closed-loop MPC runs written by tools/gpu_closed_loop.sh (gpurun_out/cl_<arm>_<mode>_<knots>.bin) into
profiles/r02_closed_loop.json.  Arms: ref = reference GBD-PCG headers, dropin = include/gbd_dropin (bit-exact bodies), fast =
include/gbd_dropin with -DGBD_DROPIN_FAST=1, direct = include/gbd_dropin with -DGBD_DROPIN_DIRECT=1 (block cyclic reduction instead of PCG), refp = reference headers with pcg_exit_tol x 1.001 (the experiment's noise floor).
Modes: b = behaviour build (TIME_LINSYS=0: SQP iterations per control step), t = timing build (the reference's linsys stopwatch)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out")


def load(arm, mode, knots):
    p = os.path.join(SRC, f"cl_{arm}_{mode}_{knots}.bin")
    if not os.path.exists(p):
        return None
    raw = np.fromfile(p, np.uint8)
    ca, cb = np.frombuffer(raw[:8], np.uint32)
    if mode == "t":
        top = np.frombuffer(raw[8:8 + 8 * ca], np.float64)
        off = 8 + 8 * ca
    else:
        top = np.frombuffer(raw[8:8 + 4 * ca], np.uint32)
        off = 8 + 4 * ca
    err = np.frombuffer(raw[off:off + 4 * cb], np.float32)
    return top, err


out = {"what": "closed-loop MPC: the reference's simulateMPC (include/mpcsim.cuh:146-149, unchanged) around sqpSolvePcg, IIWA track "
               "examples/trajfiles/0_0, built against four header sets (oracle/Makefile closed_loop); SQP_MAX_TIME_US lifted so the "
               "SQP exit does not depend on the wall clock", "runs": []}
for knots, tol in ((32, 5e-6), (128, 1e-4)):
    for mode in ("b", "t"):
        ref = load("ref", mode, knots)
        if ref is None:
            continue
        for arm in ("ref", "dropin", "fast", "direct", "refp"):
            d = load(arm, mode, knots)
            if d is None:
                continue
            top, err = d
            r = {"knot_points": knots, "pcg_exit_tol": tol * (1.001 if arm == "refp" else 1.0), "arm": arm,
                 "build": "behaviour (TIME_LINSYS=0)" if mode == "b" else "timing (TIME_LINSYS=1)", "control_steps": int(err.size),
                 "tracking_error_mean": float(err.astype(np.float64).mean()), "tracking_error_final": float(err[-1]),
                 "tracking_error_mean_vs_ref": float(err.astype(np.float64).mean() / ref[1].astype(np.float64).mean())}
            if mode == "b":
                r.update(sqp_calls=int(top.size), sqp_iters_total=int(top.sum()), sqp_iters_mean=float(top.mean()),
                         sqp_iters_total_vs_ref=float(top.sum() / ref[0].sum()),
                         identical_to_ref=bool(top.size == ref[0].size and np.array_equal(top, ref[0]) and np.array_equal(err, ref[1])))
            else:
                r.update(linsys_calls=int(top.size), linsys_us_mean=float(top.mean()), linsys_us_median=float(np.median(top)),
                         linsys_speedup_vs_ref=float(ref[0].mean() / top.mean()),
                         identical_tracking_to_ref=bool(np.array_equal(err, ref[1])))
            out["runs"].append(r)
with open(os.path.join(ROOT, "profiles", "r02_closed_loop.json"), "w") as f:
    json.dump(out, f, indent=1)
for r in out["runs"]:
    print({k: (round(v, 5) if isinstance(v, float) else v) for k, v in r.items() if k not in ("build",)})
