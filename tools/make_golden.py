#!/usr/bin/env python
"""Mints tests/golden/*.npz ON THE B200 BOX (gpurun) -- the generating script the fixtures were made by.

Inputs come from the REFERENCE: oracle/_ref/ref_capture_N (the reference's own generate_kkt_submatrices +
form_schur_system, include/pcg/linsys_setup.cuh:565-657, compiled from the reference headers where they
lie) run on the reference trajectory examples/trajfiles/0_0_* at rho = 1e-3.  Answers come from the
REFERENCE kernel: oracle/_ref/libref_gbdpcg.so (the unmodified GBD-PCG/include/pcg.cuh for sm_100a, launched
as include/pcg/sqp.cuh:230 does).  Nothing of this repo's product is on that path.

One fixture per (knot_points, offset): S, Pinv (pad tiles as the reference leaves them: 0xFF bytes = NaN),
gamma, and per (tol, cap) the reference kernel's lambda, r, p, iters, max_iter_exit from lambda0 = 0.
Written to gpurun_out/golden/ (merged back by gpurun); copy into tests/golden/ and commit.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(ROOT, "gpurun_out", "golden")

# (knot_points, [offsets], [(tol, cap)]): caps are settings.cuh:123-138; tolerances from track_iiwa_pcg.cu:62-68
CASES = [(32, [0, 200], [(1e-4, 173), (1e-6, 173)]),
         (128, [0, 300], [(1e-4, 167), (1e-6, 167)]),
         (512, [0], [(1e-4, 67)])]


def main():
    import torch
    from oracle import refgpu
    os.makedirs(OUT, exist_ok=True)
    n = 14
    for N, offsets, runs in CASES:
        exe = os.path.join(REF, f"ref_capture_{N}")
        for off in offsets:
            raw = os.path.join(OUT, f"cap_{N}_{off}.bin")
            subprocess.check_call([exe, os.path.join(REF, "0_0_traj.csv"), os.path.join(REF, "0_0_eepos.traj"), raw,
                                   str(off), "1", "0"], timeout=300)
            a = np.fromfile(raw, np.float32)
            os.remove(raw)
            mat, vec = 3 * n * n * N, n * N
            assert a.size == 2 * mat + vec
            S, P, g = a[:mat], a[mat:2 * mat], a[2 * mat:]
            rec = dict(n=np.int32(n), N=np.int32(N), offset=np.int32(off), rho=np.float32(1e-3), S=S, Pinv=P, gamma=g)
            dS, dP, dg = (torch.from_numpy(x.copy()).cuda() for x in (S, P, g))
            l0 = torch.zeros(vec, device="cuda")
            for j, (tol, cap) in enumerate(runs):
                ref = refgpu.solve(n, N, dS, dP, dg, l0, cap, tol, block=128)
                rec[f"run{j}_tol"] = np.float32(tol)
                rec[f"run{j}_cap"] = np.int32(cap)
                rec[f"run{j}_lam"] = ref["lam"].cpu().numpy()
                rec[f"run{j}_r"] = ref["r"].cpu().numpy()
                rec[f"run{j}_p"] = ref["p"].cpu().numpy()
                rec[f"run{j}_iters"] = np.int32(ref["iters"])
                rec[f"run{j}_flag"] = np.uint8(ref["max_iter_exit"])
                print(f"N={N} offset={off} tol={tol} cap={cap}: reference kernel iters={ref['iters']} "
                      f"max_iter_exit={ref['max_iter_exit']} |lam|max={np.abs(rec[f'run{j}_lam']).max():.4g}", flush=True)
            rec["nruns"] = np.int32(len(runs))
            np.savez_compressed(os.path.join(OUT, f"iiwa_{N}_{off}.npz"), **rec)
    print("golden fixtures:", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
