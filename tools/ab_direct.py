#!/usr/bin/env python
"""Direct solver (block cyclic reduction, gbd_bcr_*) next to our PCG kernel and the reference's CPU QDLDL on the same
systems: time per solve and accuracy against the fp64 solution.  Test-side tool (uses oracle/); writes gpurun_out/ab_direct.json."""
import glob
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpcgpu_b200 as mp  # noqa: E402
from mpcgpu_b200 import _capi, synth  # noqa: E402
from oracle import pcg as op  # noqa: E402
from oracle import qdldl  # noqa: E402

CAPS = {32: 173, 64: 167, 128: 167, 256: 118, 512: 67}


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


def one(name, n, N, S, P, g, out):
    L = _capi.lib()
    st = torch.cuda.current_stream().cuda_stream
    S0 = np.nan_to_num(S).astype(np.float32)
    dS, dP, dg = (torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (S0, np.nan_to_num(P).astype(np.float32), g))
    lam = torch.zeros(n * N, device="cuda")
    it = torch.zeros(1, dtype=torch.int32, device="cuda")
    fl = torch.zeros(1, dtype=torch.uint8, device="cuda")
    truth = op.solve_f64(S0, g, n, N)
    t_direct = timed(lambda: L.gbd_bcr_solve_f32(n, N, dS.data_ptr(), dg.data_ptr(), lam.data_ptr(), st))
    torch.cuda.synchronize()
    ld = lam.cpu().numpy().copy()

    def pcg():
        lam.zero_()
        L.gbd_pcg_solve_f32(n, N, dS.data_ptr(), dP.data_ptr(), dg.data_ptr(), lam.data_ptr(), 0, 0, 0, 0, it.data_ptr(), fl.data_ptr(),
                            CAPS[N], 1e-4, st)
    t_zero = timed(lambda: lam.zero_())
    t_pcg = timed(pcg) - t_zero
    torch.cuda.synchronize()
    lp = lam.cpu().numpy().copy()
    rec = dict(system=name, n=n, N=N, direct_us=t_direct, pcg_us=t_pcg, pcg_iters=int(it.item()), pcg_hit_cap=bool(fl.item()),
               direct_rel_err=float(np.abs(ld - truth).max() / np.abs(truth).max()), direct_rel_res=op.rel_residual(S0, g, ld, n, N),
               pcg_rel_err=float(np.abs(lp - truth).max() / np.abs(truth).max()), pcg_rel_res=op.rel_residual(S0, g, lp, n, N))
    if qdldl.available():
        vals = qdldl.values(S0[None], n, N)
        sec, x = qdldl.time_batched(vals, g[None], n, N, reps=200, nthreads=1)
        rec.update(qdldl_cpu_us=1e6 * sec / 200, qdldl_rel_err=float(np.abs(x[0] - truth).max() / np.abs(truth).max()))
    out.append(rec)
    print({k: (round(v, 6) if isinstance(v, float) else v) for k, v in rec.items()}, flush=True)


def main():
    out = []
    for N in (32, 128, 512):
        d = synth.make_systems(14, N, batch=1, seed=3)
        one(f"synthetic N={N}", 14, N, d["S"][0], d["Pinv"][0], d["gamma"][0], out)
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "iiwa_*_0.npz"))):
        z = np.load(path)
        one(os.path.basename(path), int(z["n"]), int(z["N"]), z["S"], z["Pinv"], z["gamma"], out)
    # batched: 1024 systems
    n, N, B = 14, 128, 1024
    d = synth.make_systems(n, N, batch=B, seed=1234)
    dS, dg = torch.from_numpy(d["S"]).cuda(), torch.from_numpy(d["gamma"]).cuda()
    lam = torch.zeros(B, n * N, device="cuda")
    ms = timed(lambda: mp.solve_direct(n, N, dS.reshape(-1), dg.reshape(-1), lam.reshape(-1), batch=B), reps=5, warm=2) / 1e3
    out.append(dict(system="synthetic batch", n=n, N=N, batch=B, direct_ms=ms, traj_per_sec=B / (ms * 1e-3)))
    print(out[-1])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ab_direct.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
