#!/usr/bin/env python
"""BASELINE.json configs[3] on REAL IIWA data: a batch of perturbed 128-knot trajectories assembled by the REFERENCE's own
generate_kkt_submatrices + form_schur_system (oracle/_ref/ref_capture_128, count = B, perturb = 1: SURVEY.md 8d config 4),
solved by our batched kernel; a sample of systems is checked bit-for-bit against the unmodified reference pcg<> kernel.
Test-side tool (uses oracle/); writes gpurun_out/iiwa_batch.json."""
import json
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpcgpu_b200 as mp  # noqa: E402
from oracle import refgpu  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


def main():
    n, N, B = 14, 128, int(os.environ.get("IIWA_BATCH", "256"))
    raw = "/tmp/iiwa_batch.bin"
    subprocess.check_call([os.path.join(REF, "ref_capture_128"), os.path.join(REF, "0_0_traj.csv"), os.path.join(REF, "0_0_eepos.traj"),
                           raw, "0", str(B), "1", "1"], timeout=900, stdout=subprocess.DEVNULL)
    mat, vec = 3 * n * n * N, n * N
    a = np.fromfile(raw, np.float32).reshape(B, 2 * mat + vec)
    os.remove(raw)
    S, P, g = (torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (a[:, :mat], a[:, mat:2 * mat], a[:, 2 * mat:]))
    lam = torch.zeros(B, vec, device="cuda")
    it = torch.zeros(B, dtype=torch.int32, device="cuda")
    fl = torch.zeros(B, dtype=torch.uint8, device="cuda")
    cap, tol = 167, 1e-4

    def run():
        lam.zero_()
        mp.solve_batched(n, N, B, S, P, g, lam, it, fl, cap, tol)

    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    same = 0
    for i in (0, 1, B // 2, B - 1):
        ref = refgpu.solve(n, N, S[i], P[i], g[i], torch.zeros(vec, device="cuda"), cap, tol, block=128)
        ok = ref["iters"] == int(it[i].item()) and bool(fl[i].item()) == ref["max_iter_exit"] and torch.equal(ref["lam"], lam[i])
        same += int(ok)
    out = dict(n=n, N=N, batch=B, ms_per_batch=ms, traj_per_sec=B / (ms * 1e-3), mean_iters=float(it.float().mean().item()),
               max_iter_exit_frac=float(fl.float().mean().item()), iters_min=int(it.min().item()), iters_max=int(it.max().item()),
               bit_identical_to_reference_kernel=f"{same}/4 sampled systems",
               source="reference generate_kkt_submatrices + form_schur_system on examples/trajfiles/0_0 + N(0,0.05^2) q, N(0,0.01^2) qd, N(0,1) u")
    print(out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "iiwa_batch.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
