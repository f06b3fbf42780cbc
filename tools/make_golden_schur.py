#!/usr/bin/env python
"""Mints tests/golden/schur_iiwa_*.npz ON THE B200 BOX (gpurun): golden vectors for rows f1 / f2.

Everything on this path is the REFERENCE: oracle/_ref/ref_capture_32 (mode 8|16) runs the reference's own
generate_kkt_submatrices on examples/trajfiles/0_0_* (the inputs G, C, g, c), its form_schur_system at rho = 1e-3
(outputs S, Pinv, gamma and G overwritten with the block inverses), its pcg<> (lambda) and its compute_dz (dz).
Nothing of this repo's product is involved.  Written to gpurun_out/golden/; copy into tests/golden/ and commit."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(ROOT, "gpurun_out", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    n, m, N = 14, 7, 32
    exe = os.path.join(REF, f"ref_capture_{N}")
    for off in (0, 200):
        raw = os.path.join(OUT, f"cap_{N}_{off}.bin")
        subprocess.check_call([exe, os.path.join(REF, "0_0_traj.csv"), os.path.join(REF, "0_0_eepos.traj"), raw, str(off), "1", "0",
                               str(8 | 16)], timeout=300)
        a = np.fromfile(raw, np.float32)
        mat, vec = 3 * n * n * N, n * N
        S, P, gam, lam = a[:mat], a[mat:2 * mat], a[2 * mat:2 * mat + vec], a[2 * mat + vec:2 * mat + 2 * vec]
        tail = a[2 * mat + 2 * vec:].view(np.uint32)
        k = np.fromfile(raw + ".kkt", np.float32)
        nG, nC, ng, nc = (n * n + m * m) * (N - 1) + n * n, (n * n + n * m) * (N - 1), (n + m) * (N - 1) + n, n * N
        assert k.size == 2 * nG + nC + 2 * ng + nc, (k.size, nG, nC, ng, nc)
        o = 0
        parts = {}
        for name, cnt in (("G", nG), ("C", nC), ("g", ng), ("c", nc), ("Ginv", nG), ("dz", ng)):
            parts[name] = k[o:o + cnt].copy()
            o += cnt
        np.savez_compressed(os.path.join(OUT, f"schur_iiwa_{N}_{off}.npz"), n=np.int32(n), m=np.int32(m), N=np.int32(N),
                            offset=np.int32(off), rho=np.float32(1e-3), S=S, Pinv=P, gamma=gam, lam=lam, pcg_iters=np.int32(tail[0]),
                            **parts)
        os.remove(raw)
        os.remove(raw + ".kkt")
        print(f"offset {off}: pcg iters {tail[0]} |S|max {np.nanmax(np.abs(S)):.4g} |dz|max {np.abs(parts['dz']).max():.4g}", flush=True)
    print("golden fixtures:", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
