#!/usr/bin/env python
"""Stress of the assembly kernels (rows f1 / f2): random dense and DIAGONAL-cost KKT systems (exact zeros everywhere: the
case the zero-numerator division path of gbd_schur.cuh exists for), both team mappings, against the reference's own kernels
(oracle/_ref/libref_schur.so) -- compared as BIT PATTERNS (uint32 views: +0 and -0 are different here, stricter than the
tests) and as values.  Test-side tool (uses oracle/).  Writes gpurun_out/stress_schur.log."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpcgpu_b200 as mp  # noqa: E402
from mpcgpu_b200 import _capi  # noqa: E402
from oracle import refgpu, schur  # noqa: E402


def diagonalise(G, n, m, N):
    """Keep only the diagonals of Q_k / R_k (the IIWA cost structure), scaled to a spread of magnitudes."""
    G = G.copy()
    nn, mm = n * n, m * m
    rng = np.random.default_rng(int(G.size))
    for k in range(N):
        off = k * (nn + mm)
        Q = G[off:off + nn].reshape(n, n)
        G[off:off + nn] = np.diag(np.diag(Q) * 10.0 ** rng.integers(-3, 3, n)).ravel()
        if k < N - 1:
            R = G[off + nn:off + nn + mm].reshape(m, m)
            G[off + nn:off + nn + mm] = np.diag(np.diag(R) * 10.0 ** rng.integers(-3, 3, m)).ravel()
    return G.astype(np.float32)


def run(n, m, N, G, C, g, c, team):
    prev = _capi.lib().gbd_schur_set_team(team)
    try:
        dG, dC, dg, dc = (torch.from_numpy(x.copy()).cuda() for x in (G, C, g, c))
        dS = torch.zeros(3 * n * n * N, device="cuda")
        dP = torch.zeros(3 * n * n * N, device="cuda")
        dgam = torch.zeros(n * N, device="cuda")
        mp.form_schur_system(n, m, N, dG, dC, dg, dc, dS, dP, dgam, 1e-3)
        torch.cuda.synchronize()
        return [x.cpu().numpy() for x in (dS, dP, dgam, dG)]
    finally:
        _capi.lib().gbd_schur_set_team(prev)


def ref(n, m, N, G, C, g, c):
    rG, rC, rg, rc = (torch.from_numpy(x.copy()).cuda() for x in (G, C, g, c))
    rS = torch.zeros(3 * n * n * N, device="cuda")
    rP = torch.zeros(3 * n * n * N, device="cuda")
    rgam = torch.zeros(n * N, device="cuda")
    refgpu.form_schur_system(n, m, N, rG, rC, rg, rc, rS, rP, rgam, 1e-3)
    torch.cuda.synchronize()
    return [x.cpu().numpy() for x in (rS, rP, rgam, rG)]


def mask(x, n, N):
    x = x.reshape(N, 3, n, n).copy()
    x[0, 0] = 0
    x[N - 1, 2] = 0
    return x.ravel()


def main():
    have_ref = refgpu.schur_available()
    cases = value_bad = bits_bad = 0
    for (n, m) in ((14, 7), (6, 3)):
        for N in ((8, 32, 128) if n == 14 else (12,)):
            for seed in range(int(os.environ.get("SEEDS", "12"))):
                G, C, g, c = schur.make_kkt(n, m, N, seed=1000 + seed)
                for kind in ("dense", "diagonal"):
                    Gk = G if kind == "dense" else diagonalise(G, n, m, N)
                    outs = {t: run(n, m, N, Gk, C, g, c, t) for t in (0, 1)}
                    if have_ref:
                        outs["ref"] = ref(n, m, N, Gk, C, g, c)
                    base = outs["ref"] if have_ref else outs[0]
                    for t in (0, 1):
                        cases += 1
                        for i, (a, b) in enumerate(zip(outs[t], base)):
                            if i < 2:
                                a, b = mask(a, n, N), mask(b, n, N)
                            if not np.array_equal(a, b):
                                value_bad += 1
                                print("VALUE mismatch", n, m, N, seed, kind, "team", t, "array", i)
                            elif not np.array_equal(a.view(np.uint32), b.view(np.uint32)):
                                bits_bad += 1
                                print("zero-sign mismatch", n, m, N, seed, kind, "team", t, "array", i,
                                      int((a.view(np.uint32) != b.view(np.uint32)).sum()), "elements")
    msg = (f"stress_schur: {cases} (system, team) cases against {'the reference kernels' if have_ref else 'team 0 (no oracle/_ref)'}: "
           f"{value_bad} arrays differ in value, {bits_bad} arrays equal in value but not in bit pattern (signed zeros)")
    print(msg)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "stress_schur.log"), "w").write(msg + "\n")
    return 1 if value_bad else 0


if __name__ == "__main__":
    sys.exit(main())
