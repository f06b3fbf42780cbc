"""CPU tests of the Schur-assembly / dz oracle (oracle/schur_oracle.c), rows f1 and f2 of SURVEY.md 8f.

The oracle restates the reference's operation order (fp32, explicit FMAs); here it is checked against an independent
fp64 statement of the same mathematics, against the synthetic-system generator's stair preconditioner, and for the
properties the downstream solver relies on (symmetry of the assembled system, S * S^-1-like structure of Pinv)."""
import numpy as np
import pytest

from oracle import schur


@pytest.mark.parametrize("n,m,N", [(14, 7, 8), (14, 7, 32), (6, 3, 5), (4, 2, 4), (2, 1, 3)])
def test_oracle_matches_fp64_mathematics(n, m, N):
    G, C, g, c = schur.make_kkt(n, m, N, seed=n + N)
    o = schur.form(G, C, g, c, n, m, N, 1e-3)
    t = schur.form_f64(G, C, g, c, n, m, N, 1e-3)
    for k in ("S", "Pinv", "gamma"):
        a, b = o[k].astype(np.float64), t[k]
        assert (np.isnan(a) == np.isnan(b)).all(), f"{k}: pad tiles differ"     # only the two pad tiles stay untouched
        mask = ~np.isnan(b)
        assert np.abs(a[mask] - b[mask]).max() <= 2e-5 * np.abs(b[mask]).max(), k


def test_inverses_left_in_G():
    n, m, N = 14, 7, 6
    G, C, g, c = schur.make_kkt(n, m, N, seed=3)
    o = schur.form(G, C, g, c, n, m, N, 1e-3)
    nn, mm = n * n, m * m
    for k in range(N):
        Q = G[k * (nn + mm):k * (nn + mm) + nn].reshape(n, n).T.astype(np.float64) + 1e-3 * np.eye(n)
        Qi = o["Ginv"][k * (nn + mm):k * (nn + mm) + nn].reshape(n, n).T
        assert np.abs(Q @ Qi - np.eye(n)).max() < 1e-4
        if k < N - 1:
            R = G[k * (nn + mm) + nn:(k + 1) * (nn + mm)].reshape(m, m).T.astype(np.float64) + 1e-3 * np.eye(m)
            Ri = o["Ginv"][k * (nn + mm) + nn:(k + 1) * (nn + mm)].reshape(m, m).T
            assert np.abs(R @ Ri - np.eye(m)).max() < 1e-4


def test_assembled_system_feeds_the_pcg_oracle():
    """S, Pinv, gamma from the assembly oracle are a valid pcg<> input: symmetric S, and PCG converges on it."""
    from oracle import pcg as opcg
    n, m, N = 14, 7, 16
    G, C, g, c = schur.make_kkt(n, m, N, seed=11)
    o = schur.form(G, C, g, c, n, m, N, 1e-3, pad=0.0)
    S = o["S"].reshape(N, 3, n, n)
    for b in range(N - 1):
        np.testing.assert_allclose(S[b, 2], S[b + 1, 0].transpose(1, 0), rtol=0, atol=0)     # right tile = (left tile of next)^T
    r = opcg.pcg(o["S"], o["Pinv"], o["gamma"], np.zeros(n * N, np.float32), n, N, 200, 1e-8)
    assert not r["max_iter_exit"]
    assert opcg.rel_residual(o["S"], o["gamma"], r["lam"], n, N) < 1e-3


def test_dz_matches_fp64():
    n, m, N = 14, 7, 9
    G, C, g, c = schur.make_kkt(n, m, N, seed=5)
    o = schur.form(G, C, g, c, n, m, N, 1e-3)
    lam = np.random.default_rng(0).standard_normal(n * N).astype(np.float32)
    dz = schur.dz(o["Ginv"], C, g, lam, n, m, N)
    nn, mm, nm = n * n, m * m, n * m
    for k in range(N):
        Qi = o["Ginv"][k * (nn + mm):k * (nn + mm) + nn].reshape(n, n).T.astype(np.float64)
        rhs = g[k * (n + m):k * (n + m) + n].astype(np.float64) - lam[k * n:(k + 1) * n]
        if k < N - 1:
            A = C[k * (nn + nm):k * (nn + nm) + nn].reshape(n, n).T.astype(np.float64)
            B = C[k * (nn + nm) + nn:(k + 1) * (nn + nm)].reshape(m, n).T.astype(np.float64)
            rhs = rhs - A.T @ lam[(k + 1) * n:(k + 2) * n]
            Ri = o["Ginv"][k * (nn + mm) + nn:(k + 1) * (nn + mm)].reshape(m, m).T.astype(np.float64)
            du = Ri @ (g[k * (n + m) + n:(k + 1) * (n + m)].astype(np.float64) - B.T @ lam[(k + 1) * n:(k + 2) * n])
            np.testing.assert_allclose(dz[k * (n + m) + n:(k + 1) * (n + m)], du, rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(dz[k * (n + m):k * (n + m) + n], Qi @ rhs, rtol=2e-4, atol=2e-5)


# ---- pinned against the REFERENCE: tests/golden/schur_iiwa_*.npz were minted on a B200 by tools/make_golden_schur.py from the
# reference's own generate_kkt_submatrices -> form_schur_system -> pcg<> -> compute_dz on examples/trajfiles/0_0_*.
import glob
import os

GOLDEN_SCHUR = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "schur_iiwa_*.npz")))


def _mask_pads(x, n, N):
    x = np.array(x, np.float32).reshape(N, 3, n, n).copy()
    x[0, 0] = 0
    x[N - 1, 2] = 0
    return x


@pytest.mark.parametrize("path", GOLDEN_SCHUR, ids=[os.path.basename(p) for p in GOLDEN_SCHUR])
def test_oracle_bit_exact_on_reference_minted_vectors(path):
    z = np.load(path)
    n, m, N, rho = int(z["n"]), int(z["m"]), int(z["N"]), float(z["rho"])
    o = schur.form(z["G"], z["C"], z["g"], z["c"], n, m, N, rho)
    assert np.array_equal(o["gamma"], z["gamma"])
    assert np.array_equal(o["Ginv"], z["Ginv"])
    for k in ("S", "Pinv"):
        assert np.array_equal(_mask_pads(o[k], n, N), _mask_pads(z[k], n, N)), k
    assert np.array_equal(schur.dz(z["Ginv"], z["C"], z["g"], z["lam"], n, m, N), z["dz"])


def test_golden_schur_fixtures_present():
    assert len(GOLDEN_SCHUR) >= 2
