"""CPU tests: the oracle against the reference's known answers and fp64 ground truth."""
import numpy as np
import pytest

from mpcgpu_b200 import synth


def test_glass_known_answers(oracle_pcg):
    # GLASS/GTests/test.cu:157-171 (DotProduct, a=i, b=2i, n=100 -> 656700) and :275-282 (reduce 0..99 -> 4950)
    a = np.arange(100, dtype=np.float32)
    assert oracle_pcg.glass_dot(a, 2 * a) == np.float32(656700.0)
    assert oracle_pcg.glass_reduce(a) == np.float32(4950.0)
    assert oracle_pcg.glass_reduce(a.astype(np.float64)) == 4950.0


@pytest.mark.parametrize("cnt", [1, 2, 3, 4, 5, 7, 14, 31, 32, 33, 64, 100, 128, 512])
def test_glass_tree_order(oracle_pcg, cnt):
    """The tree is a fixed association of the same terms: exact on integers, and equal to a direct
    Python restatement of reduce.cuh's loop on random floats."""
    rng = np.random.default_rng(cnt)
    x = rng.standard_normal(cnt).astype(np.float32)
    v = x.copy()
    s = cnt
    while s > 3:
        odd = s & 1
        s = (s - odd) // 2
        v[:s] = v[:s] + v[s:2 * s]
        if odd:
            v[0] = np.float32(v[0] + v[2 * s])
    for i in range(1, s):
        v[0] = np.float32(v[0] + v[i])
    assert oracle_pcg.glass_reduce(x) == v[0]
    ints = np.arange(cnt, dtype=np.float32)
    assert oracle_pcg.glass_reduce(ints) == np.float32(cnt * (cnt - 1) // 2)


def test_gemv_known_answer_through_bdmv(oracle_pcg):
    # GLASS/GTests/test.cu:46-68,327-344: a[i] = i (5x7), x[j] = 2j -> {910,952,994,1036,1078} reading a
    # column-major, {182,476,770,1064,1358} reading it row-major ("transposed").  bdmv is the same
    # column-major MAC, so embed the 5x7 in the diagonal 7x7 tile of block row 0.
    n, N = 7, 2
    a = np.arange(35, dtype=np.float32)
    x = np.zeros(n * N, np.float32)
    x[:7] = 2 * np.arange(7)
    for block, want in ((a.reshape(7, 5).T, [910, 952, 994, 1036, 1078]),      # block[r,c] = a[r + 5c]
                        (a.reshape(5, 7), [182, 476, 770, 1064, 1358])):       # block[r,c] = a[7r + c]
        S = np.zeros((N, 3, n, n), np.float32)                                  # [b][t][c][r]
        S[0, 1, :, :5] = block.T
        y = oracle_pcg.bdmv(S, x, n, N)
        assert y[:5].tolist() == want


def test_bdmv_matches_dense(oracle_pcg):
    for n, N in [(2, 3), (6, 12), (14, 32)]:
        d = synth.make_systems(n, N, seed=n * N)
        S = d["S"][0]
        x = np.random.default_rng(1).standard_normal(n * N).astype(np.float32)
        y = oracle_pcg.bdmv(S, x, n, N)
        A = oracle_pcg.band_to_dense(S, n, N)
        np.testing.assert_allclose(y, A @ x.astype(np.float64), rtol=2e-5, atol=2e-5)


def test_bdmv_ignores_pad_tiles(oracle_pcg):
    n, N = 6, 12
    a = synth.make_systems(n, N, seed=3)
    b = synth.make_systems(n, N, seed=3, nan_pads=True)
    x = np.random.default_rng(2).standard_normal(n * N).astype(np.float32)
    assert np.array_equal(oracle_pcg.bdmv(a["S"][0], x, n, N), oracle_pcg.bdmv(b["S"][0], x, n, N))


def test_demo_system_G1(oracle_pcg, g1):
    """GBD-PCG/examples/pcg_solve.cu system: PCG with the stair preconditioner reaches the fp64 solution."""
    n, N = g1["n"], g1["N"]
    from conftest import G1_S, G1_GAMMA
    x = oracle_pcg.solve_f64(np.array(G1_S), np.array(G1_GAMMA), n, N)            # decimal literals in fp64
    np.testing.assert_allclose(x, g1["lam"], rtol=1e-7)
    x32 = oracle_pcg.solve_f64(g1["S"], g1["gamma"], n, N)                        # after the fp32 cast (cond ~1.6e3)
    np.testing.assert_allclose(x32, g1["lam"], rtol=1e-4)
    A = oracle_pcg.band_to_dense(g1["S"], n, N)
    ev = np.linalg.eigvalsh(A)
    assert ev.max() < 0 and abs(ev.min() + 5.02) < 0.01            # symmetric negative definite, SURVEY 8c
    r = oracle_pcg.pcg(g1["S"], g1["Pinv"], g1["gamma"], np.zeros(6, np.float32), n, N, 100, 1e-10)
    assert not r["max_iter_exit"] and r["iters"] <= 12
    np.testing.assert_allclose(r["lam"], g1["lam"], rtol=2e-3)
    r64 = oracle_pcg.pcg(g1["S"].astype(np.float64), g1["Pinv"].astype(np.float64), g1["gamma"].astype(np.float64),
                         np.zeros(6), n, N, 100, 1e-20)
    np.testing.assert_allclose(r64["lam"], x32, rtol=1e-8)


@pytest.mark.parametrize("n,N,tol,cap", [(6, 12, 1e-6, 60), (14, 32, 1e-6, 173), (14, 128, 1e-4, 167)])
def test_oracle_converges_to_truth(oracle_pcg, n, N, tol, cap):
    d = synth.make_systems(n, N, seed=11)
    S, P, g, l0 = d["S"][0], d["Pinv"][0], d["gamma"][0], d["lambda0"][0]
    r = oracle_pcg.pcg(S, P, g, l0, n, N, cap, tol)
    assert not r["max_iter_exit"] and 1 <= r["iters"] <= cap
    assert abs(r["eta"]) < tol
    x = oracle_pcg.solve_f64(S, g, n, N)
    assert oracle_pcg.rel_residual(S, g, r["lam"], n, N) < 2e-3
    assert np.abs(r["lam"] - x).max() / np.abs(x).max() < 5e-3
    # r_out really is gamma - S*lambda up to fp32 drift
    rr = g.astype(np.float64) - oracle_pcg.bdmv(S.astype(np.float64), r["lam"].astype(np.float64), n, N)
    assert np.abs(rr - r["r"]).max() < 1e-3


def test_oracle_exit_semantics(oracle_pcg):
    n, N = 6, 12
    d = synth.make_systems(n, N, seed=5)
    S, P, g, l0 = d["S"][0], d["Pinv"][0], d["gamma"][0], d["lambda0"][0]
    # cap hit: iters == max_iter, flag set (pcg.cuh:154,212)
    r = oracle_pcg.pcg(S, P, g, l0, n, N, 3, 1e-30)
    assert r["iters"] == 3 and r["max_iter_exit"]
    # max_iter = 0: lambda untouched
    r0 = oracle_pcg.pcg(S, P, g, l0 + 1, n, N, 0, 1e-6)
    assert r0["iters"] == 0 and r0["max_iter_exit"] and np.array_equal(r0["lam"], l0 + 1)
    # huge tolerance: exits after exactly one iteration (no pre-loop test)
    r1 = oracle_pcg.pcg(S, P, g, l0, n, N, 50, 1e30)
    assert r1["iters"] == 1 and not r1["max_iter_exit"]
    # warm start from the converged solution takes fewer iterations
    full = oracle_pcg.pcg(S, P, g, l0, n, N, 200, 1e-8)
    warm = oracle_pcg.pcg(S, P, g, full["lam"], n, N, 200, 1e-8)
    assert warm["iters"] < full["iters"]
    # contraction on/off is the same algorithm (iteration counts within 2)
    nc = oracle_pcg.pcg(S, P, g, l0, n, N, 200, 1e-8, contract=False)
    assert abs(nc["iters"] - full["iters"]) <= 2


def test_oracle_batched_equals_single(oracle_pcg):
    n, N, B = 6, 12, 5
    d = synth.make_systems(n, N, batch=B, seed=9)
    rb = oracle_pcg.pcg_batched(d["S"], d["Pinv"], d["gamma"], d["lambda0"], n, N, B, 60, 1e-6)
    for i in range(B):
        r = oracle_pcg.pcg(d["S"][i], d["Pinv"][i], d["gamma"][i], d["lambda0"][i], n, N, 60, 1e-6)
        assert np.array_equal(r["lam"], rb["lam"][i]) and r["iters"] == rb["iters"][i]


def test_stair_preconditioner_is_D_minus_DOD(oracle_pcg):
    n, N = 6, 12
    d = synth.make_systems(n, N, seed=2, dtype=np.float64)
    A = oracle_pcg.band_to_dense(d["S"][0], n, N)
    Pm = oracle_pcg.band_to_dense(d["Pinv"][0], n, N)
    D = np.zeros_like(A)
    for b in range(N):
        D[b * n:(b + 1) * n, b * n:(b + 1) * n] = A[b * n:(b + 1) * n, b * n:(b + 1) * n]
    Di = np.linalg.inv(D)
    np.testing.assert_allclose(Pm, Di - Di @ (A - D) @ Di, atol=1e-10)
