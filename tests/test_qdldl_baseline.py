"""CPU tests of the baseline wiring: the reference's own QDLDL (oracle/_ref) on our band layout."""
import numpy as np
import pytest

from mpcgpu_b200 import synth
from oracle import qdldl


pytestmark = pytest.mark.skipif(not qdldl.build(), reason="oracle/_ref/libqdldl_ref.so not built and no /root/reference")


def test_nnz_formula():
    # include/qdldl/sqp.cuh:148 ; SURVEY 3.5: 9 436 / 38 332 / 153 916 for N = 32 / 128 / 512 at n = 14
    assert [qdldl.nnz(14, N) for N in (32, 128, 512)] == [9436, 38332, 153916]


def test_pattern_is_upper_triangular_csc():
    n, N = 6, 5
    cp, ri = qdldl.pattern(n, N)
    assert cp[0] == 0 and cp[-1] == qdldl.nnz(n, N)
    for j in range(n * N):
        rows = ri[cp[j]:cp[j + 1]]
        assert np.all(np.diff(rows) == 1) and rows[-1] == j              # contiguous, ends on the diagonal
        b = j // n
        assert rows[0] == max(0, (b - 1) * n)


@pytest.mark.parametrize("n,N", [(2, 3), (6, 12), (14, 32), (14, 128)])
def test_qdldl_solves_the_band_system(oracle_pcg, n, N):
    d = synth.make_systems(n, N, seed=4)
    S, g = d["S"][0], d["gamma"][0]
    # the packed CSC, expanded symmetrically, is the dense matrix the band layout denotes
    cp, ri = qdldl.pattern(n, N)
    val = qdldl.values(S, n, N)[0]
    A = np.zeros((n * N, n * N))
    for j in range(n * N):
        for k in range(cp[j], cp[j + 1]):
            A[ri[k], j] = val[k]
            A[j, ri[k]] = val[k]
    np.testing.assert_array_equal(A, oracle_pcg.band_to_dense(S, n, N))
    x = qdldl.solve(S, g, n, N)
    truth = oracle_pcg.solve_f64(S, g, n, N)
    assert np.abs(x - truth).max() / np.abs(truth).max() < 5e-3         # fp32 LDL^T, cond ~ 1e3-1e4
    assert oracle_pcg.rel_residual(S, g, x, n, N) < 1e-3


def test_batched_threads_agree():
    n, N, B = 6, 12, 16
    d = synth.make_systems(n, N, batch=B, seed=8)
    vals = qdldl.values(d["S"], n, N)
    _, x1 = qdldl.time_batched(vals, d["gamma"], n, N, reps=1, nthreads=1)
    sec, x4 = qdldl.time_batched(vals, d["gamma"], n, N, reps=2, nthreads=4)
    assert sec > 0 and np.array_equal(x1, x4)
    np.testing.assert_array_equal(x1[3], qdldl.solve(d["S"][3], d["gamma"][3], n, N))
