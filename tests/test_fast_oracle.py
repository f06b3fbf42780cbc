"""CPU tests of oracle/pcg_fast_oracle.c, the restatement of the tolerance-parity kernels' operation order: on the
reference's own IIWA systems it must agree with the reference kernel's stored answers (tests/golden, minted on a B200 by
tools/make_golden.py) within the stated tolerance of SURVEY.md 8(c)(ii) for every cluster shape the kernels use."""
import os

import numpy as np
import pytest

from test_golden import GOLDEN, load


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_fast_order_within_tolerance_of_reference_kernel(oracle_pcg, path):
    g = load(path)
    n, N = g["n"], g["N"]
    for run in g["runs"]:
        r_ref = oracle_pcg.rel_residual(g["S"], g["gamma"], run["lam"], n, N)
        # (cluster size, lanes per knot row): the single-solve shapes, then the packed batch shapes (32 knot rows per CTA)
        for C, lanes in ((4, 16), (8, 16), (16, 16), (N // 32, n)):
            if C < 1 or N // C < 2:
                continue
            f = oracle_pcg.pcg_fast(g["S"], g["Pinv"], g["gamma"], np.zeros(n * N, np.float32), n, N, C, run["cap"], run["tol"], lanes=lanes)
            assert abs(f["iters"] - run["iters"]) <= 2
            if run["iters"] < run["cap"] - 2 or run["max_iter_exit"]:
                assert f["max_iter_exit"] == run["max_iter_exit"]
            assert np.abs(f["lam"].astype(np.float64) - run["lam"]).max() / np.abs(run["lam"]).max() <= 1e-3
            assert oracle_pcg.rel_residual(g["S"], g["gamma"], f["lam"], n, N) <= 1.1 * r_ref + 1e-6


def test_fast_order_solves_a_small_system_to_fp64_truth(oracle_pcg):
    """A well-conditioned synthetic system (n = 6, N = 12) against the fp64 block-Thomas solution."""
    from mpcgpu_b200 import synth
    d = synth.make_systems(6, 12, seed=3)
    S, P, g = (d[k][0] for k in ("S", "Pinv", "gamma"))
    for C in (2, 3):
        f = oracle_pcg.pcg_fast(S, P, g, np.zeros(72, np.float32), 6, 12, C, 100, 1e-12)
        assert not f["max_iter_exit"]
        x = oracle_pcg.solve_f64(S, g, 6, 12)
        assert np.abs(f["lam"] - x).max() / np.abs(x).max() < 1e-3


def test_fast_order_exit_semantics_match_reference_order(oracle_pcg):
    from mpcgpu_b200 import synth
    n, N = 14, 32
    d = synth.make_systems(n, N, seed=21)
    S, P, g, l0 = (d[k][0] for k in ("S", "Pinv", "gamma", "lambda0"))
    for cap, tol in ((3, 1e-30), (0, 1e-6), (50, 1e30), (1, 1e-30)):
        a = oracle_pcg.pcg_fast(S, P, g, l0, n, N, 4, cap, tol)
        b = oracle_pcg.pcg(S, P, g, l0, n, N, cap, tol)
        assert (a["iters"], a["max_iter_exit"]) == (b["iters"], b["max_iter_exit"])
    a = oracle_pcg.pcg_fast(S, P, g, l0, n, N, 4, 0, 1e-6)
    assert np.array_equal(a["lam"], l0)
