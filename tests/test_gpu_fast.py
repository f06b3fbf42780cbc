"""GPU tests of the tolerance-parity ("fast") kernels -- the library's default numerics (include/gbd_pcg.h).

Two bars, both through the C ABI:

1. BIT-EXACT against oracle/pcg_fast_oracle.c, the CPU restatement of exactly the operation order these kernels use
   (Chronopoulos-Gear recurrence, per-tile chains, per-CTA reductions): lambda, r, p, iteration count, exit flag.
   A tolerance can hide a bug; this bar cannot.
2. TOLERANCE PARITY against the reference kernel pcg<T,n,N> (GBD-PCG/include/pcg.cuh:54-218) -- its outputs stored in
   tests/golden/iiwa_*.npz (minted on a B200 from the reference's own assembly + kernel, tools/make_golden.py), the
   kernel itself run on this GPU when oracle/_ref is present, and the bit-faithful C oracle elsewhere -- with the
   policy of SURVEY.md 8(c)(ii), written here:
       iteration count within +-2 of the reference's;
       max_iter_exit identical unless the reference is within 2 iterations of the cap;
       max|lambda - lambda_ref| / max|lambda_ref| <= 1e-3;
       fp64 relative residual ||gamma - S lambda|| / ||gamma|| <= 1.1 x the reference kernel's (+1e-6 absolute).
"""
import numpy as np
import pytest

from mpcgpu_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.fast_numerics]

ITER_TOL = 2
LAMBDA_TOL = 1e-3
RESID_FACTOR = 1.1


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    torch.cuda.init()
    return torch


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _gpu_solve(torch, m, S, P, g, l0, n, N, max_iter, tol):
    dS, dP, dg = (_dev(torch, np.asarray(x, np.float32)) for x in (S, P, g))
    lam = _dev(torch, np.asarray(l0, np.float32))
    r = torch.full((n * N,), float("nan"), device="cuda")
    p = torch.full((n * N,), float("nan"), device="cuda")
    it = torch.full((1,), -1, dtype=torch.int32, device="cuda")
    fl = torch.full((1,), 7, dtype=torch.uint8, device="cuda")
    m.pcg_launch(n, N, dS, dP, dg, lam, r, p, None, None, it, fl, max_iter, tol)
    torch.cuda.synchronize()
    return dict(lam=lam.cpu().numpy(), iters=int(it.item()), max_iter_exit=bool(fl.item()), r=r.cpu().numpy(), p=p.cpu().numpy())


def _assert_same(a, b, what=""):
    assert a["iters"] == b["iters"], f"{what}: iters {a['iters']} vs {b['iters']}"
    assert a["max_iter_exit"] == b["max_iter_exit"], what
    for k in ("lam", "r", "p"):
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), f"{what}: {k} differs, max abs " \
            f"{np.abs(np.asarray(a[k], np.float64) - np.asarray(b[k], np.float64)).max()}"


def _assert_parity(oracle_pcg, got, ref, S, g, n, N, cap, what=""):
    """SURVEY.md 8(c)(ii) against a reference-order result `ref` (dict with lam, iters, max_iter_exit)."""
    assert abs(got["iters"] - ref["iters"]) <= ITER_TOL, f"{what}: iters {got['iters']} vs reference {ref['iters']}"
    if ref["iters"] < cap - ITER_TOL or ref["max_iter_exit"]:
        if not (ref["max_iter_exit"] is False and ref["iters"] >= cap - ITER_TOL):
            assert got["max_iter_exit"] == ref["max_iter_exit"], what
    scale = np.abs(ref["lam"]).max()
    if scale > 0:
        err = np.abs(got["lam"].astype(np.float64) - ref["lam"].astype(np.float64)).max() / scale
        assert err <= LAMBDA_TOL, f"{what}: lambda differs from the reference by {err:.2e} (relative, max norm)"
    r_got = oracle_pcg.rel_residual(S, g, got["lam"], n, N)
    r_ref = oracle_pcg.rel_residual(S, g, ref["lam"], n, N)
    assert r_got <= RESID_FACTOR * r_ref + 1e-6, f"{what}: residual {r_got:.3e} vs reference {r_ref:.3e}"


def test_default_numerics_is_fast_and_resolves_to_a_fast_kernel(torch_cuda, capi):
    assert capi.lib().gbd_pcg_get_numerics() == capi.NUMERICS_FAST
    for N in (32, 64, 128, 256):
        v = capi.resolved_variant(14, N)
        assert v["fast"] and v["kernel"].endswith("fast"), v
    # shapes without a fast kernel are served by the bit-exact family
    assert not capi.resolved_variant(14, 128, f64=True)["fast"]
    prev = capi.set_numerics(capi.NUMERICS_BITEXACT)
    try:
        assert not capi.resolved_variant(14, 128)["fast"]
    finally:
        capi.set_numerics(prev)


def test_every_fast_variant_bit_exact_vs_its_own_oracle(torch_cuda, capi, oracle_pcg):
    """Each compiled fast variant against the CPU restatement of its operation order; NaN pad tiles; warm start."""
    import mpcgpu_b200 as m
    L = capi.lib()
    seen = 0
    for v in capi.variants():
        if not v["fast"] or v["mode"] not in (20, 21, 24, 25, 26, 27, 28, 31):
            continue
        n, N, C = v["n"], v["N"], v["cluster"]
        lanes = {24: 1, 25: 1, 26: n, 27: 0, 28: 0, 31: 0}.get(v["mode"], 16)  # the oracle's reduction order: 16 / n lanes per knot row, 0 = batch kernel, 1 = grid kernel
        d = synth.make_systems(n, N, batch=2, seed=300 + n + N, nan_pads=True)
        cap, tol = (40, 1e-7) if N > 128 else (60, 1e-6)
        if n == 64 and N == 256:
            cap = 12                                           # the CPU restatement of a 25 MB system is slow
        assert L.gbd_pcg_set_tuning(n, N, 0, C, v["mode"]) == 0
        try:
            for i in range(2):
                l0 = d["lambda0"][i] if i == 0 else (0.1 * np.random.default_rng(i).standard_normal(n * N)).astype(np.float32)
                got = _gpu_solve(torch_cuda, m, d["S"][i], d["Pinv"][i], d["gamma"][i], l0, n, N, cap, tol)
                want = oracle_pcg.pcg_fast(d["S"][i], d["Pinv"][i], d["gamma"][i], l0, n, N, C, cap, tol, lanes=lanes)
                _assert_same(got, want, f"variant {v} system {i}")
                assert np.isfinite(got["lam"]).all()
        finally:
            L.gbd_pcg_set_tuning(n, N, 0, 0, -1)
        seen += 1
    assert seen >= 20


def test_golden_iiwa_systems_tolerance_parity_vs_reference_kernel(torch_cuda, capi, oracle_pcg):
    """The reference's own IIWA systems and the reference kernel's own answers (tests/golden): every fast variant of the
    size within the stated tolerance, and bit-exact against its own oracle on the real data too."""
    import mpcgpu_b200 as m
    from test_golden import GOLDEN, load
    L = capi.lib()
    checked = 0
    for path in GOLDEN:
        g = load(path)
        n, N = g["n"], g["N"]
        l0 = np.zeros(n * N, np.float32)
        for v in [v for v in capi.variants() if v["n"] == n and v["N"] == N and v["mode"] in (20, 21)]:
            assert L.gbd_pcg_set_tuning(n, N, 0, v["cluster"], v["mode"]) == 0
            try:
                for run in g["runs"]:
                    got = _gpu_solve(torch_cuda, m, g["S"], g["Pinv"], g["gamma"], l0, n, N, run["cap"], run["tol"])
                    _assert_parity(oracle_pcg, got, run, g["S"], g["gamma"], n, N, run["cap"], f"{g['name']} tol {run['tol']} {v}")
                    want = oracle_pcg.pcg_fast(g["S"], g["Pinv"], g["gamma"], l0, n, N, v["cluster"], run["cap"], run["tol"])
                    _assert_same(got, want, f"{g['name']} own oracle {v}")
                    checked += 1
            finally:
                L.gbd_pcg_set_tuning(n, N, 0, 0, -1)
    assert checked >= 8


def test_config5_grid_kernel_tolerance_parity(torch_cuda, capi, oracle_pcg):
    """BASELINE config 5 (n = 64, N = 256, tol 1e-6, cap 200): the default launch is the tolerance-parity grid kernel
    (gbd_grid_pcg_fast.cuh); policy of SURVEY.md 8(c)(ii) against the unmodified reference kernel on this GPU (else the oracle)."""
    import mpcgpu_b200 as m
    from oracle import refgpu
    torch = torch_cuda
    n, N, cap, tol = 64, 256, 200, 1e-6
    v = capi.resolved_variant(n, N)
    assert v["fast"] and v["mode"] == 24, v
    d = synth.make_systems(n, N, batch=1, seed=4242, nan_pads=True)
    S, P, g, l0 = (d[k][0] for k in ("S", "Pinv", "gamma", "lambda0"))
    got = _gpu_solve(torch, m, S, P, g, l0, n, N, cap, tol)
    assert not got["max_iter_exit"] and np.isfinite(got["lam"]).all()
    if refgpu.available():
        rk = refgpu.solve(n, N, _dev(torch, S), _dev(torch, P), _dev(torch, g), _dev(torch, l0), cap, tol, block=128)
        ref = dict(lam=rk["lam"].cpu().numpy(), iters=rk["iters"], max_iter_exit=rk["max_iter_exit"])
    else:
        ref = oracle_pcg.pcg(S, P, g, l0, n, N, cap, tol)
    _assert_parity(oracle_pcg, got, ref, S, g, n, N, cap, "config 5 vs reference")


@pytest.mark.parametrize("n,N,cap,tol", [(14, 32, 173, 1e-6), (14, 64, 167, 1e-5), (14, 128, 167, 1e-4), (14, 128, 167, 1e-6),
                                         (14, 256, 118, 1e-5), (6, 12, 60, 1e-6)])
def test_synthetic_rings_tolerance_parity(torch_cuda, capi, oracle_pcg, n, N, cap, tol):
    """Default (fast) launch vs the bit-faithful reference oracle and, when present, the unmodified reference kernel on this GPU."""
    import mpcgpu_b200 as m
    from oracle import refgpu
    torch = torch_cuda
    d = synth.make_systems(n, N, batch=3, seed=40 + N, nan_pads=True)
    for i in range(3):
        S, P, g, l0 = (d[k][i] for k in ("S", "Pinv", "gamma", "lambda0"))
        got = _gpu_solve(torch, m, S, P, g, l0, n, N, cap, tol)
        ref = oracle_pcg.pcg(S, P, g, l0, n, N, cap, tol)
        _assert_parity(oracle_pcg, got, ref, S, g, n, N, cap, f"({n},{N}) system {i} vs oracle")
        if i == 0 and refgpu.available():
            rk = refgpu.solve(n, N, _dev(torch, S), _dev(torch, P), _dev(torch, g), _dev(torch, l0), cap, tol, block=128)
            rk = dict(lam=rk["lam"].cpu().numpy(), iters=rk["iters"], max_iter_exit=rk["max_iter_exit"])
            _assert_parity(oracle_pcg, got, rk, S, g, n, N, cap, f"({n},{N}) vs reference kernel")


@pytest.mark.parametrize("n,N,C,mode", [(14, 32, 4, 20), (14, 32, 1, 27), (14, 64, 2, 27), (14, 128, 4, 31), (32, 8, 4, 24), (32, 32, 16, 25)])
def test_exit_semantics_and_warm_start(torch_cuda, capi, oracle_pcg, n, N, C, mode):
    """pcg.cuh:195,212: iters = k+1 when iteration k passed the test, max_iter with the flag set otherwise; lambda is in/out.
    One kernel of every tolerance-parity family: single-solve, batch (one CTA per system, 2- and 4-CTA clusters), grid (flat and
    two-level exchange)."""
    import mpcgpu_b200 as m
    lanes = {24: 1, 25: 1, 26: n, 27: 0, 28: 0, 31: 0}.get(mode, 16)
    d = synth.make_systems(n, N, seed=21)
    S, P, g, l0 = (d[k][0] for k in ("S", "Pinv", "gamma", "lambda0"))
    big = 173
    assert capi.lib().gbd_pcg_set_tuning(n, N, 0, C, mode) == 0
    try:
        for cap, tol in ((3, 1e-30), (0, 1e-6), (50, 1e30), (1, 1e-30)):
            got = _gpu_solve(torch_cuda, m, S, P, g, l0, n, N, cap, tol)
            ref = oracle_pcg.pcg(S, P, g, l0, n, N, cap, tol)
            assert (got["iters"], got["max_iter_exit"]) == (ref["iters"], ref["max_iter_exit"])
            _assert_same(got, oracle_pcg.pcg_fast(S, P, g, l0, n, N, C, cap, tol, lanes=lanes), f"cap {cap} tol {tol}")
        full = _gpu_solve(torch_cuda, m, S, P, g, l0, n, N, big, 1e-7)
        warm = _gpu_solve(torch_cuda, m, S, P, g, full["lam"], n, N, big, 1e-7)
        assert warm["iters"] < full["iters"]
        _assert_same(warm, oracle_pcg.pcg_fast(S, P, g, full["lam"], n, N, C, big, 1e-7, lanes=lanes), "warm start")
    finally:
        capi.lib().gbd_pcg_set_tuning(n, N, 0, 0, -1)


@pytest.mark.parametrize("C,mode", [(2, 20), (2, 26), (1, 27), (2, 28)])
def test_batched_equals_single_and_is_deterministic(torch_cuda, capi, oracle_pcg, C, mode):
    """More systems than resident clusters, iteration counts from 1 to the cap (right-hand sides over six decades): clusters
    draw systems from the work counter in a data-dependent order, yet every system equals its own-oracle solution bit for
    bit, lands in its slot, and a second launch reproduces the first."""
    import mpcgpu_b200 as m
    torch = torch_cuda
    n, N, B, cap, tol = 14, 32, 300, 60, 1e-4
    lanes = {26: n, 27: 0, 28: 0}.get(mode, 16)
    d = synth.make_systems(n, N, batch=B, seed=77, nan_pads=True)
    scale = (10.0 ** np.random.default_rng(5).uniform(-4.0, 2.0, size=B)).astype(np.float32)
    gam = (d["gamma"] * scale[:, None]).astype(np.float32)
    S, P, g = (_dev(torch, x) for x in (d["S"], d["Pinv"], gam))
    outs = []
    assert capi.lib().gbd_pcg_set_tuning(n, N, 0, C, mode) == 0    # pin the kernel under test
    try:
        for _ in range(2):
            lam = _dev(torch, d["lambda0"])
            it = torch.zeros(B, dtype=torch.int32, device="cuda")
            fl = torch.zeros(B, dtype=torch.uint8, device="cuda")
            m.solve_batched(n, N, B, S, P, g, lam, it, fl, cap, tol)
            torch.cuda.synchronize()
            outs.append((lam.cpu().numpy(), it.cpu().numpy(), fl.cpu().numpy()))
    finally:
        capi.lib().gbd_pcg_set_tuning(n, N, 0, 0, -1)
    assert all(np.array_equal(a, b) for a, b in zip(outs[0], outs[1]))
    lam, it, fl = outs[0]
    for i in list(range(0, B, 7)) + [B - 1]:
        want = oracle_pcg.pcg_fast(d["S"][i], d["Pinv"][i], gam[i], d["lambda0"][i], n, N, C, cap, tol, lanes=lanes)
        assert int(it[i]) == want["iters"] and bool(fl[i]) == want["max_iter_exit"], i
        assert np.array_equal(lam[i], want["lam"]), i
    assert it.min() <= 5 and fl.any()


def test_full_size_batch_properties(torch_cuda, capi, oracle_pcg):
    """BASELINE config 4 size (1024 x N=128): every system's fp64 residual is within the parity bound of the bit-exact
    kernels' on the same batch, iteration counts within +-2, on all 1024 systems (GPU vs GPU)."""
    import mpcgpu_b200 as m
    torch = torch_cuda
    n, N, B, cap, tol = 14, 128, 1024, 167, 1e-4
    d = synth.make_systems(n, N, batch=B, seed=11)
    S, P, g = (_dev(torch, d[k]) for k in ("S", "Pinv", "gamma"))
    res = {}
    for name, num in (("fast", capi.NUMERICS_FAST), ("exact", capi.NUMERICS_BITEXACT)):
        prev = capi.set_numerics(num)
        try:
            lam = _dev(torch, d["lambda0"])
            it = torch.zeros(B, dtype=torch.int32, device="cuda")
            fl = torch.zeros(B, dtype=torch.uint8, device="cuda")
            m.solve_batched(n, N, B, S, P, g, lam, it, fl, cap, tol)
            torch.cuda.synchronize()
            res[name] = (lam.cpu().numpy(), it.cpu().numpy(), fl.cpu().numpy())
        finally:
            capi.set_numerics(prev)
    lf, itf, ff = res["fast"]
    le, ite, fe = res["exact"]
    assert np.abs(itf.astype(int) - ite.astype(int)).max() <= ITER_TOL
    near_cap = ite >= cap - ITER_TOL
    assert np.array_equal(ff[~near_cap], fe[~near_cap])
    err = np.abs(lf.astype(np.float64) - le).max(axis=1) / np.abs(le).max(axis=1)
    assert err.max() <= LAMBDA_TOL, err.max()
    for i in range(0, B, 64):
        rf = oracle_pcg.rel_residual(d["S"][i], d["gamma"][i], lf[i], n, N)
        re_ = oracle_pcg.rel_residual(d["S"][i], d["gamma"][i], le[i], n, N)
        assert rf <= RESID_FACTOR * re_ + 1e-6


@pytest.mark.parametrize("knots,block", [(32, 160), (64, 160), (128, 160), (32, 128), (64, 96), (128, 128)])
def test_dropin_headers_fast_body_reference_launch_geometry(torch_cuda, oracle_pcg, tmp_path, knots, block):
    """include/gbd_dropin built with -DGBD_DROPIN_FAST=1: pcg<float,14,N> launched exactly like include/pcg/sqp.cuh:230
    (cooperative, grid = N, smem = pcgSharedMemSize).  Blocks of >= 160 threads run the tolerance-parity body (clusters of N/8
    CTAs): bit-exact vs ITS oracle.  Smaller blocks run the bit-exact body: bit-exact vs the reference-order oracle."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", f"dropin_demo_fast_{knots}")
    if not os.path.exists(exe):
        pytest.skip("tests/_build/dropin_demo_fast_* not built (run __graft_entry__.build())")
    n, cap, tol = 14, 167, 1e-5
    d = synth.make_systems(n, knots, seed=9, nan_pads=True)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    np.concatenate([d["S"][0], d["Pinv"][0], d["gamma"][0], d["lambda0"][0]]).astype(np.float32).tofile(fin)
    subprocess.check_call([exe, str(fin), str(fout), str(cap), repr(tol), str(block), "3"], timeout=120)
    raw = np.fromfile(fout, np.float32)
    vec = n * knots
    tail = raw[3 * vec:].view(np.uint32)
    got = dict(lam=raw[:vec], r=raw[vec:2 * vec], p=raw[2 * vec:3 * vec], iters=int(tail[0]), max_iter_exit=bool(tail[1]))
    if block >= 160:
        want = oracle_pcg.pcg_fast(d["S"][0], d["Pinv"][0], d["gamma"][0], d["lambda0"][0], n, knots, knots // 8, cap, tol)
    else:
        want = oracle_pcg.pcg(d["S"][0], d["Pinv"][0], d["gamma"][0], d["lambda0"][0], n, knots, cap, tol)
    _assert_same(got, want, f"fast drop-in N={knots} block={block}")


@pytest.mark.parametrize("n,N,C,mode,batch", [(14, 32, 4, 20, 1), (14, 32, 1, 27, 5), (14, 64, 2, 27, 3), (32, 8, 4, 24, 1)])
def test_unaligned_pointers_take_the_non_tma_path(torch_cuda, capi, oracle_pcg, n, N, C, mode, batch):
    """S / Pinv that are only 4-byte aligned cannot be staged by TMA bulk copies: every tolerance-parity family falls back to plain
    loads and still equals its oracle bit for bit."""
    import mpcgpu_b200 as m
    torch = torch_cuda
    lanes = {24: 1, 25: 1, 26: n, 27: 0, 28: 0, 31: 0}.get(mode, 16)
    cap, tol = 50, 1e-6
    d = synth.make_systems(n, N, batch=batch, seed=17)
    mat = 3 * n * n * N
    bigS, bigP = torch.zeros(batch * mat + 1, device="cuda"), torch.zeros(batch * mat + 1, device="cuda")
    S, P = bigS[1:], bigP[1:]                                 # 4-byte aligned only
    S.copy_(_dev(torch, d["S"]).reshape(-1))
    P.copy_(_dev(torch, d["Pinv"]).reshape(-1))
    assert S.data_ptr() % 16 != 0 and P.data_ptr() % 16 != 0
    g, lam = _dev(torch, d["gamma"]).reshape(-1), _dev(torch, d["lambda0"]).reshape(-1)
    it = torch.zeros(batch, dtype=torch.int32, device="cuda")
    fl = torch.zeros(batch, dtype=torch.uint8, device="cuda")
    assert capi.lib().gbd_pcg_set_tuning(n, N, 0, C, mode) == 0
    try:
        if batch == 1:
            m.pcg_launch(n, N, S, P, g, lam, None, None, None, None, it, fl, cap, tol)
        else:
            m.solve_batched(n, N, batch, S, P, g, lam, it, fl, cap, tol)
        torch.cuda.synchronize()
    finally:
        capi.lib().gbd_pcg_set_tuning(n, N, 0, 0, -1)
    lam = lam.cpu().numpy().reshape(batch, n * N)
    for i in range(batch):
        want = oracle_pcg.pcg_fast(d["S"][i], d["Pinv"][i], d["gamma"][i], d["lambda0"][i], n, N, C, cap, tol, lanes=lanes)
        assert int(it[i]) == want["iters"] and bool(fl[i]) == want["max_iter_exit"], i
        assert np.array_equal(lam[i], want["lam"]), i
