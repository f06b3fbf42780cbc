"""GPU tests of the direct solver (block cyclic reduction, include/gbd/gbd_bcr.cuh) -- SURVEY.md 8f row f4.

This is NOT the reference's algorithm (the reference has pcg<> and, on the CPU, QDLDL): the bar is a floating-point
tolerance, stated here.  Synthetic systems (condition number 3e2 .. 5e3): relative error against the fp64 solution of the
same band system < 1e-3 (measured: 1e-5 .. 2e-5) and fp64 relative residual ||gamma - S lam|| / ||gamma|| < 5e-5.
Reference-minted IIWA systems (condition number 6e4 .. 2e7, fp32): relative error < 2e-2 (measured 3e-5 .. 6e-3), residual
< 5e-3 and never larger than the residual the reference's own pcg<> leaves on the same system at its documented tolerance
(measured: 13 .. 27 %, relative error of lambda 12 .. 100 %); agreement with the reference's own QDLDL -- the CPU direct
path this row stands beside, itself fp32 -- within 5e-2 relative."""
import numpy as np
import pytest

from mpcgpu_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    torch.cuda.init()
    return torch


def _solve(torch, n, N, S, g, batch=1):
    import mpcgpu_b200 as mp
    dS, dg = torch.from_numpy(np.ascontiguousarray(S)).cuda(), torch.from_numpy(np.ascontiguousarray(g)).cuda()
    lam = torch.full((batch * n * N,), float("nan"), device="cuda")
    mp.solve_direct(n, N, dS.reshape(-1), dg.reshape(-1), lam, batch=batch)
    torch.cuda.synchronize()
    return lam.cpu().numpy()


@pytest.mark.parametrize("n,N", [(14, 8), (14, 16), (14, 32), (14, 64), (14, 128), (14, 256), (14, 512), (6, 16), (2, 4)])
def test_direct_solve_vs_fp64_truth(torch_cuda, oracle_pcg, n, N):
    d = synth.make_systems(n, N, batch=2, seed=40 + N, nan_pads=True)
    for i in range(2):
        lam = _solve(torch_cuda, n, N, d["S"][i], d["gamma"][i])
        truth = oracle_pcg.solve_f64(d["S"][i], d["gamma"][i], n, N)
        assert np.isfinite(lam).all()
        rel = np.abs(lam - truth).max() / np.abs(truth).max()
        res = oracle_pcg.rel_residual(d["S"][i], d["gamma"][i], lam, n, N)
        assert res < 5e-5, (n, N, res)
        assert rel < 1e-3, (n, N, rel)


def test_direct_solve_batched_matches_single(torch_cuda):
    n, N, B = 14, 128, 40                      # more systems than resident clusters: the persistent loop is exercised
    d = synth.make_systems(n, N, batch=B, seed=77)
    lam = _solve(torch_cuda, n, N, d["S"], d["gamma"], batch=B).reshape(B, n * N)
    for i in (0, 1, 17, B - 1):
        one = _solve(torch_cuda, n, N, d["S"][i], d["gamma"][i])
        assert np.array_equal(lam[i], one)


def test_direct_solve_on_reference_minted_iiwa_systems(torch_cuda, oracle_pcg):
    """tests/golden/iiwa_*.npz (reference KKT + Schur assembly): where the reference's PCG stops at its cap, the direct solve
    is complete; it agrees with the reference's own QDLDL on the same matrix."""
    import glob
    import os
    import ctypes as C
    from oracle import qdldl
    paths = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "iiwa_*.npz")))
    assert paths
    for path in paths:
        z = np.load(path)
        n, N = int(z["n"]), int(z["N"])
        lam = _solve(torch_cuda, n, N, z["S"], z["gamma"])
        res = oracle_pcg.rel_residual(np.nan_to_num(z["S"]), z["gamma"], lam, n, N)
        ref_res = oracle_pcg.rel_residual(np.nan_to_num(z["S"]), z["gamma"], z["run0_lam"], n, N)
        truth = oracle_pcg.solve_f64(np.nan_to_num(z["S"]), z["gamma"], n, N)
        assert res < 5e-3 and res <= ref_res, (path, res, ref_res)
        assert np.abs(lam - truth).max() <= 2e-2 * np.abs(truth).max(), path
        if qdldl.available():
            S0 = np.nan_to_num(z["S"]).astype(np.float32)
            val = qdldl.values(S0, n, N)[0]
            ws = qdldl.lib().qdldl_ref_create(n, N)
            x = np.zeros(n * N, np.float32)
            g = np.ascontiguousarray(z["gamma"], np.float32)
            fp = C.POINTER(C.c_float)
            assert qdldl.lib().qdldl_ref_solve(ws, val.ctypes.data_as(fp), g.ctypes.data_as(fp), x.ctypes.data_as(fp)) >= 0
            qdldl.lib().qdldl_ref_destroy(ws)
            assert np.abs(lam - x).max() <= 5e-2 * np.abs(x).max(), path


def test_direct_argument_errors(torch_cuda):
    from mpcgpu_b200 import _capi
    L = _capi.lib()
    assert L.gbd_bcr_supported(14, 128) == 1 and L.gbd_bcr_supported(14, 48) == 0
    assert L.gbd_bcr_solve_f32(14, 48, 1, 1, 1, 0) == _capi.ERR_UNSUPPORTED
    assert L.gbd_bcr_solve_f32(14, 128, 0, 1, 1, 0) == _capi.ERR_BADARG


def test_step_with_direct_fallback(torch_cuda, oracle_pcg):
    """gbd_step_run_fallback_f32: trajectories whose PCG solve hits the cap get the direct solution (fp64 residual small), the
    converged ones keep their PCG solution bit for bit; dz is formed from whichever lambda was kept."""
    torch = torch_cuda
    import mpcgpu_b200 as mp
    from oracle import schur
    n, m, N, B = 14, 7, 32, 6
    kk = [schur.make_kkt(n, m, N, seed=500 + i) for i in range(B)]
    G, C, g, c = (np.concatenate([k[j] for k in kk]) for j in range(4))
    out = {}
    for fb in (False, True):
        dG, dC, dg, dc = (torch.from_numpy(x.copy()).cuda() for x in (G, C, g, c))
        dl = torch.zeros(B * n * N, device="cuda")
        dz = torch.zeros(B * ((n + m) * (N - 1) + n), device="cuda")
        plan = mp.StepPlan(n, m, N, B)
        plan.run(dG, dC, dg, dc, 1e-3, dl, dz, 12, 1e-7, direct_fallback=fb)      # cap of 12: most trajectories hit it
        it, fl = plan.results()
        out[fb] = (dl.cpu().numpy().reshape(B, -1), dz.cpu().numpy().reshape(B, -1), fl.copy())
        plan.close()
    assert np.array_equal(out[False][2], out[True][2]) and out[True][2].any()
    nz = (n + m) * (N - 1) + n
    for i in range(B):
        o = schur.form(kk[i][0], kk[i][1], kk[i][2], kk[i][3], n, m, N, 1e-3, pad=0.0)
        if out[True][2][i]:
            res_fb = oracle_pcg.rel_residual(o["S"], o["gamma"], out[True][0][i], n, N)
            res_pcg = oracle_pcg.rel_residual(o["S"], o["gamma"], out[False][0][i], n, N)
            assert res_fb < 5e-5 and res_fb < res_pcg, (i, res_fb, res_pcg)
            assert np.array_equal(out[True][1][i], schur.dz(o["Ginv"], kk[i][1], kk[i][2], out[True][0][i], n, m, N))
        else:
            assert np.array_equal(out[True][0][i], out[False][0][i]) and np.array_equal(out[True][1][i], out[False][1][i])


@pytest.mark.parametrize("knots,block", [(32, 128), (128, 128), (128, 64)])
def test_dropin_headers_direct_body_reference_launch_geometry(torch_cuda, oracle_pcg, tmp_path, knots, block):
    """include/gbd_dropin built with -DGBD_DROPIN_DIRECT=1: pcg<float,14,N> launched exactly like include/pcg/sqp.cuh:230 solves the
    system by block cyclic reduction (block of 128 threads: bit-identical to the C-ABI direct solver, iters = 0, flag = 0); any
    other block size runs the bit-exact PCG body."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", f"dropin_demo_direct_{knots}")
    if not os.path.exists(exe):
        pytest.skip("tests/_build/dropin_demo_direct_* not built (run __graft_entry__.build())")
    n, cap, tol = 14, 167, 1e-5
    d = synth.make_systems(n, knots, seed=9, nan_pads=True)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    np.concatenate([d["S"][0], d["Pinv"][0], d["gamma"][0], d["lambda0"][0]]).astype(np.float32).tofile(fin)
    subprocess.check_call([exe, str(fin), str(fout), str(cap), repr(tol), str(block), "3"], timeout=120)
    raw = np.fromfile(fout, np.float32)
    vec = n * knots
    tail = raw[3 * vec:].view(np.uint32)
    lam, iters, flag = raw[:vec], int(tail[0]), bool(tail[1])
    if block == 128:
        assert iters == 0 and not flag
        assert np.array_equal(lam, _solve(torch_cuda, n, knots, d["S"][0], d["gamma"][0]))
        truth = oracle_pcg.solve_f64(d["S"][0], d["gamma"][0], n, knots)
        assert np.abs(lam - truth).max() / np.abs(truth).max() < 1e-3
    else:
        want = oracle_pcg.pcg(d["S"][0], d["Pinv"][0], d["gamma"][0], d["lambda0"][0], n, knots, cap, tol)
        assert iters == want["iters"] and flag == want["max_iter_exit"] and np.array_equal(lam, want["lam"])
