"""GPU parity tests for rows f1 / f2 of SURVEY.md 8f: Schur-complement + preconditioner assembly (form_schur_system,
include/pcg/linsys_setup.cuh:621-657) and step recovery (compute_dz, include/common/dz.cuh:125-136), through the C ABI.

Tolerance: the kernels keep the reference's floating-point operation order, so the bar is BIT-EXACT (== on every
element, +0/-0 equal) against the C oracle and -- when oracle/_ref/libref_schur.so is present -- against the
reference's own kernels run on the same GPU.  Pad tiles (left of block row 0, right of block row N-1) are excluded:
the reference never writes them."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    torch.cuda.init()
    return torch


def _mask_pads(x, n, N):
    x = np.array(x, np.float32).reshape(N, 3, n, n).copy()
    x[0, 0] = 0
    x[N - 1, 2] = 0
    return x


@pytest.fixture(params=[0, 1], ids=["cta_per_row", "warp_per_row"])
def team(request):
    """Both mappings of the assembly's first phase (gbd_schur_set_team): the latency launch and the batch launch."""
    from mpcgpu_b200 import _capi
    prev = _capi.lib().gbd_schur_set_team(request.param)
    yield request.param
    _capi.lib().gbd_schur_set_team(prev)


def _ours(torch, n, m, N, G, C, g, c, rho):
    import mpcgpu_b200 as mp
    dG, dC, dg, dc = (torch.from_numpy(x.copy()).cuda() for x in (G, C, g, c))
    dS = torch.full((3 * n * n * N,), float("nan"), device="cuda")
    dP = torch.full((3 * n * n * N,), float("nan"), device="cuda")
    dgam = torch.full((n * N,), float("nan"), device="cuda")
    mp.form_schur_system(n, m, N, dG, dC, dg, dc, dS, dP, dgam, rho)
    torch.cuda.synchronize()
    return dict(S=dS, Pinv=dP, gamma=dgam, Ginv=dG, C=dC, g=dg)


@pytest.mark.parametrize("n,m,N", [(14, 7, 8), (14, 7, 32), (14, 7, 128), (14, 7, 512), (6, 3, 12), (4, 2, 5), (2, 1, 3)])
def test_form_schur_bit_exact_vs_oracle(torch_cuda, team, n, m, N):
    from oracle import schur
    G, C, g, c = schur.make_kkt(n, m, N, seed=100 + n + N)
    want = schur.form(G, C, g, c, n, m, N, 1e-3)
    got = _ours(torch_cuda, n, m, N, G, C, g, c, 1e-3)
    assert np.array_equal(got["gamma"].cpu().numpy(), want["gamma"])
    assert np.array_equal(got["Ginv"].cpu().numpy(), want["Ginv"])
    for k in ("S", "Pinv"):
        assert np.array_equal(_mask_pads(got[k].cpu().numpy(), n, N), _mask_pads(want[k], n, N)), k


@pytest.mark.parametrize("n,m,N", [(14, 7, 32), (14, 7, 128), (6, 3, 12)])
def test_form_schur_and_dz_bit_exact_vs_reference_kernels(torch_cuda, team, n, m, N):
    """A/B against the reference's own form_schur_system / compute_dz compiled for sm_100a (oracle/_ref/libref_schur.so)."""
    torch = torch_cuda
    from oracle import refgpu, schur
    if not refgpu.schur_available():
        pytest.skip("oracle/_ref/libref_schur.so not built (make -C oracle ref needs /root/reference)")
    import mpcgpu_b200 as mp
    G, C, g, c = schur.make_kkt(n, m, N, seed=7)
    got = _ours(torch, n, m, N, G, C, g, c, 1e-3)
    rG, rC, rg, rc = (torch.from_numpy(x.copy()).cuda() for x in (G, C, g, c))
    rS = torch.zeros(3 * n * n * N, device="cuda")
    rP = torch.zeros(3 * n * n * N, device="cuda")
    rgam = torch.zeros(n * N, device="cuda")
    refgpu.form_schur_system(n, m, N, rG, rC, rg, rc, rS, rP, rgam, 1e-3)
    torch.cuda.synchronize()
    assert np.array_equal(got["gamma"].cpu().numpy(), rgam.cpu().numpy())
    assert np.array_equal(got["Ginv"].cpu().numpy(), rG.cpu().numpy())
    for k, r in (("S", rS), ("Pinv", rP)):
        assert np.array_equal(_mask_pads(got[k].cpu().numpy(), n, N), _mask_pads(r.cpu().numpy(), n, N)), k
    lam = torch.from_numpy(np.random.default_rng(1).standard_normal(n * N).astype(np.float32)).cuda()
    dz_o = torch.zeros((n + m) * (N - 1) + n, device="cuda")
    dz_r = torch.zeros_like(dz_o)
    mp.compute_dz(n, m, N, got["Ginv"], got["C"], got["g"], lam, dz_o)
    refgpu.compute_dz(n, m, N, rG, rC, rg, lam, dz_r)
    torch.cuda.synchronize()
    assert np.array_equal(dz_o.cpu().numpy(), dz_r.cpu().numpy())


@pytest.mark.parametrize("n,m,N", [(14, 7, 32), (14, 7, 128)])
def test_assemble_solve_recover_pipeline_vs_oracle(torch_cuda, n, m, N):
    """form_schur_system -> pcg -> compute_dz on the GPU equals the same chain of oracles bit for bit."""
    torch = torch_cuda
    import mpcgpu_b200 as mp
    from oracle import pcg as opcg
    from oracle import schur
    G, C, g, c = schur.make_kkt(n, m, N, seed=21)
    got = _ours(torch, n, m, N, G, C, g, c, 1e-3)
    lam = torch.zeros(n * N, device="cuda")
    it = torch.zeros(1, dtype=torch.int32, device="cuda")
    fl = torch.zeros(1, dtype=torch.uint8, device="cuda")
    mp.pcg_launch(n, N, got["S"], got["Pinv"], got["gamma"], lam, None, None, None, None, it, fl, 200, 1e-6)
    dz = torch.zeros((n + m) * (N - 1) + n, device="cuda")
    mp.compute_dz(n, m, N, got["Ginv"], got["C"], got["g"], lam, dz)
    torch.cuda.synchronize()
    o = schur.form(G, C, g, c, n, m, N, 1e-3)
    w = opcg.pcg(o["S"], o["Pinv"], o["gamma"], np.zeros(n * N, np.float32), n, N, 200, 1e-6)
    assert int(it.item()) == w["iters"] and bool(fl.item()) == w["max_iter_exit"]
    assert np.array_equal(lam.cpu().numpy(), w["lam"])
    assert np.array_equal(dz.cpu().numpy(), schur.dz(o["Ginv"], C, g, w["lam"], n, m, N))


def test_schur_argument_errors(torch_cuda):
    from mpcgpu_b200 import _capi
    L = _capi.lib()
    assert L.gbd_schur_supported(14, 7) == 1 and L.gbd_schur_supported(14, 6) == 0
    assert L.gbd_form_schur_system_f32(14, 6, 8, 1, 1, 1, 1, 1, 1, 1, 1e-3, 0) == _capi.ERR_UNSUPPORTED
    assert L.gbd_form_schur_system_f32(14, 7, 8, 0, 1, 1, 1, 1, 1, 1, 1e-3, 0) == _capi.ERR_BADARG
    assert L.gbd_compute_dz_f32(14, 7, 1, 1, 1, 1, 1, 1, 0) == _capi.ERR_BADARG


def test_form_schur_and_dz_on_reference_minted_iiwa_vectors(torch_cuda):
    """tests/golden/schur_iiwa_*.npz: inputs from the reference's generate_kkt_submatrices on examples/trajfiles/0_0_*, answers
    from the reference's form_schur_system / pcg<> / compute_dz (tools/make_golden_schur.py) -- bit-exact."""
    import glob
    import os
    import mpcgpu_b200 as mp
    torch = torch_cuda
    paths = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "schur_iiwa_*.npz")))
    assert paths
    for path in paths:
        z = np.load(path)
        n, m, N, rho = int(z["n"]), int(z["m"]), int(z["N"]), float(z["rho"])
        got = _ours(torch, n, m, N, z["G"], z["C"], z["g"], z["c"], rho)
        assert np.array_equal(got["gamma"].cpu().numpy(), z["gamma"])
        assert np.array_equal(got["Ginv"].cpu().numpy(), z["Ginv"])
        for k in ("S", "Pinv"):
            assert np.array_equal(_mask_pads(got[k].cpu().numpy(), n, N), _mask_pads(z[k], n, N)), k
        # solve with the reference's tolerance / cap, then recover dz: the whole chain equals the reference's
        lam = torch.zeros(n * N, device="cuda")
        it = torch.zeros(1, dtype=torch.int32, device="cuda")
        fl = torch.zeros(1, dtype=torch.uint8, device="cuda")
        mp.pcg_launch(n, N, got["S"], got["Pinv"], got["gamma"], lam, None, None, None, None, it, fl, 173, 1e-4)
        dz = torch.zeros((n + m) * (N - 1) + n, device="cuda")
        mp.compute_dz(n, m, N, got["Ginv"], got["C"], got["g"], lam, dz)
        torch.cuda.synchronize()
        assert int(it.item()) == int(z["pcg_iters"])
        assert np.array_equal(lam.cpu().numpy(), z["lam"])
        assert np.array_equal(dz.cpu().numpy(), z["dz"])


@pytest.mark.parametrize("n,m,N,B", [(14, 7, 32, 5), (14, 7, 128, 3)])
def test_batched_step_plan_equals_per_system_oracle_chain(torch_cuda, team, n, m, N, B):
    """gbd_step_run_f32 (row f3): assembly -> warm-started solve -> dz for a batch, one enqueue; every trajectory equals the
    oracle chain bit for bit, and the flags the multi-GPU driver gathers are the per-trajectory max_iter_exit."""
    torch = torch_cuda
    import mpcgpu_b200 as mp
    from mpcgpu_b200 import sharding
    from oracle import pcg as opcg
    from oracle import schur
    kk = [schur.make_kkt(n, m, N, seed=300 + i) for i in range(B)]
    G, C, g, c = (np.concatenate([k[j] for k in kk]) for j in range(4))
    lam0 = (0.01 * np.random.default_rng(3).standard_normal(B * n * N)).astype(np.float32)
    dG, dC, dg, dc, dl = (torch.from_numpy(x.copy()).cuda() for x in (G, C, g, c, lam0))
    dz = torch.zeros(B * ((n + m) * (N - 1) + n), device="cuda")
    plan = mp.StepPlan(n, m, N, B)
    cap, tol = 30, 1e-5                                   # a cap some trajectories hit, so both flag values occur
    plan.run(dG, dC, dg, dc, 1e-3, dl, dz, cap, tol)
    it, fl = plan.results()
    flags_dev = plan.device_flags().cpu().numpy()
    gathered = sharding.gather_converged(plan.device_flags(), B, 1, 0).cpu().numpy()
    nl, nz = n * N, (n + m) * (N - 1) + n
    for i in range(B):
        o = schur.form(kk[i][0], kk[i][1], kk[i][2], kk[i][3], n, m, N, 1e-3)
        w = opcg.pcg(o["S"], o["Pinv"], o["gamma"], lam0[i * nl:(i + 1) * nl], n, N, cap, tol)
        assert int(it[i]) == w["iters"] and bool(fl[i]) == w["max_iter_exit"], i
        assert np.array_equal(dl.cpu().numpy()[i * nl:(i + 1) * nl], w["lam"]), i
        assert np.array_equal(dz.cpu().numpy()[i * nz:(i + 1) * nz], schur.dz(o["Ginv"], kk[i][1], kk[i][2], w["lam"], n, m, N)), i
    assert np.array_equal(flags_dev, fl) and np.array_equal(gathered, fl)
    plan.close()


@pytest.mark.parametrize("n,N", [(14, 32), (14, 128), (6, 12), (2, 3)])
def test_csr_wire_format_bit_exact_and_feeds_qdldl(torch_cuda, n, N):
    """Row f4: band S -> upper-triangular CSC (include/utils/csr.cuh:10-74).  Index/byte work: identical to the oracle's
    restatement; and the reference's own QDLDL, fed with the GPU-packed matrix, agrees with the GPU PCG solution."""
    torch = torch_cuda
    import mpcgpu_b200 as mp
    from mpcgpu_b200 import _capi, synth
    from oracle import qdldl
    if not qdldl.available():
        pytest.skip("oracle/_ref/libqdldl_ref.so not built")
    L = _capi.lib()
    d = synth.make_systems(n, N, batch=1, seed=5, nan_pads=True)
    S = torch.from_numpy(d["S"][0]).cuda()
    nnz = L.gbd_schur_csr_nnz(n, N)
    assert nnz == qdldl.nnz(n, N)
    cp = torch.full((n * N + 1,), -1, dtype=torch.int32, device="cuda")
    ri = torch.full((nnz,), -1, dtype=torch.int32, device="cuda")
    val = torch.full((nnz,), float("nan"), device="cuda")
    assert L.gbd_schur_csr_pattern_i32(n, N, cp.data_ptr(), ri.data_ptr(), 0) == 0
    assert L.gbd_schur_csr_values_f32(n, N, S.data_ptr(), val.data_ptr(), 0) == 0
    torch.cuda.synchronize()
    wcp, wri = qdldl.pattern(n, N)
    assert np.array_equal(cp.cpu().numpy(), wcp) and np.array_equal(ri.cpu().numpy(), wri)
    wval = qdldl.values(d["S"][0], n, N)[0]
    assert np.array_equal(val.cpu().numpy(), wval)
    if n == 14:
        # the reference's QDLDL on the GPU-packed matrix vs our PCG on the band form of the same system
        import ctypes as C
        ws = qdldl.lib().qdldl_ref_create(n, N)
        x = np.zeros(n * N, np.float32)
        v = np.ascontiguousarray(val.cpu().numpy())
        g = np.ascontiguousarray(d["gamma"][0])
        fp = C.POINTER(C.c_float)
        assert qdldl.lib().qdldl_ref_solve(ws, v.ctypes.data_as(fp), g.ctypes.data_as(fp), x.ctypes.data_as(fp)) >= 0
        qdldl.lib().qdldl_ref_destroy(ws)
        P, gam = torch.from_numpy(d["Pinv"][0]).cuda(), torch.from_numpy(d["gamma"][0]).cuda()
        lam = torch.zeros(n * N, device="cuda")
        it = torch.zeros(1, dtype=torch.int32, device="cuda")
        fl = torch.zeros(1, dtype=torch.uint8, device="cuda")
        mp.pcg_launch(n, N, S, P, gam, lam, None, None, None, None, it, fl, 500, 1e-9)
        torch.cuda.synchronize()
        lam = lam.cpu().numpy()
        assert np.abs(lam - x).max() <= 2e-3 * np.abs(x).max()
