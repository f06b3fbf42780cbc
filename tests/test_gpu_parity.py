"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI.

Tolerance: the CUDA path keeps the reference's floating-point operation ORDER (sequential FMA band
rows, GLASS trees, IEEE division), so against the oracle -- and against the unmodified reference
kernel when oracle/_ref/libref_gbdpcg.so is present -- the bar is BIT-EXACT lambda, iteration count
and exit flag (values compared with ==, so +0/-0 are equal).  north_star only asks for an fp32
tolerance on residual norm and iteration count; bit-exactness is strictly stronger.
"""
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


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _gpu_solve(torch, m, d, i, max_iter, tol, dtype=np.float32):
    n, N = d["n"], d["N"]
    S, P, g = (_dev(torch, d[k][i].astype(dtype)) for k in ("S", "Pinv", "gamma"))
    lam = _dev(torch, d["lambda0"][i].astype(dtype))
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    r = torch.full((n * N,), float("nan"), dtype=tdt, device="cuda")
    p = torch.full((n * N,), float("nan"), dtype=tdt, device="cuda")
    it = torch.full((1,), -1, dtype=torch.int32, device="cuda")
    fl = torch.full((1,), 7, dtype=torch.uint8, device="cuda")
    m.pcg_launch(n, N, S, P, g, lam, r, p, None, None, it, fl, max_iter, tol)
    torch.cuda.synchronize()
    return dict(lam=lam.cpu().numpy(), iters=int(it.item()), max_iter_exit=bool(fl.item()), r=r.cpu().numpy(),
                p=p.cpu().numpy())


def _assert_same(a, b, what=""):
    assert a["iters"] == b["iters"], f"{what}: iters {a['iters']} vs {b['iters']}"
    assert a["max_iter_exit"] == b["max_iter_exit"], what
    for k in ("lam", "r", "p"):
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), f"{what}: {k} differs, max abs " \
            f"{np.abs(np.asarray(a[k], np.float64) - np.asarray(b[k], np.float64)).max()}"


def test_every_variant_bit_exact_vs_oracle(torch_cuda, capi, oracle_pcg):
    """Each compiled (n, N, cluster, residency, dtype) variant against the C oracle, NaN pad tiles."""
    import mpcgpu_b200 as m
    L = capi.lib()
    cache = {}
    for v in capi.variants():
        if v["fast"]:
            continue            # tolerance-parity kernels: tests/test_gpu_fast.py
        n, N, f64 = v["n"], v["N"], v["f64"]
        dt = np.float64 if f64 else np.float32
        key = (n, N, f64)
        if key not in cache:
            d = synth.make_systems(n, N, batch=2, seed=100 + n + N, nan_pads=True, dtype=dt)
            cap, tol = (40, 1e-7) if N > 128 else (60, 1e-6)
            want = [oracle_pcg.pcg(d["S"][i], d["Pinv"][i], d["gamma"][i], d["lambda0"][i], n, N, cap, tol)
                    for i in range(2)]
            cache[key] = (d, cap, tol, want)
        d, cap, tol, want = cache[key]
        assert L.gbd_pcg_set_tuning(n, N, int(f64), v["cluster"], v["mode"]) == 0
        try:
            for i in range(2):
                got = _gpu_solve(torch_cuda, m, d, i, cap, tol, dt)
                _assert_same(got, want[i], f"variant {v} system {i}")
                assert np.isfinite(got["lam"]).all()
        finally:
            L.gbd_pcg_set_tuning(n, N, int(f64), 0, -1)


@pytest.mark.parametrize("n,N,cap,tol", [(14, 32, 173, 1e-6), (14, 128, 167, 1e-4), (14, 128, 167, 1e-6),
                                         (14, 512, 67, 1e-5), (14, 256, 118, 1e-5), (14, 64, 167, 1e-5)])
def test_reference_configs_vs_oracle_and_truth(torch_cuda, capi, oracle_pcg, n, N, cap, tol):
    """BASELINE.json configs 1-3 sizes with the reference's iteration caps (settings.cuh:123-138)."""
    import mpcgpu_b200 as m
    d = synth.make_systems(n, N, batch=1, seed=7)
    got = _gpu_solve(torch_cuda, m, d, 0, cap, tol)
    want = oracle_pcg.pcg(d["S"][0], d["Pinv"][0], d["gamma"][0], d["lambda0"][0], n, N, cap, tol)
    _assert_same(got, want, f"({n},{N})")
    # size-independent property: the solution solves the system (fp64 residual), unless the cap was hit
    if not got["max_iter_exit"]:
        assert oracle_pcg.rel_residual(d["S"][0], d["gamma"][0], got["lam"], n, N) < 5e-3


def test_exit_semantics_on_gpu(torch_cuda, capi, oracle_pcg):
    import mpcgpu_b200 as m
    n, N = 14, 32
    d = synth.make_systems(n, N, seed=21)
    for cap, tol in ((3, 1e-30), (0, 1e-6), (50, 1e30), (1, 1e-30)):
        got = _gpu_solve(torch_cuda, m, d, 0, cap, tol)
        want = oracle_pcg.pcg(d["S"][0], d["Pinv"][0], d["gamma"][0], d["lambda0"][0], n, N, cap, tol)
        assert (got["iters"], got["max_iter_exit"]) == (want["iters"], want["max_iter_exit"])
        assert np.array_equal(got["lam"], want["lam"])
    # warm start: lambda is in/out
    d2 = dict(d)
    full = _gpu_solve(torch_cuda, m, d, 0, 173, 1e-7)
    d2["lambda0"] = full["lam"][None]
    warm = _gpu_solve(torch_cuda, m, d2, 0, 173, 1e-7)
    wwant = oracle_pcg.pcg(d["S"][0], d["Pinv"][0], d["gamma"][0], full["lam"], n, N, 173, 1e-7)
    _assert_same(warm, wwant, "warm start")
    assert warm["iters"] < full["iters"]


def test_demo_system_G1_on_gpu(torch_cuda, capi, oracle_pcg, g1):
    import mpcgpu_b200 as m
    d = dict(n=2, N=3, S=g1["S"][None], Pinv=g1["Pinv"][None], gamma=g1["gamma"][None],
             lambda0=np.zeros((1, 6), np.float32))
    got = _gpu_solve(torch_cuda, m, d, 0, 100, 1e-10)
    want = oracle_pcg.pcg(g1["S"], g1["Pinv"], g1["gamma"], np.zeros(6, np.float32), 2, 3, 100, 1e-10)
    _assert_same(got, want, "G1")
    np.testing.assert_allclose(got["lam"], g1["lam"], rtol=2e-3)


def test_batched_equals_single_and_oracle(torch_cuda, capi, oracle_pcg):
    """More systems than co-resident clusters: exercises the persistent loop and smem reuse."""
    import mpcgpu_b200 as m
    torch = torch_cuda
    n, N, B, cap, tol = 14, 32, 300, 173, 1e-4
    d = synth.make_systems(n, N, batch=B, seed=33, nan_pads=True)
    S, P, g, lam = (_dev(torch, d[k]) for k in ("S", "Pinv", "gamma", "lambda0"))
    it = torch.zeros(B, dtype=torch.int32, device="cuda")
    fl = torch.zeros(B, dtype=torch.uint8, device="cuda")
    m.solve_batched(n, N, B, S, P, g, lam, it, fl, cap, tol)
    torch.cuda.synchronize()
    want = oracle_pcg.pcg_batched(d["S"], d["Pinv"], d["gamma"], d["lambda0"], n, N, B, cap, tol)
    assert np.array_equal(it.cpu().numpy().astype(np.uint32), want["iters"])
    assert np.array_equal(fl.cpu().numpy().astype(bool), want["max_iter_exit"])
    assert np.array_equal(lam.cpu().numpy(), want["lam"])
    assert len(set(want["iters"].tolist())) > 1          # systems really differ


def test_batched_uneven_iteration_counts(torch_cuda, capi, oracle_pcg):
    """Right-hand sides scaled over six decades: solves of 1 .. cap iterations in one launch, more systems than resident
    clusters -- the clusters draw systems from the work counter in a data-dependent order; every system must still equal
    the oracle bit for bit and land in its own slot."""
    import mpcgpu_b200 as m
    torch = torch_cuda
    n, N, B, cap, tol = 14, 32, 260, 60, 1e-4
    d = synth.make_systems(n, N, batch=B, seed=77)
    scale = (10.0 ** np.random.default_rng(5).uniform(-4.0, 2.0, size=B)).astype(np.float32)
    gam = (d["gamma"] * scale[:, None]).astype(np.float32)
    S, P, g, lam = (_dev(torch, x) for x in (d["S"], d["Pinv"], gam, d["lambda0"]))
    it = torch.zeros(B, dtype=torch.int32, device="cuda")
    fl = torch.zeros(B, dtype=torch.uint8, device="cuda")
    m.solve_batched(n, N, B, S, P, g, lam, it, fl, cap, tol)
    torch.cuda.synchronize()
    want = oracle_pcg.pcg_batched(d["S"], d["Pinv"], gam, d["lambda0"], n, N, B, cap, tol)
    assert np.array_equal(it.cpu().numpy().astype(np.uint32), want["iters"])
    assert np.array_equal(fl.cpu().numpy().astype(bool), want["max_iter_exit"])
    assert np.array_equal(lam.cpu().numpy(), want["lam"])
    assert want["iters"].min() <= 5 and want["max_iter_exit"].any()      # from almost-converged to capped


@pytest.mark.parametrize("cap", [8, 1])
def test_fallback_when_cluster_size_cannot_be_placed(torch_cuda, capi, oracle_pcg, cap):
    """A device that cannot place the default 16-CTA cluster (GBD_PCG_MAX_CLUSTER simulates it) falls back to the next
    variant of the shape -- cap 1 leaves only the grid kernel -- with the same bits."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import numpy as np, torch, sys\n"
        f"sys.path.insert(0, {root!r})\n"
        "import mpcgpu_b200 as m\n"
        "from mpcgpu_b200 import synth\n"
        "d = synth.make_systems(14, 128, batch=1, seed=7)\n"
        "S, P, g = (torch.from_numpy(d[k][0]).cuda() for k in ('S', 'Pinv', 'gamma'))\n"
        "lam = torch.zeros(14 * 128, device='cuda')\n"
        "it = torch.zeros(1, dtype=torch.int32, device='cuda'); fl = torch.zeros(1, dtype=torch.uint8, device='cuda')\n"
        "m.pcg_launch(14, 128, S, P, g, lam, None, None, None, None, it, fl, 167, 1e-4)\n"
        "torch.cuda.synchronize()\n"
        "np.save(sys.argv[1], lam.cpu().numpy()); print(int(it.item()), int(fl.item()))\n")
    out = os.path.join(root, "tests", "_build", f"fallback_{cap}.npy")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    env = dict(os.environ, GBD_PCG_MAX_CLUSTER=str(cap), GBD_PCG_NUMERICS="exact")
    res = subprocess.run([sys.executable, "-c", code, out], capture_output=True, text=True, timeout=150, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    iters, flag = (int(x) for x in res.stdout.split()[-2:])
    d = synth.make_systems(14, 128, batch=1, seed=7)
    want = oracle_pcg.pcg(d["S"][0], d["Pinv"][0], d["gamma"][0], d["lambda0"][0], 14, 128, 167, 1e-4)
    assert iters == want["iters"] and bool(flag) == want["max_iter_exit"]
    assert np.array_equal(np.load(out), want["lam"])


def test_full_size_batch_properties(torch_cuda, capi, oracle_pcg):
    """BASELINE config 4 size (1024 x N=128): too big for the oracle in seconds, so check
    size-independent properties: every system's fp64 residual is small, iteration counts equal the
    oracle's on a sample, and solving twice is deterministic."""
    import mpcgpu_b200 as m
    torch = torch_cuda
    n, N, B, cap, tol = 14, 128, 1024, 167, 1e-4
    d = synth.make_systems(n, N, batch=B, seed=1234)
    S, P, g = (_dev(torch, d[k]) for k in ("S", "Pinv", "gamma"))
    out = []
    for _ in range(2):
        lam = _dev(torch, d["lambda0"])
        it = torch.zeros(B, dtype=torch.int32, device="cuda")
        fl = torch.zeros(B, dtype=torch.uint8, device="cuda")
        m.solve_batched(n, N, B, S, P, g, lam, it, fl, cap, tol)
        torch.cuda.synchronize()
        out.append((lam.cpu().numpy(), it.cpu().numpy(), fl.cpu().numpy()))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    lam, it, fl = out[0]
    for i in (0, 1, 511, 1023):
        w = oracle_pcg.pcg(d["S"][i], d["Pinv"][i], d["gamma"][i], d["lambda0"][i], n, N, cap, tol)
        assert w["iters"] == it[i] and np.array_equal(w["lam"], lam[i])
    conv = fl == 0
    assert conv.mean() > 0.9
    for i in np.flatnonzero(conv)[::37]:
        assert oracle_pcg.rel_residual(d["S"][i], d["gamma"][i], lam[i], n, N) < 5e-3


def test_host_buffer_plan_and_linsys_window(torch_cuda, capi, oracle_pcg):
    import mpcgpu_b200 as m
    torch = torch_cuda
    n, N, cap, tol = 14, 128, 167, 1e-4
    d = synth.make_systems(n, N, batch=3, seed=5)
    want = [oracle_pcg.pcg(d["S"][i], d["Pinv"][i], d["gamma"][i], d["lambda0"][i], n, N, cap, tol) for i in range(3)]
    # solvePCG(h_S, ...) mirror: host numpy in, lambda overwritten
    lam = d["lambda0"][0].copy()
    iters, flag = m.solvePCG(d["S"][0], d["Pinv"][0], d["gamma"][0], lam, n, N,
                             m.PcgConfig(pcg_exit_tol=tol, pcg_max_iter=cap), return_flag=True)
    assert iters == want[0]["iters"] and flag == want[0]["max_iter_exit"] and np.array_equal(lam, want[0]["lam"])
    # batched plan
    plan = m.HostPlan(n, N, batch=3)
    lamb = d["lambda0"].copy()
    its, fls = plan.solve(d["S"], d["Pinv"], d["gamma"], lamb, cap, tol)
    for i in range(3):
        assert its[i] == want[i]["iters"] and np.array_equal(lamb[i], want[i]["lam"])
    plan.close()
    # pinned host buffers: the zero-copy path (kernel reads/writes host memory itself)
    pin = {k: torch.from_numpy(d[k]).pin_memory() for k in ("S", "Pinv", "gamma")}
    lamp = torch.zeros(3, n * N).pin_memory()
    plan = m.HostPlan(n, N, batch=3)
    its, fls = plan.solve(pin["S"].numpy(), pin["Pinv"].numpy(), pin["gamma"].numpy(), lamp.numpy(), cap, tol)
    for i in range(3):
        assert its[i] == want[i]["iters"] and np.array_equal(lamp[i].numpy(), want[i]["lam"])
    plan.close()
    # the SQP linsys window
    S, P, g = (_dev(torch, d[k][1]) for k in ("S", "Pinv", "gamma"))
    lamd = _dev(torch, d["lambda0"][1])
    r = torch.zeros(n * N, device="cuda")
    p = torch.zeros(n * N, device="cuda")
    it = torch.zeros(1, dtype=torch.int32, device="cuda")
    fl = torch.zeros(1, dtype=torch.uint8, device="cuda")
    iters, flag, us = m.linsys_window(n, N, S, P, g, lamd, r, p, it, fl, cap, tol)
    assert iters == want[1]["iters"] and flag == want[1]["max_iter_exit"] and 0 < us < 1e6
    assert np.array_equal(lamd.cpu().numpy(), want[1]["lam"])
    # device overload of solvePCG
    lamd2 = _dev(torch, d["lambda0"][2])
    S, P, g = (_dev(torch, d[k][2]) for k in ("S", "Pinv", "gamma"))
    assert m.solvePCG_device(n, N, S, P, g, lamd2, r, p, None, None,
                             m.PcgConfig(pcg_exit_tol=tol, pcg_max_iter=cap)) == want[2]["iters"]


def test_unaligned_pointers_take_the_non_tma_path(torch_cuda, capi, oracle_pcg):
    import mpcgpu_b200 as m
    torch = torch_cuda
    n, N, cap, tol = 14, 32, 60, 1e-6
    d = synth.make_systems(n, N, seed=17)
    want = oracle_pcg.pcg(d["S"][0], d["Pinv"][0], d["gamma"][0], d["lambda0"][0], n, N, cap, tol)
    big = torch.zeros(3 * n * n * N + 1, device="cuda")
    S = big[1:]                                               # 4-byte aligned only
    S.copy_(_dev(torch, d["S"][0]))
    big2 = torch.zeros(3 * n * n * N + 1, device="cuda")
    P = big2[1:]
    P.copy_(_dev(torch, d["Pinv"][0]))
    g, lam = _dev(torch, d["gamma"][0]), _dev(torch, d["lambda0"][0])
    it = torch.zeros(1, dtype=torch.int32, device="cuda")
    fl = torch.zeros(1, dtype=torch.uint8, device="cuda")
    assert S.data_ptr() % 16 != 0
    m.pcg_launch(n, N, S, P, g, lam, None, None, None, None, it, fl, cap, tol)
    torch.cuda.synchronize()
    assert int(it.item()) == want["iters"] and np.array_equal(lam.cpu().numpy(), want["lam"])


def test_ab_against_unmodified_reference_kernel(torch_cuda, capi):
    """A/B on the GPU against the reference's own pcg<> (oracle/_ref, compiled from the reference
    headers for sm_100a and launched as include/pcg/sqp.cuh:230 does): bit-identical outputs."""
    from oracle import refgpu
    if not refgpu.available():
        pytest.skip("oracle/_ref/libref_gbdpcg.so not present (built only where /root/reference exists)")
    import mpcgpu_b200 as m
    torch = torch_cuda
    for (n, N, cap, tol) in ((2, 3, 50, 1e-10), (6, 12, 60, 1e-6), (14, 32, 173, 1e-6), (14, 128, 167, 1e-4),
                             (14, 128, 167, 1e-6), (14, 512, 67, 1e-5), (64, 256, 40, 1e-6)):     # last: BASELINE config 5
        d = synth.make_systems(n, N, seed=40 + N, nan_pads=True)
        S, P, g, l0 = (_dev(torch, d[k][0]) for k in ("S", "Pinv", "gamma", "lambda0"))
        ref = refgpu.solve(n, N, S, P, g, l0, cap, tol, block=128)
        got = _gpu_solve(torch, m, d, 0, cap, tol)
        ref_np = dict(lam=ref["lam"].cpu().numpy(), iters=ref["iters"], max_iter_exit=ref["max_iter_exit"],
                      r=ref["r"].cpu().numpy(), p=ref["p"].cpu().numpy())
        _assert_same(got, ref_np, f"vs reference kernel ({n},{N})")


def test_golden_reference_vectors_every_variant(torch_cuda, capi):
    """tests/golden/iiwa_*.npz: real IIWA systems assembled by the reference, answers from the reference
    kernel on a B200 (tools/make_golden.py).  Every compiled fp32 variant of that size must reproduce
    lambda, r, p, iters and the exit flag bit for bit -- no oracle involved."""
    import mpcgpu_b200 as m
    from test_golden import GOLDEN, load
    L = capi.lib()
    for path in GOLDEN:
        g = load(path)
        n, N = g["n"], g["N"]
        d = dict(n=n, N=N, S=g["S"][None], Pinv=g["Pinv"][None], gamma=g["gamma"][None],
                 lambda0=np.zeros((1, n * N), np.float32))
        vs = [v for v in capi.variants() if v["n"] == n and v["N"] == N and not v["f64"] and not v["fast"]]
        assert vs
        for v in vs:
            assert L.gbd_pcg_set_tuning(n, N, 0, v["cluster"], v["mode"]) == 0
            try:
                for run in g["runs"]:
                    got = _gpu_solve(torch_cuda, m, d, 0, run["cap"], run["tol"])
                    _assert_same(got, run, f"{g['name']} tol {run['tol']} variant {v}")
            finally:
                L.gbd_pcg_set_tuning(n, N, 0, 0, -1)


def test_double_precision_instantiation(torch_cuda, capi, oracle_pcg):
    """USE_DOUBLES=1 equivalent (include/common/settings.cuh:41-49)."""
    import mpcgpu_b200 as m
    n, N, cap, tol = 14, 128, 167, 1e-10
    d = synth.make_systems(n, N, seed=3, dtype=np.float64)
    got = _gpu_solve(torch_cuda, m, d, 0, cap, tol, np.float64)
    want = oracle_pcg.pcg(d["S"][0], d["Pinv"][0], d["gamma"][0], d["lambda0"][0], n, N, cap, tol)
    _assert_same(got, want, "f64")


@pytest.mark.parametrize("knots,block", [(32, 128), (128, 128), (32, 256)])
def test_dropin_headers_on_golden_reference_vectors(torch_cuda, tmp_path, knots, block):
    """Drop-in pcg<float,14,N> under the reference's launch geometry on the reference-minted golden systems."""
    import os
    import subprocess
    from test_golden import GOLDEN, load
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", f"dropin_demo_{knots}")
    if not os.path.exists(exe):
        pytest.skip("tests/_build/dropin_demo_* not built (run __graft_entry__.build())")
    for path in GOLDEN:
        g = load(path)
        if g["N"] != knots:
            continue
        n = g["n"]
        vec = n * knots
        fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
        np.concatenate([g["S"], g["Pinv"], g["gamma"], np.zeros(vec, np.float32)]).astype(np.float32).tofile(fin)
        for run in g["runs"]:
            subprocess.check_call([exe, str(fin), str(fout), str(run["cap"]), repr(run["tol"]), str(block), "2"], timeout=120)
            raw = np.fromfile(fout, np.float32)
            tail = raw[3 * vec:].view(np.uint32)
            got = dict(lam=raw[:vec], r=raw[vec:2 * vec], p=raw[2 * vec:3 * vec], iters=int(tail[0]), max_iter_exit=bool(tail[1]))
            _assert_same(got, run, f"drop-in {g['name']} block={block} tol={run['tol']}")


@pytest.mark.parametrize("knots,block", [(32, 128), (32, 64), (64, 128), (64, 256), (64, 64), (128, 128), (128, 64), (32, 256), (256, 128),
                                         (512, 128), (512, 64)])
def test_dropin_headers_reference_launch_geometry(torch_cuda, oracle_pcg, tmp_path, knots, block):
    """include/gbd_dropin: pcg<float,14,N> launched exactly like include/pcg/sqp.cuh:230 (cooperative,
    grid = N, block = PCG_NUM_THREADS, smem = pcgSharedMemSize) -- bit-exact vs the oracle."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", f"dropin_demo_{knots}")
    if not os.path.exists(exe):
        pytest.skip("tests/_build/dropin_demo_* not built (run __graft_entry__.build())")
    n, cap, tol = 14, (167 if knots <= 128 else 67), 1e-5
    d = synth.make_systems(n, knots, seed=9, nan_pads=True)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    np.concatenate([d["S"][0], d["Pinv"][0], d["gamma"][0], d["lambda0"][0]]).astype(np.float32).tofile(fin)
    subprocess.check_call([exe, str(fin), str(fout), str(cap), repr(tol), str(block), "3"], timeout=120)
    raw = np.fromfile(fout, np.float32)
    vec = n * knots
    tail = raw[3 * vec:].view(np.uint32)
    got = dict(lam=raw[:vec], r=raw[vec:2 * vec], p=raw[2 * vec:3 * vec], iters=int(tail[0]), max_iter_exit=bool(tail[1]))
    want = oracle_pcg.pcg(d["S"][0], d["Pinv"][0], d["gamma"][0], d["lambda0"][0], n, knots, cap, tol)
    _assert_same(got, want, f"drop-in N={knots} block={block}")
