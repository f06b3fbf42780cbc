"""CPU tests: the C-ABI library builds, loads, exports every symbol include/gbd_pcg.h declares,
and refuses to compute without a GPU (no fallback).  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    txt = open(os.path.join(ROOT, "include", "gbd_pcg.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gbd_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_all_exported(capi):
    declared = _declared_functions()
    assert len(declared) >= 14
    assert sorted(capi.SYMBOLS) == declared
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH], text=True)
    exported = set(re.findall(r" T (gbd_\w+)", out))
    assert set(declared) <= exported
    L = capi.lib()
    for s in declared:
        assert getattr(L, s) is not None


def test_abi_version_and_variants(capi):
    assert capi.lib().gbd_pcg_abi_version() == 2
    vs = capi.variants()
    pairs = {(v["n"], v["N"]) for v in vs if not v["f64"]}
    # IIWA horizons of the reference (include/common/settings.cuh:123-138) + the GBD-PCG demo size
    assert {(14, 32), (14, 64), (14, 128), (14, 256), (14, 512), (2, 3)} <= pairs
    for v in vs:
        assert v["N"] % v["cluster"] == 0 and 1 <= v["cluster"] <= (148 if v["mode"] in (4, 24, 25) else 16)
        assert v["threads"] % 32 == 0 and v["threads"] <= 1024
        assert v["smem"] <= 227 * 1024
    L = capi.lib()
    assert L.gbd_pcg_supported(14, 128, 0) == 1 and L.gbd_pcg_supported(14, 128, 1) == 1
    assert L.gbd_pcg_supported(13, 128, 0) == 0
    assert L.gbd_pcg_strerror(-1).decode().startswith("no kernel")


def test_tuning_knob(capi):
    L = capi.lib()
    assert L.gbd_pcg_set_tuning(14, 128, 0, 16, -1) == capi.OK
    assert L.gbd_pcg_set_tuning(14, 128, 0, 5, -1) == capi.ERR_UNSUPPORTED
    assert L.gbd_pcg_set_tuning(14, 128, 0, 0, -1) == capi.OK


def test_schur_team_knob_returns_the_previous_mode(capi):
    """gbd_schur_set_team: -1 = by size (default), 0 = CTA per block row, 1 = warp per block row; any negative value means -1."""
    L = capi.lib()
    assert L.gbd_schur_set_team(1) == -1
    assert L.gbd_schur_set_team(0) == 1
    assert L.gbd_schur_set_team(-7) == 0
    assert L.gbd_schur_set_team(-1) == -1


def test_bad_arguments_are_errors_not_aborts(capi):
    L = capi.lib()
    it, fl = C.c_uint32(), C.c_uint8()
    # null pointers
    assert L.gbd_pcg_solve_f32(14, 128, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 10, 1e-6, 0) == capi.ERR_BADARG
    # unsupported size (pointers non-null but never dereferenced)
    assert L.gbd_pcg_solve_f32(13, 128, 8, 8, 8, 8, 0, 0, 0, 0, 8, 8, 10, 1e-6, 0) == capi.ERR_UNSUPPORTED
    assert L.gbd_pcg_plan_create(13, 128, 1, 0, C.byref(C.c_void_p())) == capi.ERR_UNSUPPORTED
    assert L.gbd_pcg_plan_create(14, 128, 0, 0, C.byref(C.c_void_p())) == capi.ERR_BADARG
    assert L.gbd_pcg_plan_destroy(None) == capi.OK


def test_no_gpu_means_error_not_fallback(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import mpcgpu_b200 as m
    with pytest.raises(capi.GbdPcgError) as ei:
        m.HostPlan(14, 128)
    assert ei.value.status in (capi.ERR_NODEVICE, capi.ERR_CUDA)
    d = m.synth.make_systems(2, 3)
    with pytest.raises(capi.GbdPcgError):
        m.solvePCG(d["S"][0], d["Pinv"][0], d["gamma"][0], d["lambda0"][0].copy(), 2, 3, m.PcgConfig())


def test_product_does_not_import_oracle():
    """The product path may not route through the oracle (tier rule 3)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mpcgpu_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "oracle/" not in txt and "pcg_oracle" not in txt, f
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "oracle" not in open(p).read(), f


def test_synth_layout_and_roofline_bytes():
    from mpcgpu_b200 import synth
    # B_iter(14,32)=158144, (14,128)=641984, (14,512)=2577344, (64,256)=25493504
    assert [synth.bytes_per_iteration(*a) for a in ((14, 32), (14, 128), (14, 512), (64, 256))] == \
        [158144, 641984, 2577344, 25493504]
    d = synth.make_systems(6, 12, batch=3, seed=1, nan_pads=True)
    T = d["S"].reshape(3, 12, 3, 6, 6)
    assert np.isnan(T[:, 0, 0]).all() and np.isnan(T[:, 11, 2]).all()
    assert np.isfinite(T[:, 1:, 0]).all() and np.isfinite(T[:, :, 1]).all() and np.isfinite(T[:, :11, 2]).all()
    # symmetry of the assembled matrix: right tile of row b == (left tile of row b+1)^T
    np.testing.assert_array_equal(T[:, :-1, 2], np.swapaxes(T[:, 1:, 0], -1, -2))
    # negated (negative definite) storage convention
    assert (np.diagonal(T[:, :, 1], axis1=-2, axis2=-1) < 0).all()


def _pcg_instantiations(binary):
    out = subprocess.run(["cuobjdump", "-res-usage", binary], capture_output=True, text=True).stdout
    return sorted(set(re.findall(r"Function (_Z3pcgI\w+?EEv)", out)))


def test_dropin_build_instantiates_the_reference_shapes():
    """Regression for an include-order bug: the reference's headers include "gpuassert.cuh" / "utils.cuh"
    BEFORE include/common/settings.cuh defines STATE_SIZE, so a drop-in file that defines more than its
    namesake (e.g. the STATE_SIZE default of constants.cuh) silently turns sqp.cuh's
    pcg<T,STATE_SIZE,KNOT_POINTS> into pcg<float,3,N>.  The reference example built against the drop-in
    headers must contain exactly the pcg<> instantiations of the build against GBD-PCG/include."""
    run = os.path.join(ROOT, "oracle", "_ref", "run")
    pairs = [(os.path.join(run, f"track_iiwa_pcg_ref_{k}"), os.path.join(run, f"track_iiwa_pcg_dropin_{k}")) for k in (32, 128)]
    pairs = [p for p in pairs if os.path.exists(p[0]) and os.path.exists(p[1])]
    if not pairs:
        pytest.skip("oracle/_ref/run/track_iiwa_pcg_* not built (make -C oracle examples; needs /root/reference)")
    for ref, drop in pairs:
        want = _pcg_instantiations(ref)
        assert want and all("Lj14E" in w for w in want), want
        assert _pcg_instantiations(drop) == want


def test_dropin_headers_define_what_their_namesakes_define():
    """gpuassert.cuh must not leak STATE_SIZE / KNOT_POINTS; constants.cuh must (as defaults only)."""
    inc = os.path.join(ROOT, "include", "gbd_dropin")
    src = "#include \"%s\"\n#if defined(STATE_SIZE) || defined(KNOT_POINTS)\n#error leaked\n#endif\nint main(){return 0;}\n"
    for hdr, leaks in (("gpuassert.cuh", False), ("constants.cuh", True), ("types.cuh", True), ("utils.cuh", True)):
        p = subprocess.run(["/usr/local/cuda/bin/nvcc", "-std=c++17", "-E", "-x", "cu", "-I" + inc, "-I" + os.path.join(ROOT, "include"),
                            "-"], input=src % hdr, capture_output=True, text=True)
        assert (p.returncode != 0) == leaks, (hdr, p.stderr[-300:])
