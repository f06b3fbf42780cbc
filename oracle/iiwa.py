"""Real IIWA Schur systems minted by the REFERENCE's own assembly -- TEST / BENCHMARK-INPUT INFRASTRUCTURE ONLY.

`oracle/_ref/ref_capture_<N>` (oracle/ref_capture.cu compiled against the reference headers where they lie) runs the reference's
generate_kkt_submatrices + form_schur_system (include/pcg/sqp.cuh:94-219, include/pcg/linsys_setup.cuh:565-657) on the reference
trajectory examples/trajfiles/0_0_* at rho = 1e-3 and dumps (S, Pinv, gamma).  Two families:

* ``ring(N, count, stride)``  system i = the N-knot window at knot offset i * stride of 0_0_traj.csv: what successive MPC steps of
  examples/track_iiwa_pcg.cu hand to the solver at their first SQP iteration (BASELINE.json configs 1-3).
* ``perturbed(N, count)``     system i = the window at offset 0 plus N(0, 0.05^2) on q, N(0, 0.01^2) on qd, N(0, 1) on u,
  std::mt19937_64(1234 + i): BASELINE.json configs[3], the 1024-trajectory batch.

Needs a GPU (the reference's assembly kernels run on it) and the binaries built by `make -C oracle capture` (which needs
/root/reference; the built binaries travel to the GPU box).  The product never imports this module; bench.py uses it only to
obtain INPUTS (and says so in its `data` field), tests use it for parity on real data.
"""
from __future__ import annotations

import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(_HERE, "_ref")
N_STATE = 14
TRAJ_ROWS = 666          # knots in examples/trajfiles/0_0_traj.csv


def available(N: int) -> bool:
    return all(os.path.exists(os.path.join(REF, f)) for f in (f"ref_capture_{N}", "0_0_traj.csv", "0_0_eepos.traj"))


def _capture(N: int, offset: int, count: int, mode: int, stride: int = 1, timeout: int = 900):
    n = N_STATE
    mat, vec = 3 * n * n * N, n * N
    with tempfile.TemporaryDirectory(prefix="gbd_iiwa_") as td:
        raw = os.path.join(td, "cap.bin")
        subprocess.check_call([os.path.join(REF, f"ref_capture_{N}"), os.path.join(REF, "0_0_traj.csv"),
                               os.path.join(REF, "0_0_eepos.traj"), raw, str(offset), str(count), str(mode), "1", str(stride)],
                              timeout=timeout, stdout=subprocess.DEVNULL)
        a = np.fromfile(raw, np.float32)
    assert a.size == count * (2 * mat + vec), (a.size, count, mat, vec)
    a = a.reshape(count, 2 * mat + vec)
    return dict(n=n, N=N, S=np.ascontiguousarray(a[:, :mat]), Pinv=np.ascontiguousarray(a[:, mat:2 * mat]),
                gamma=np.ascontiguousarray(a[:, 2 * mat:]), lambda0=np.zeros((count, vec), np.float32))


def max_ring(N: int, stride: int) -> int:
    return (TRAJ_ROWS - N) // stride + 1


def ring(N: int, count: int, stride: int = 2):
    """`count` distinct systems from sliding windows of the reference trajectory (pad tiles zeroed)."""
    count = min(count, max_ring(N, stride))
    d = _capture(N, 0, count, 2, stride)
    d["source"] = (f"reference generate_kkt_submatrices + form_schur_system (rho 1e-3) on examples/trajfiles/0_0, {count} windows of "
                   f"{N} knots at offsets 0, {stride}, ..")
    return d


def perturbed(N: int, count: int):
    """BASELINE.json configs[3]: `count` randomly perturbed copies of the first window."""
    d = _capture(N, 0, count, 1)
    d["source"] = (f"reference generate_kkt_submatrices + form_schur_system (rho 1e-3) on the first {N}-knot window of "
                   f"examples/trajfiles/0_0 + N(0,0.05^2) q, N(0,0.01^2) qd, N(0,1) u, mt19937_64(1234+i), {count} trajectories")
    return d
