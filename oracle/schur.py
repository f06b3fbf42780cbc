"""ctypes binding of oracle/_build/libschur_oracle.so (oracle/schur_oracle.c): CPU restatement of the reference's
form_schur_system (include/pcg/linsys_setup.cuh:621-657) and compute_dz (include/common/dz.cuh:125-136).
TEST INFRASTRUCTURE ONLY -- imported by tests/ and tools/, never by mpcgpu_b200."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libschur_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "schur_oracle.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
        L = C.CDLL(LIB)
        fp = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
        L.schur_oracle_form_f32.restype = C.c_int
        L.schur_oracle_form_f32.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, fp, fp, fp, fp, fp, fp, fp, C.c_float]
        L.schur_oracle_dz_f32.restype = C.c_int
        L.schur_oracle_dz_f32.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, fp, fp, fp, fp, fp]
        _lib = L
    return _lib


def sizes(n: int, m: int, N: int):
    return dict(G=(n * n + m * m) * (N - 1) + n * n, C=(n * n + n * m) * (N - 1), g=(n + m) * (N - 1) + n, c=n * N,
                S=3 * n * n * N, gamma=n * N)


def form(G, Cm, g, c, n, m, N, rho, pad=np.nan):
    """Returns dict(S, Pinv, gamma, Ginv).  Pad tiles (left of row 0, right of row N-1) are filled with `pad`:
    the reference never writes them."""
    sz = sizes(n, m, N)
    Ginv = np.ascontiguousarray(G, np.float32).copy()
    S = np.full(sz["S"], pad, np.float32)
    P = np.full(sz["S"], pad, np.float32)
    gamma = np.zeros(sz["gamma"], np.float32)
    rc = lib().schur_oracle_form_f32(n, m, N, Ginv, np.ascontiguousarray(Cm, np.float32), np.ascontiguousarray(g, np.float32),
                                     np.ascontiguousarray(c, np.float32), S, P, gamma, np.float32(rho))
    assert rc == 0
    return dict(S=S, Pinv=P, gamma=gamma, Ginv=Ginv)


def dz(Ginv, Cm, g, lam, n, m, N):
    out = np.zeros((n + m) * (N - 1) + n, np.float32)
    rc = lib().schur_oracle_dz_f32(n, m, N, np.ascontiguousarray(Ginv, np.float32), np.ascontiguousarray(Cm, np.float32),
                                   np.ascontiguousarray(g, np.float32), np.ascontiguousarray(lam, np.float32), out)
    assert rc == 0
    return out


def make_kkt(n: int, m: int, N: int, seed: int = 0):
    """Seeded synthetic KKT blocks in the reference's dense layouts: SPD Q_k, R_k, A_k = I + small, B_k small."""
    rng = np.random.default_rng(seed)
    G, Cm, g = [], [], []
    for k in range(N):
        M = rng.standard_normal((n, n))
        Q = M @ M.T / n + np.eye(n)
        G.append(Q.T.ravel())                                     # column-major
        g.append(rng.standard_normal(n))
        if k < N - 1:
            Mr = rng.standard_normal((m, m))
            G.append((Mr @ Mr.T / m + np.eye(m)).T.ravel())
            Cm.append((np.eye(n) + rng.standard_normal((n, n)) / 16).T.ravel())
            Cm.append((rng.standard_normal((n, m)) / 16).T.ravel())
            g.append(rng.standard_normal(m))
    c = rng.standard_normal(n * N) * 0.1
    return (np.concatenate(G).astype(np.float32), np.concatenate(Cm).astype(np.float32), np.concatenate(g).astype(np.float32),
            c.astype(np.float32))


def form_f64(G, Cm, g, c, n, m, N, rho):
    """Independent fp64 statement of the same mathematics (numpy inverses / matmuls) for sanity checks of the oracle."""
    nn, mm, nm = n * n, m * m, n * m
    G, Cm, g, c = (np.asarray(x, np.float64) for x in (G, Cm, g, c))
    Q = [G[k * (nn + mm):k * (nn + mm) + nn].reshape(n, n).T + rho * np.eye(n) for k in range(N)]
    R = [G[k * (nn + mm) + nn:(k + 1) * (nn + mm)].reshape(m, m).T + rho * np.eye(m) for k in range(N - 1)]
    A = [Cm[k * (nn + nm):k * (nn + nm) + nn].reshape(n, n).T for k in range(N - 1)]
    B = [Cm[k * (nn + nm) + nn:(k + 1) * (nn + nm)].reshape(m, n).T for k in range(N - 1)]
    q = [g[k * (n + m):k * (n + m) + n] for k in range(N)]
    r = [g[k * (n + m) + n:(k + 1) * (n + m)] for k in range(N - 1)]
    S = np.full((N, 3, n, n), np.nan)
    P = np.full((N, 3, n, n), np.nan)
    gam = np.zeros((N, n))
    Qi = [np.linalg.inv(x) for x in Q]
    Ri = [np.linalg.inv(x) for x in R]
    theta = [Qi[0]]
    S[0, 1] = -Qi[0].T
    P[0, 1] = -Q[0].T
    gam[0] = -Qi[0] @ q[0]
    phis = [None]
    for b in range(1, N):
        phi = A[b - 1] @ Qi[b - 1]
        BR = B[b - 1] @ Ri[b - 1]
        th = phi @ A[b - 1].T + BR @ B[b - 1].T + Qi[b]
        gam[b] = -(Qi[b] @ q[b] - c[b * n:(b + 1) * n] + phi @ q[b - 1] + BR @ r[b - 1])
        S[b, 0] = -phi.T
        S[b, 1] = -th.T
        S[b - 1, 2] = -phi                                      # (phi^T) stored column-major = phi row-major view
        P[b, 1] = -np.linalg.inv(th).T
        theta.append(th)
        phis.append(phi)
    Td = [P[b, 1].T for b in range(N)]                          # stored diagonal tiles as matrices
    for b in range(N):
        if b:
            P[b, 0] = -(Td[b] @ (S[b, 0].T) @ Td[b - 1]).T
        if b < N - 1:
            P[b, 2] = -(Td[b] @ (S[b + 1, 0].T).T @ Td[b + 1]).T
    return dict(S=S.reshape(-1), Pinv=P.reshape(-1), gamma=gam.reshape(-1))
