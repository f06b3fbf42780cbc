"""Seeded synthetic Schur-complement systems in the reference's storage convention.

The reference builds ``S``, ``Pinv`` and ``gamma`` on the GPU from the KKT blocks of the MPC problem
(``include/pcg/linsys_setup.cuh:139-657``).  This module produces systems with the same structure,
sign convention and storage layout from random LQR-like data, so the solver can be exercised without
the robot dynamics:

* dynamics  ``x_{k+1} = A_k x_k + B_k u_k``,  ``A_k = I + a_scale * N(0,1)``,  ``B_k = a_scale * N(0,1)``
* costs     ``Q_k = M M^T / n + (1 + rho) I``,  ``R_k`` likewise (m x m)
* Schur complement of the KKT system ``S = C G^{-1} C^T`` (block tridiagonal, SPD)::

      D_0 = Q_0^{-1}                D_k = A Q^{-1} A^T + B R^{-1} B^T + Q_k^{-1}   (k >= 1, A,B,Q,R of step k-1)
      L_k = -A_{k-1} Q_{k-1}^{-1}   (block (k, k-1));   block (k, k+1) = L_{k+1}^T

* STORED matrix is ``-S`` (the reference stores every tile with multiplier -1,
  ``include/pcg/linsys_setup.cuh:202-210,249-255,491-524``), so ``eta = r . Pinv r <= 0``.
* STORED preconditioner is the symmetric-stair one (``linsys_setup.cuh:9-137``):
  ``Pinv_kk = (S_kk)^{-1}``, ``Pinv_{k,k-1} = -Pinv_kk S_{k,k-1} Pinv_{k-1,k-1}``,
  ``Pinv_{k,k+1} = -Pinv_kk S_{k,k+1} Pinv_{k+1,k+1}``  (all in terms of the stored ``S``).
* layout ``[N][3][n][n]``: tile t in {0: left, 1: diag, 2: right}, column-major inside a tile
  (``GBD-PCG/include/utils.cuh:58-84,98-161``); pad tiles (0,0) and (N-1,2) are filled with NaN on
  request to prove nobody reads them.

Everything is computed in fp64 and cast once to the requested dtype.
"""
from __future__ import annotations

import numpy as np


def stair_preconditioner(S: np.ndarray, n: int, N: int) -> np.ndarray:
    """Pinv in band layout from a stored band matrix S (any leading batch dims), in fp64."""
    T = np.asarray(S, np.float64).reshape(-1, N, 3, n, n)
    # tiles are column-major: T[..., c, r] = block[r, c]  ->  block = swapaxes
    L = np.swapaxes(T[:, :, 0], -1, -2)
    D = np.swapaxes(T[:, :, 1], -1, -2)
    Rr = np.swapaxes(T[:, :, 2], -1, -2)
    Dinv = np.linalg.inv(D)
    P = np.zeros_like(T)
    P[:, :, 1] = np.swapaxes(Dinv, -1, -2)
    PL = -(Dinv[:, 1:] @ L[:, 1:] @ Dinv[:, :-1])
    PR = -(Dinv[:, :-1] @ Rr[:, :-1] @ Dinv[:, 1:])
    P[:, 1:, 0] = np.swapaxes(PL, -1, -2)
    P[:, :-1, 2] = np.swapaxes(PR, -1, -2)
    return P.reshape(S.shape)


def make_systems(n: int, N: int, batch: int = 1, seed: int = 0, m: int | None = None, rho: float = 1e-3,
                 a_scale: float = 1.0 / 64.0, dtype=np.float32, nan_pads: bool = False, chunk: int = 32):
    """Return dict(S, Pinv, gamma, lambda0) with shapes [batch, N*3*n*n] x2, [batch, N*n] x2."""
    if m is None:
        m = max(1, n // 2)
    rng = np.random.default_rng(seed)
    S_out = np.empty((batch, N * 3 * n * n), dtype)
    P_out = np.empty_like(S_out)
    g_out = np.empty((batch, N * n), dtype)
    eye_n, eye_m = np.eye(n), np.eye(m)
    for b0 in range(0, batch, chunk):
        B = min(chunk, batch - b0)
        M = rng.standard_normal((B, N, n, n))
        Q = M @ np.swapaxes(M, -1, -2) / n + (1.0 + rho) * eye_n
        Mr = rng.standard_normal((B, N, m, m))
        Rm = Mr @ np.swapaxes(Mr, -1, -2) / m + (1.0 + rho) * eye_m
        A = eye_n + a_scale * rng.standard_normal((B, N, n, n))
        Bm = a_scale * rng.standard_normal((B, N, n, m))
        Qi = np.linalg.inv(Q)
        Ri = np.linalg.inv(Rm)
        AQi = A @ Qi                                                     # A_k Q_k^{-1}
        D = np.empty((B, N, n, n))
        D[:, 0] = Qi[:, 0]
        D[:, 1:] = (AQi[:, :-1] @ np.swapaxes(A[:, :-1], -1, -2)
                    + Bm[:, :-1] @ Ri[:, :-1] @ np.swapaxes(Bm[:, :-1], -1, -2) + Qi[:, 1:])
        D = 0.5 * (D + np.swapaxes(D, -1, -2))
        Lb = np.zeros((B, N, n, n))
        Lb[:, 1:] = -AQi[:, :-1]
        T = np.zeros((B, N, 3, n, n))
        # stored = -S ; tiles column-major -> store block^T
        T[:, :, 1] = -np.swapaxes(D, -1, -2)
        T[:, 1:, 0] = -np.swapaxes(Lb[:, 1:], -1, -2)
        T[:, :-1, 2] = -Lb[:, 1:]                                        # (L_{k+1}^T)^T = L_{k+1}
        Tq = T.astype(dtype)                                             # what the solver will see
        P = stair_preconditioner(Tq.reshape(B, -1), n, N).reshape(B, N, 3, n, n)
        Pq = P.astype(dtype)
        if nan_pads:
            Tq[:, 0, 0] = np.nan
            Tq[:, N - 1, 2] = np.nan
            Pq[:, 0, 0] = np.nan
            Pq[:, N - 1, 2] = np.nan
        S_out[b0:b0 + B] = Tq.reshape(B, -1)
        P_out[b0:b0 + B] = Pq.reshape(B, -1)
        g_out[b0:b0 + B] = rng.standard_normal((B, N * n)).astype(dtype)
    return dict(S=S_out, Pinv=P_out, gamma=g_out, lambda0=np.zeros((batch, N * n), dtype), n=n, N=N, batch=batch)


def bytes_per_iteration(n: int, N: int, elem: int = 4) -> int:
    """Algorithmic bytes of one PCG iteration of one system:
    both band matrices (real tiles only) read once + r, p, lambda each read and written once."""
    return elem * (2 * (3 * N - 2) * n * n + 6 * N * n)


def flops_per_iteration(n: int, N: int) -> int:
    return 2 * 2 * (3 * N - 2) * n * n + 10 * N * n


def make_kkt_batch(n: int, m: int, N: int, batch: int = 1, seed: int = 0, chunk: int = 64):
    """Seeded synthetic KKT blocks for `batch` trajectories in the reference's dense layouts (the inputs of form_schur_system,
    include/pcg/linsys_setup.cuh:621-657): G = per knot [Q_k (n*n) | R_k (m*m)] (last knot Q only), C = per knot
    [A_k (n*n) | B_k (n*m)], g = per knot [q_k | r_k], c = n per knot; blocks column-major.  SPD Q, R; A = I + small; B small.
    Returns float32 arrays G [batch, .], C [batch, .], g [batch, .], c [batch, .]."""
    rng = np.random.default_rng(seed)
    nn, mm, nm = n * n, m * m, n * m
    G = np.empty((batch, (nn + mm) * (N - 1) + nn), np.float32)
    C = np.empty((batch, (nn + nm) * (N - 1)), np.float32)
    g = np.empty((batch, (n + m) * (N - 1) + n), np.float32)
    c = np.empty((batch, n * N), np.float32)
    for b0 in range(0, batch, chunk):
        B = min(chunk, batch - b0)
        M = rng.standard_normal((B, N, n, n))
        Q = M @ np.swapaxes(M, -1, -2) / n + np.eye(n)
        Mr = rng.standard_normal((B, N - 1, m, m))
        R = Mr @ np.swapaxes(Mr, -1, -2) / m + np.eye(m)
        A = np.eye(n) + rng.standard_normal((B, N - 1, n, n)) / 16
        Bm = rng.standard_normal((B, N - 1, n, m)) / 16
        q = rng.standard_normal((B, N, n))
        r = rng.standard_normal((B, N - 1, m))
        Gk = np.concatenate([np.swapaxes(Q[:, :-1], -1, -2).reshape(B, N - 1, nn), np.swapaxes(R, -1, -2).reshape(B, N - 1, mm)], axis=2)
        G[b0:b0 + B] = np.concatenate([Gk.reshape(B, -1), np.swapaxes(Q[:, -1], -1, -2).reshape(B, nn)], axis=1)
        C[b0:b0 + B] = np.concatenate([np.swapaxes(A, -1, -2).reshape(B, N - 1, nn), np.swapaxes(Bm, -1, -2).reshape(B, N - 1, nm)],
                                      axis=2).reshape(B, -1)
        g[b0:b0 + B] = np.concatenate([np.concatenate([q[:, :-1], r], axis=2).reshape(B, -1), q[:, -1]], axis=1)
        c[b0:b0 + B] = 0.1 * rng.standard_normal((B, n * N))
    return G, C, g, c
