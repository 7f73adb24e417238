"""Lane-level model of the bordered register Cholesky of the wide kernel build (csrc/sim_core.cuh `spd_solve<36>`, CUDA branch:
the host emulation uses the plain shared-memory routine, so this is the CPU check of the algorithm itself).  32 "lanes" own
row i of the leading 32 x 32 block plus their entries of the KB border rows; the KB x KB corner is replicated; shuffles become
array reads of another lane's value.  Compared with numpy's dense solve for nv = 33 (hammer) ... 36 (relocate)."""
import numpy as np
import pytest


def spd_solve_lanes(A, x, dadd=None, hh=0.0, NA=32, KB=4):
    nv = A.shape[0]
    h = np.zeros((32, NA))                       # h[i][j]: lane i, register j
    for i in range(32):
        for j in range(NA):
            h[i, j] = A[i, j] if (i < nv and j < nv) else (1.0 if i == j else 0.0)
    cb = np.zeros((32, KB)); S = np.eye(KB); bb = np.zeros(KB)
    for r in range(KB):
        row = 32 + r
        if row < nv:
            cb[:min(nv, 32), r] = A[row, :min(nv, 32)]
            for q in range(KB):
                if 32 + q < nv:
                    S[r, q] = A[row, 32 + q]
            bb[r] = x[row]
    if dadd is not None:
        for i in range(min(nv, 32)):
            h[i, i] += hh * dadd[i]
        for r in range(KB):
            if 32 + r < nv:
                S[r, r] += hh * dadd[32 + r]
    b = np.array([x[i] if i < nv else 0.0 for i in range(32)])
    dinv = np.ones(32)
    for k in range(NA):
        inv = 1.0 / np.sqrt(max(h[k, k], 1e-30))            # shfl(h[k], k)
        lik = np.where(np.arange(32) > k, h[:, k] * inv, 0.0)
        dinv[k] = inv
        h[np.arange(32) > k, k] = lik[np.arange(32) > k]
        for j in range(k + 1, NA):
            h[:, j] -= lik * lik[j]                          # ljk = value of lane j
        lr = cb[k, :] * inv                                  # shfl(cb[r] * inv, k)
        for r in range(KB):
            new = cb[:, r] - lr[r] * lik
            new[k] = lr[r]
            cb[:, r] = new
        S -= np.outer(lr, lr)
    sinv = np.zeros(KB)
    for k in range(KB):
        sinv[k] = 1.0 / np.sqrt(max(S[k, k], 1e-30))
        for r in range(k + 1, KB):
            S[r, k] *= sinv[k]
        for r in range(k + 1, KB):
            for q in range(k + 1, r + 1):
                S[r, q] -= S[r, k] * S[q, k]
    for k in range(NA):                                      # L y = b
        yk = b[k] * dinv[k]
        bn = b - h[:, k] * yk
        b = np.where(np.arange(32) > k, bn, np.where(np.arange(32) == k, yk, b))
    for r in range(KB):
        bb[r] -= np.sum(cb[:, r] * b)                        # warp reduction
    for k in range(KB):
        bb[k] *= sinv[k]
        for r in range(k + 1, KB):
            bb[r] -= S[r, k] * bb[k]
    for k in range(KB - 1, -1, -1):
        bb[k] *= sinv[k]
        for r in range(k):
            bb[r] -= S[k, r] * bb[k]
    for r in range(KB):
        b = b - cb[:, r] * bb[r]
    sacc = np.zeros(32); z = np.zeros(32)
    for k in range(NA - 1, -1, -1):                          # L^T z = y
        zk = (b[k] - dinv[k] * sacc[k]) * dinv[k]
        sn = sacc + h[:, k] * zk                             # lane i < k holds h_i[k] = H-row entry frozen as L_ki / dinv
        sacc = np.where(np.arange(32) < k, sn, sacc)
        z[k] = zk
    out = np.array(x, dtype=float)
    out[:min(nv, 32)] = z[:min(nv, 32)]
    for r in range(KB):
        if 32 + r < nv:
            out[32 + r] = bb[r]
    return out


@pytest.mark.parametrize("nv", [33, 34, 36])
@pytest.mark.parametrize("damped", [False, True])
def test_bordered_register_cholesky_matches_dense_solve(nv, damped):
    rng = np.random.default_rng(nv)
    G = rng.normal(size=(nv, nv + 8))
    A = G @ G.T + 0.5 * np.eye(nv)
    x = rng.normal(size=nv)
    d = rng.uniform(0, 3, nv) if damped else None
    got = spd_solve_lanes(A, x, d, 0.002)
    want = np.linalg.solve(A + (0.002 * np.diag(d) if damped else 0), x)
    assert np.allclose(got, want, rtol=1e-9, atol=1e-10)
