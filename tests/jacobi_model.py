"""numpy model of the device Procrustes algorithm (csrc/procrustes.cuh): real-basis change + one-sided Jacobi with
round-robin ordering and singular-value cut-off.  Used by tests to show the ALGORITHM agrees with the reference's
SVD-based formula independent of the GPU."""
import numpy as np


def to_real(Il):
    l = (Il.shape[1] - 1) // 2
    X = np.empty(Il.shape)
    X[:, 0] = Il[:, l].real
    for m in range(1, l + 1):
        X[:, 2 * m - 1] = np.sqrt(2) * Il[:, l + m].real
        X[:, 2 * m] = np.sqrt(2) * Il[:, l + m].imag
    return X


def to_cplx(X):
    l = (X.shape[1] - 1) // 2
    I = np.empty(X.shape, complex)
    I[:, l] = X[:, 0]
    for m in range(1, l + 1):
        c = (X[:, 2 * m - 1] + 1j * X[:, 2 * m]) / np.sqrt(2)
        I[:, l + m] = c
        I[:, l - m] = (-1) ** m * np.conj(c)
    return I


def jacobi_project(V, qs, Il, eps=1e-15, tol=1e-15, max_sweeps=40):
    """T = V polar(V^T D^2 X) in the real basis, returned as complex coefficients [N_r, 2l+1]; also sweeps used."""
    V = np.asarray(V).real
    X = to_real(Il)
    M = (V.T * qs[None, :] ** 2) @ X
    G, P = M.T.copy(), V.copy()
    n = G.shape[1]
    sweeps = 0
    for sweeps in range(1, max_sweeps + 1):
        nrm2 = (G * G).sum(0)
        thr = eps * eps * nrm2.max()
        lst = [c for c in range(n) if nrm2[c] > thr]
        nact = len(lst)
        if nact < 2:
            break
        npad = nact + (nact & 1)
        rot = False
        for r in range(npad - 1):
            for i in range(npad // 2):
                ka, kb = i, npad - 1 - i
                pa = 0 if ka == 0 else 1 + ((ka - 1 - r) % (npad - 1))
                pb = 1 + ((kb - 1 - r) % (npad - 1))
                if pa >= nact or pb >= nact:
                    continue
                p, q = sorted((lst[pa], lst[pb]))
                a, b = G[:, p].copy(), G[:, q].copy()
                app, aqq, apq = a @ a, b @ b, a @ b
                if app <= thr or aqq <= thr or abs(apq) <= tol * np.sqrt(app * aqq):
                    continue
                zeta = (aqq - app) / (2 * apq)
                t = (1.0 if zeta >= 0 else -1.0) / (abs(zeta) + np.sqrt(1 + zeta * zeta))
                c = 1 / np.sqrt(1 + t * t)
                s = c * t
                G[:, p], G[:, q] = c * a - s * b, s * a + c * b
                pa_, pb_ = P[:, p].copy(), P[:, q].copy()
                P[:, p], P[:, q] = c * pa_ - s * pb_, s * pa_ + c * pb_
                rot = True
        if not rot:
            break
    nrm2 = (G * G).sum(0)
    keep = (nrm2 > eps * eps * nrm2.max()) & (nrm2 > 0)
    Gn = np.where(keep[None, :], G / np.sqrt(np.where(keep, nrm2, 1.0))[None, :], 0.0)
    return to_cplx(P @ Gn.T), sweeps


def mgs2(A, selective=True):
    """Column modified Gram-Schmidt with re-orthogonalisation (csrc/procrustes.cuh:mgs2_qr): Q, R.
    selective: the second projection pass runs only for groups of 4 columns (one warp) in which the first pass removed more
    than half of a column's squared norm, or whose downdated cached norm fell below 1e-6 of the last fresh value -- the
    device rule; selective=False is the always-twice variant."""
    A = A.copy()
    m, r = A.shape
    R, Q = np.zeros((r, r)), np.zeros((m, r))
    n2 = (A * A).sum(0)
    nref = n2.copy()
    for j in range(r):
        nrm = np.sqrt(A[:, j] @ A[:, j])
        q = A[:, j] / nrm if nrm > 0 else np.zeros(m)
        Q[:, j], R[j, j] = q, nrm
        if j + 1 == r:
            break
        t = slice(j + 1, r)
        c = q @ A[:, t]
        A[:, t] -= np.outer(q, c)
        R[j, t] += c
        n2n = n2[t] - c * c
        need = (c * c > 0.5 * n2[t]) | (n2n < 1e-6 * nref[t]) if selective else np.ones(r - j - 1, bool)
        for k0 in range(0, r - j - 1, 4):                  # warp granularity
            if need[k0:k0 + 4].any():
                need[k0:k0 + 4] = True
        idx = np.nonzero(need)[0] + j + 1
        if len(idx):
            c2 = q @ A[:, idx]
            A[:, idx] -= np.outer(q, c2)
            R[j, idx] += c2
        n2[t] = n2n
        fresh = np.union1d(idx, np.arange(j + 1, min(r, j + 5)))
        n2[fresh] = (A[:, fresh] * A[:, fresh]).sum(0)
        nref[fresh] = n2[fresh]
    return Q, R


def qr_polar(V, qs, Il, eps=1e-15, tol=1e-15, max_sweeps=40):
    """numpy model of the QR-preconditioned device algorithm (csrc/procrustes.cuh:jacobi_qr_problem):
    G_a = Q1 R1, R1^T = Q2 R2, Jacobi on L = R2^T with W = Q2 rotated along, polar(G_a) = Q1 U~ W^T.
    Returns T = V polar(V^T D^2 X) as complex coefficients [N_r, 2l+1] and the number of sweeps."""
    V = np.asarray(V).real
    M = (V.T * qs[None, :] ** 2) @ to_real(Il)
    G = M.T.copy()
    nrm = (G * G).sum(0)
    act = np.nonzero(nrm > eps * eps * nrm.max())[0]
    Q1, R1 = mgs2(G[:, act])
    Q2, R2 = mgs2(R1.T.copy())
    L, W = R2.T.copy(), Q2.copy()
    r = L.shape[1]
    sweeps = 0
    for sweeps in range(1, max_sweeps + 1):
        n2 = (L * L).sum(0)
        thr = eps * eps * n2.max()
        lst = [c for c in range(r) if n2[c] > thr]
        rot = False
        for i in range(len(lst)):
            for j in range(i + 1, len(lst)):
                p, q = lst[i], lst[j]
                a, b = L[:, p].copy(), L[:, q].copy()
                app, aqq, apq = a @ a, b @ b, a @ b
                if app <= thr or aqq <= thr or abs(apq) <= tol * np.sqrt(app * aqq):
                    continue
                zeta = (aqq - app) / (2 * apq)
                t = (1.0 if zeta >= 0 else -1.0) / (abs(zeta) + np.sqrt(1 + zeta * zeta))
                c = 1 / np.sqrt(1 + t * t)
                s = c * t
                L[:, p], L[:, q] = c * a - s * b, s * a + c * b
                wa, wb = W[:, p].copy(), W[:, q].copy()
                W[:, p], W[:, q] = c * wa - s * wb, s * wa + c * wb
                rot = True
        if not rot:
            break
    n2 = (L * L).sum(0)
    keep = (n2 > eps * eps * n2.max()) & (n2 > 0)
    U = np.where(keep[None, :], L / np.sqrt(np.where(keep, n2, 1.0))[None, :], 0.0)
    P = U @ W.T                                  # polar(R1)
    Tt = Q1 @ (P @ V.T[act])                     # [2l+1, N_r]
    return to_cplx(Tt.T), sweeps
