"""CPU model of the warm-started Procrustes step: runs the numpy oracle for a few MTIP iterations at the bench
configuration, records G_l = M_l^T of every iteration and reports, per iteration,
  * how far G_l V_prev is from having orthogonal columns (max scaled off-diagonal of its Gram matrix),
  * the number of cyclic one-sided Jacobi sweeps needed from that start (tol 1e-15) against a cold start.
Usage: python tools/jacobi_warm_model.py [n_iter] [L] [N_r]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from jacobi_model import to_real  # noqa: E402
from oracle import mtip as O  # noqa: E402
from xframe_b200.settings import tutorial_settings  # noqa: E402


def sweeps_to_converge(G, tol=1e-15, eps=1e-15, max_sweeps=40):
    """vectorised cyclic one-sided Jacobi (row-cyclic order, numpy): returns sweeps, rotations, rotated G, J."""
    G = G.copy()
    n = G.shape[1]
    J = np.eye(n)
    rots = 0
    for sweep in range(1, max_sweeps + 1):
        n2 = (G * G).sum(0)
        thr = eps * eps * n2.max()
        act = [c for c in range(n) if n2[c] > thr]
        rot = 0
        for ii, p in enumerate(act):
            for q in act[ii + 1:]:
                a, b = G[:, p], G[:, q]
                app, aqq, apq = a @ a, b @ b, a @ b
                if app <= thr or aqq <= thr or abs(apq) <= tol * np.sqrt(app * aqq):
                    continue
                zeta = (aqq - app) / (2 * apq)
                t = (1.0 if zeta >= 0 else -1.0) / (abs(zeta) + np.sqrt(1 + zeta * zeta))
                c = 1 / np.sqrt(1 + t * t)
                s = c * t
                G[:, p], G[:, q] = c * a - s * b, s * a + c * b
                ja, jb = J[:, p].copy(), J[:, q].copy()
                J[:, p], J[:, q] = c * ja - s * jb, s * ja + c * jb
                rot += 1
        rots += rot
        if rot == 0:
            return sweep, rots, G, J
    return max_sweeps, rots, G, J


def main():
    n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 63
    n_r = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    n_theta, n_phi = (L + 1 + 7) // 8 * 8, 2 * (L + 1)
    max_q = 0.322416
    qs = O.radial_grids('midpoint', max_q, n_r, 2.0)[1]
    sd = tutorial_settings(grid={'max_q': max_q, 'max_order': L, 'n_phi': n_phi, 'n_theta': n_theta, 'n_radial_points': n_r})
    boot = O.MTIP(sd, {'data_radial_points': qs, 'average_intensity': np.ones(n_r), 'max_order': L,
                       'data_projection_matrices': [np.zeros((n_r, min(n_r, 2 * l + 1)), complex) for l in range(L + 1)]})
    data = O.invariants_from_density(O.six_sphere_density(boot.real_grid), boot.ft, boot.sh, boot.qs)
    m = O.MTIP(sd, data)
    m.results['errors'] = {'real': {'l2_projection_diff': []}, 'reciprocal': {}, 'main': []}
    rec = []
    orig = m.rp.approximate_unknowns

    def hook(I):
        rec.append([to_real(np.asarray(I[l])) for l in range(L + 1)])
        return orig(I)
    m.rp.approximate_unknowns = hook
    rho = m.density_guess(np.random.default_rng(1000))
    rho = m.ift(m.ft(rho))
    m.beta = 0.5
    t0 = time.time()
    for it in range(n_iter):
        rho = m.io_step('HIO', rho, True)[1]
    print(f'{n_iter} oracle iterations in {time.time() - t0:.1f}s', flush=True)
    orders = [l for l in range(2, L + 1, 2 if L > 15 else 1) if np.abs(m.rp.projection_matrices[l]).max() > 0]
    pick = orders[::max(1, len(orders) // 6)]
    print('orders modelled:', pick)
    Vprev = {}
    for it in range(n_iter):
        line = []
        for l in pick:
            V = m.rp.projection_matrices[l].real
            Mt = ((V.T * qs[None, :] ** 2) @ rec[it][l]).T          # G = M^T  [2l+1, n_l]
            n2 = (Mt * Mt).sum(0)
            act = np.nonzero(n2 > 1e-30 * n2.max())[0]
            G = Mt[:, act]
            cold_s, cold_r, _, Jc = sweeps_to_converge(G)
            if l in Vprev and Vprev[l].shape[0] == G.shape[1]:
                Gw = G @ Vprev[l]
                gram = Gw.T @ Gw
                d = np.sqrt(np.abs(np.diag(gram)))
                d = np.where(d > 0, d, 1.0)
                off = np.abs(gram / d[:, None] / d[None, :] - np.eye(len(d))).max()
                warm_s, warm_r, _, Jw = sweeps_to_converge(Gw)
                Vprev[l] = Vprev[l] @ Jw
                line.append(f'l={l}: r={G.shape[1]} cold {cold_s}sw/{cold_r}rot warm {warm_s}sw/{warm_r}rot off {off:.1e}')
            else:
                Vprev[l] = Jc
                line.append(f'l={l}: r={G.shape[1]} cold {cold_s}sw/{cold_r}rot')
        print(f'it {it}: ' + ' | '.join(line), flush=True)


if __name__ == '__main__':
    main()
