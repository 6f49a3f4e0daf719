"""Diagnostic sweep on a GPU box: every C-ABI operator against the oracle, printing relative-L2 errors.
Not a test (tests/ holds the asserting versions) -- used while bringing kernels up."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import load_golden, golden_settings, golden_data, rel_l2  # noqa: E402
from oracle import mtip as O  # noqa: E402
from xframe_b200.plan import Plan, HIO, ER  # noqa: E402

dev = torch.device('cuda')


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def N(t):
    return t.cpu().numpy()


def report(name, got, ref, tol):
    e = rel_l2(got, ref)
    flag = 'ok ' if e < tol else 'BAD'
    print(f'  [{flag}] {name:34s} rel_l2={e:.3e} (tol {tol:g})', flush=True)
    return e < tol


def run_case(tag):
    print(f'== {tag}', flush=True)
    g = load_golden(tag)
    sd = golden_settings(g)
    m = O.MTIP(sd, golden_data(g))
    L, n_r = m.l_max, len(m.rs)
    plan = Plan(L, n_r, float(g['max_q']), n_theta=int(g['n_theta']), n_phi=int(g['n_phi']), max_batch=3)
    ok = True
    x = g['x_grid']
    c = plan.sht_forward(T(x))
    ok &= report('sht_forward', N(c), g['sht_forward_direct'], 1e-12)
    band = plan.sht_inverse(T(g['sht_forward_direct']))
    ok &= report('sht_inverse', N(band), g['sht_inverse_of_forward'], 1e-12)
    hk = plan.hankel(T(g['hankel_in_direct'])[None])
    ok &= report('hankel fwd', N(hk)[0], g['hankel_fwd_direct'], 1e-12)
    hk = plan.hankel(T(g['hankel_in_direct'])[None], inverse=True)
    ok &= report('hankel inv', N(hk)[0], g['hankel_inv_direct'], 1e-12)
    b3 = np.stack([g['sht_inverse_of_forward']] * 3)
    b3[1] *= 2.0
    f = plan.ft(T(b3))
    ok &= report('ft (batch 3, run 0)', N(f)[0], g['ft_x'], 1e-11)
    ok &= report('ft (batch 3, run 1)', N(f)[1], 2 * g['ft_x'], 1e-11)
    f = plan.ift(T(b3))
    ok &= report('ift (batch 3, run 2)', N(f)[2], g['ift_x'], 1e-11)
    # projection
    plan.set_projection(m.rp.projection_matrices, m.rp.radial_mask, m.rp.number_of_particles[0])
    ip = plan.project_invariants(T(np.stack([g['I_direct'], g['I_direct']])))
    ok &= report('project_invariants', N(ip)[1], g['Iproj_direct'], 1e-6)
    splits = np.arange(1, L + 1) ** 2
    for l, (a, b) in enumerate(zip(np.split(N(ip)[0], splits, axis=1), np.split(g['Iproj_direct'], splits, axis=1))):
        nb = np.linalg.norm(b)
        print(f'      l={l:3d} rel={np.linalg.norm(a - b) / (nb if nb > 0 else 1):.2e} |ref|={nb:.2e}')
    mi = plan.modify_intensity(T(g['rho_hat0'])[None], T(g['I_proj_grid'])[None])
    ok &= report('modify_intensity', N(mi)[0], g['rho_hat_mod'], 1e-14)
    # real side
    popt = sd['projections']['real']['projections']
    plan.set_real(popt['apply'], m.real_pr.initial_support, popt['value_threshold']['threshold'], popt['limit_imag']['threshold'])
    sup = torch.ones((1,) + plan.grid_shape, dtype=torch.uint8, device=dev)
    nxt, err = plan.real_update(HIO, float(g['hio_beta']), T(g['rho_new'])[None], T(g['rho0'])[None], sup)
    ok &= report('real_update HIO', N(nxt)[0], g['hio_out'], 1e-14)
    e = N(err)[0]
    print(f'      err {e[0] / e[1]:.15e} ref {float(g["real_err"]):.15e}')
    ok &= abs(e[0] / e[1] - float(g['real_err'])) < 1e-11 * float(g['real_err'])
    nxt, err = plan.real_update(ER, 0.0, T(g['rho_new'])[None], T(g['rho0'])[None], sup)
    ok &= report('real_update ER', N(nxt)[0], g['rho_proj'], 1e-14)
    swm = plan.shrinkwrap(T(g['rho0'])[None], 12.5, 0.09)
    d = (N(swm)[0] != g['sw_mask']).mean()
    print(f'  [{"ok " if d < 1e-3 else "BAD"}] shrinkwrap mask mismatch fraction {d:.2e}')
    ok &= d < 1e-3
    # loop
    if 'loop_main_error' in g:
        loops = sd['main_loop']['sub_loops']
        res = run_loop(plan, m, sd, g['rho0'])
        n = len(g['loop_main_error'])
        print('      gpu  errors', np.array2string(res['errors'][0][:n], precision=6))
        print('      ref  errors', np.array2string(g['loop_main_error'], precision=6))
        ok &= report('loop error history', res['errors'][0], g['loop_main_error'], 1e-6)
        ok &= report('loop last_real_density', res['last_real'][0], g['loop_last_real_density'], 1e-6)
        ok &= report('loop best real_density', res['best_real'][0], g['loop_real_density'], 1e-6)
        ok &= report('loop last_reciprocal', res['last_reciprocal'][0], g['loop_last_reciprocal_density'], 1e-6)
        d = (res['last_support'][0] != g['loop_last_support_mask']).mean()
        d2 = (res['best_support'][0] != g['loop_support_mask']).mean()
        print(f'      support mismatch last {d:.2e} best {d2:.2e}')
        ok &= d < 1e-3 and d2 < 1e-3
    plan.close()
    return ok


def run_loop(plan, m, sd, rho0, nb=1):
    """Host driver mirroring reconstruct.py:854-951 on the device-resident batch (same as xframe_b200.reconstruct)."""
    from xframe_b200.reconstruct import run_schedule
    rho = T(np.stack([rho0] * nb))
    return run_schedule(plan, sd, rho, default_sigma=np.pi / plan.qs.max())


if __name__ == '__main__':
    allok = True
    for tag in ['ref_small_ftstab', 'ref_small_plain', 'ref_medium_ops']:
        try:
            allok &= run_case(tag)
        except Exception:
            traceback.print_exc()
            allok = False
    print('ALL OK' if allok else 'FAILURES')
    sys.exit(0 if allok else 1)
