"""Full tutorial reconstructions (600 iterations + 6 SW, L=63, N_r=128) with the numpy oracle for a few seeds.
Writes tests/golden/full_run_oracle.json: per-seed error history summary + final metrics.  Run in the build container:
    OMP_NUM_THREADS=1 python tools/oracle_full_run.py 1000 1001 1002
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mtip as O
from xframe_b200.settings import tutorial_settings
import multiprocessing as mp

L, NR, NT, NP, MAXQ = 63, 128, 64, 128, 0.322416


def setup():
    sd = tutorial_settings(grid={'max_q': MAXQ, 'max_order': L, 'n_phi': NP, 'n_theta': NT, 'n_radial_points': NR})
    qs = O.radial_grids('midpoint', MAXQ, NR, 2.0)[1]
    boot = O.MTIP(sd, {'data_radial_points': qs, 'average_intensity': np.ones(NR), 'max_order': L,
                       'data_projection_matrices': [np.zeros((NR, min(NR, 2 * l + 1)), complex) for l in range(L + 1)]})
    data = O.invariants_from_density(O.six_sphere_density(boot.real_grid), boot.ft, boot.sh, boot.qs)
    return sd, data


def run(seed):
    sd, data = setup()
    m = O.MTIP(sd, data)
    t0 = time.time()
    res = m.run(rng=np.random.default_rng(seed))
    errs = res['error_dict']['main']
    # invariant error of the last density (fxs_IO_methods.py:432-446 with all-true mask, N_particles = 1)
    Bl, Bref = res['last_deg2_invariant'], m.rp.deg2_invariants
    inv_err = [float(np.sum(np.abs(Bref[l] - Bl[l]) ** 2) / max(np.sum(np.abs(Bref[l]) ** 2), 1e-300)) for l in range(0, L + 1, 2)]
    out = {'seed': seed, 'final_error': float(res['final_error']), 'last_error': float(errs[-1]), 'n_errors': len(errs),
           'errors_every_20': [float(e) for e in errs[::20]], 'support_fraction': float(res['last_support_mask'].mean()),
           'deg2_invariant_l2_diff_even_orders': inv_err, 'seconds': time.time() - t0}
    print(json.dumps(out)[:300], flush=True)
    return out


if __name__ == '__main__':
    seeds = [int(s) for s in sys.argv[1:]] or [1000]
    with mp.get_context('fork').Pool(len(seeds)) as pool:
        outs = pool.map(run, seeds)
    path = os.path.join(ROOT, 'tests', 'golden', 'full_run_oracle.json')
    with open(path, 'w') as f:
        json.dump({'config': 'tutorial schedule 5x(60 HIO,SW,40 ER)+1x(SW,100 ER), L=63, N_r=128, 64x128, six-sphere invariants',
                   'runs': outs}, f, indent=1)
    print('wrote', path)
