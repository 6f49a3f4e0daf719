"""Full tutorial reconstructions on the GPU (600 iterations + 6 SW per run) for a list of seeds, through the public
worker API; prints one JSON object with per-run summaries (same fields as tools/oracle_full_run.py) and the measured
reconstructions/hour including init, shrink-wrap steps and result read-back.
    python tools/gpu_full_run.py --runs 16 [--seed0 1000]
"""
import argparse, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xframe_b200 import setup_host as S
from xframe_b200.plan import Plan
from xframe_b200.settings import tutorial_settings
from xframe_b200.worker import ProjectWorker

L, NR, NT, NP, MAXQ = 63, 128, 64, 128, 0.322416

ap = argparse.ArgumentParser()
ap.add_argument('--runs', type=int, default=16)
ap.add_argument('--seed0', type=int, default=1000)
ap.add_argument('--out', default=None)
args = ap.parse_args()

sd = tutorial_settings(grid={'max_q': MAXQ, 'max_order': L, 'n_phi': NP, 'n_theta': NT, 'n_radial_points': NR})
sd['GPU'] = {'use': True, 'batch': args.runs, 'seed': args.seed0}
boot = Plan(L, NR, MAXQ, n_theta=NT, n_phi=NP, max_batch=1)
data = S.invariants_from_density(boot, S.six_sphere_density(boot))
boot.close()
w = ProjectWorker(sd, data, n_reconstructions=args.runs)
torch.cuda.synchronize()
t0 = time.time()
res, _ = w.run()
torch.cuda.synchronize()
dt = time.time() - t0
Bref = [p @ p.conj().T for p in w.proj.projection_matrices]
runs = []
for r in res:
    errs = r['error_dict']['main']
    Bl = r['last_deg2_invariant']
    inv_err = [float(np.sum(np.abs(Bref[l] - Bl[l]) ** 2) / max(np.sum(np.abs(Bref[l]) ** 2), 1e-300)) for l in range(0, L + 1, 2)]
    runs.append({'seed': args.seed0 + r['run_id'], 'final_error': float(r['final_error']), 'last_error': float(errs[-1]), 'n_errors': len(errs),
                 'errors_every_20': [float(e) for e in errs[::20]], 'support_fraction': float(r['last_support_mask'].mean()),
                 'deg2_invariant_l2_diff_even_orders': inv_err, 'finite': bool(np.isfinite(r['real_density']).all())})
out = {'runs': runs, 'seconds_total': dt, 'reconstructions_per_hour': args.runs / dt * 3600.0, 'n_runs': args.runs,
       'launches': w.plan.launch_count()}
s = json.dumps(out)
if args.out:
    open(args.out, 'w').write(s)
print(json.dumps({k: v for k, v in out.items() if k != 'runs'}))
for r in runs[:4]:
    print(r['seed'], 'final', r['final_error'], 'last', r['last_error'], 'support', r['support_fraction'], 'inv_err l=2,10,30:',
          r['deg2_invariant_l2_diff_even_orders'][1], r['deg2_invariant_l2_diff_even_orders'][5], r['deg2_invariant_l2_diff_even_orders'][15])
