"""Check: results of a run do not depend on the batch it is in (nb = 1 vs nb = 32), and every Jacobi problem is solved."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from xframe_b200.plan import HIO
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 32
plan, sd, rho0 = bench.build_problem(nb, 0, [1000 + i for i in range(nb)])
plan.mtip_init(rho0)
for _ in range(4):
    plan.mtip_iterate(HIO, True, [0.5])
torch.cuda.synchronize()
orders, sw = plan.jacobi_sweeps()
print('zero-sweep problems:', int((sw == 0).sum()), 'of', sw.size, ' sweeps min/mean/max', sw.min(), sw.mean(), sw.max())
hist, best = plan.mtip_errors()
big = plan.mtip_grid('last_real').cpu().numpy()
plan.close()
for k in (0, 5, nb - 1):
    p1, _, r1 = bench.build_problem(1, 0, [1000 + k])
    p1.mtip_init(r1)
    for _ in range(4):
        p1.mtip_iterate(HIO, True, [0.5])
    one = p1.mtip_grid('last_real').cpu().numpy()[0]
    h1, _ = p1.mtip_errors()
    print('run', k, 'rel diff nb=1 vs batch', np.linalg.norm(one - big[k]) / np.linalg.norm(one), 'errors', hist[k].tolist()[-1], h1[0].tolist()[-1])
    p1.close()
