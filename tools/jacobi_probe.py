"""Diagnostics: Jacobi sweep counts per order during the first iterations of the bench workload."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from xframe_b200.plan import HIO
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
plan, sd, rho0 = bench.build_problem(nb, 0, [1000 + i for i in range(nb)])
plan.mtip_init(rho0)
for it in range(4):
    plan.mtip_iterate(HIO, True, [0.5])
    torch.cuda.synchronize()
    orders, sw = plan.jacobi_sweeps()
    print('iter', it, 'orders', orders[:6], '...', 'sweeps run0:', sw[0].tolist())
    print('   mean sweeps', sw.mean(), 'max', sw.max())
