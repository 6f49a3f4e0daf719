"""Diagnostics: where the cycles of the QR-preconditioned Jacobi kernel go (library built with -DJAC_TIMING, selected with XFB200_LIB).
usage (GPU box): XFB200_LIB=$PWD/xframe_b200/lib/libxfb200_jtiming.so python tools/jacobi_phase_probe.py [runs]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from xframe_b200 import _lib  # noqa: E402
from xframe_b200.plan import HIO  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 128
plan, sd, rho0 = bench.build_problem(nb, 0, [1000 + i for i in range(nb)])
plan.mtip_init(rho0)
plan.mtip_iterate(HIO, True, [0.5] * 3)
torch.cuda.synchronize()
buf = (C.c_double * 8)()
_lib.check(plan.lib.xfb_debug_jacobi_phase_cycles(buf))            # reset
for split in (1,):
    plan.mtip_iterate(HIO, True, [0.5] * 4)
    _lib.check(plan.lib.xfb_debug_jacobi_phase_cycles(buf))
    v = list(buf)
    tot = sum(v[:5])
    names = ['active list + gather', 'first QR', 'second QR', 'sweeps', 'final norms + polar product']
    print(f'problems {v[5]:.0f}, sweeps per problem {v[6] / max(v[5], 1):.2f}, cycles per problem {tot / max(v[5], 1):.0f}')
    for n, c in zip(names, v[:5]):
        print(f'  {n:30s} {100 * c / tot:5.1f} %   {c / max(v[5], 1):10.0f} cycles per problem')
orders, sw = plan.jacobi_sweeps()
print('sweeps by order (run 0):', dict(zip(orders, sw[0].tolist())))
