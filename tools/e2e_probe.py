"""Diagnostics: e2e (host-buffer) step time vs pipeline chunk size, and device-resident step time vs batch size."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from xframe_b200.plan import HIO
nb = 128
plan, sd, rho0 = bench.build_problem(nb, 0, [1000 + i for i in range(nb)])
plan.mtip_init(rho0)
for _ in range(3):
    plan.mtip_iterate(HIO, True, [0.5])
torch.cuda.synchronize()
h_in = torch.empty((nb,) + plan.grid_shape, dtype=torch.complex128).pin_memory()
h_out = torch.empty_like(h_in).pin_memory()
h_err = torch.empty((nb, 2), dtype=torch.float64).pin_memory()
h_in.copy_(plan.mtip_grid('last_real').cpu())
for chunk in (8, 16, 24, 32, 64):
    plan.set_host_chunk(chunk)
    plan.mtip_step_host(HIO, True, 0.5, h_in, h_out, h_err)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        plan.mtip_step_host(HIO, True, 0.5, h_in, h_out, h_err)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f'chunk {chunk}: {dt * 1e3:.1f} ms/step  {nb / dt:.0f} it/s', flush=True)
# pure copies
t0 = time.perf_counter(); d = h_in.cuda(non_blocking=True); torch.cuda.synchronize(); print('H2D 2GiB ms', (time.perf_counter() - t0) * 1e3)
t0 = time.perf_counter(); h_out.copy_(d, non_blocking=True); torch.cuda.synchronize(); print('D2H 2GiB ms', (time.perf_counter() - t0) * 1e3)
del d
plan.close()
for n in (8, 16, 32, 64):
    plan, sd, rho0 = bench.build_problem(n, 0, [1000 + i for i in range(n)])
    plan.mtip_init(rho0)
    for _ in range(3):
        plan.mtip_iterate(HIO, True, [0.5])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    plan.profile(True)
    a.record()
    for _ in range(5):
        plan.mtip_iterate(HIO, True, [0.5])
    b.record(); torch.cuda.synchronize()
    pr = plan.profile_read()
    print(f'device nb={n}: {a.elapsed_time(b) / 5:.2f} ms/step = {a.elapsed_time(b) / 5 / n:.3f} ms/run; jacobi {pr["procrustes_jacobi"]["ms"] / 5:.2f} ms', flush=True)
    plan.close()
