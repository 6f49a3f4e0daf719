"""Harmonic-transform round-trip microbench (BASELINE.json configs[1], SURVEY.md section 8d "Config 2"):
batched sh.inverse(sh.forward(x)) and ift(ft(x)) for L in {31,63,127,255}, N_r in {64,128,256,512},
n_theta = L+1 (rounded up to a multiple of 8), n_phi = 2(L+1); x = band-limited synthesis of N(0,1)+iN(0,1)
coefficients, numpy.random.default_rng(1234); batch chosen to fill ~2 GiB of grid data.

Prints one JSON line per (L, N_r): round-trip error (relative L2, the parity property: analysis o synthesis = id for
band-limited input, shtns_plugin.py:263-267), time per round trip (CUDA events), achieved HBM GB/s against the
algorithmic bytes 2*(G + C)*16 per SHT pair, and Hankel FP64 TFLOP/s inside the FT.
    python tools/transform_bench.py [--quick] > gpurun_out/transform_bench.jsonl
"""
import argparse, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xframe_b200.plan import Plan


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    except Exception:
        return 6650.0, 'fallback'


def bench(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def run(L, n_r, target_bytes, reps):
    n_theta = ((L + 1 + 7) // 8) * 8
    n_phi = 2 * (L + 1)
    G = n_r * n_theta * n_phi
    C = n_r * (L + 1) ** 2
    nb = max(1, min(256, int(target_bytes // (G * 16))))
    plan = Plan(L, n_r, 0.3, n_theta=n_theta, n_phi=n_phi, max_batch=nb)
    rng = np.random.default_rng(1234)
    c0 = torch.from_numpy(rng.standard_normal((n_r, (L + 1) ** 2)) + 1j * rng.standard_normal((n_r, (L + 1) ** 2))).cuda()
    c = c0[None].repeat(nb, 1, 1).contiguous()
    x = plan.sht_inverse(c)                                   # band-limited grid data [nb, N_r, n_theta, n_phi]
    back = plan.sht_forward(x)
    err_sht = float((back - c).norm() / c.norm())
    y = plan.ift(plan.ft(x))
    # ift(ft(x)) is not the identity for the discrete Hankel pair; the parity property used here is linearity + SHT exactness
    lin = plan.ft((2.0 * x).contiguous())
    err_lin = float((lin - 2.0 * plan.ft(x)).norm() / lin.norm())
    ms_sht = bench(lambda: plan.sht_inverse(plan.sht_forward(x)), reps)
    ms_ft = bench(lambda: plan.ift(plan.ft(x)), reps)
    peak, src = peaks()
    bytes_pair = nb * 2 * (G + C) * 16                         # forward: read G write C; inverse: read C write G
    out = {'L': L, 'n_r': n_r, 'n_theta': n_theta, 'n_phi': n_phi, 'batch': nb, 'grid_GiB': nb * G * 16 / 2 ** 30,
           'sht_roundtrip_rel_l2': err_sht, 'ft_linearity_rel_l2': err_lin,
           'sht_pair_ms': ms_sht, 'sht_pair_GBps': bytes_pair / ms_sht / 1e6, 'sht_pair_frac_of_hbm_peak': bytes_pair / ms_sht / 1e6 / peak,
           'ft_ift_pair_ms': ms_ft, 'sht_pairs_per_s': nb * n_r / (ms_sht * 1e-3), 'peak_GBps': peak, 'peak_source': src,
           'finite': bool(torch.isfinite(torch.view_as_real(y)).all())}
    plan.close()
    del x, y, c, back, lin
    torch.cuda.empty_cache()
    return out


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--gib', type=float, default=2.0)
    ap.add_argument('--reps', type=int, default=5)
    a = ap.parse_args()
    cases = [(31, 64), (63, 128)] if a.quick else [(31, 64), (63, 128), (127, 256), (255, 512)]
    for L, n_r in cases:
        print(json.dumps(run(L, n_r, a.gib * 2 ** 30, a.reps)), flush=True)
