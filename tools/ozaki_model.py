"""CPU model of an FP64-equivalent Hankel contraction on the tcgen05 int8 tensor path (Ozaki scheme, error-free slicing):
accuracy on the real Hankel weights and a cost model against the measured FP64 DMMA kernel.

  rows  A [R x N_r]  (real / imaginary parts of the coefficient rows of one order l), per-ROW power-of-two scale
  W_l   [N_r x N_r]  (hankel_transforms.py:399-410), per-COLUMN power-of-two scale
  slices of `bits` bits (int8 operands: 6 value bits + sign keeps every partial sum exact in the int32 accumulator:
  2*6 + log2(128) + log2(#pairs of one weight) <= 31), products A_i W_j kept for i + j < n_slices (the triangle), summed in FP64.

Usage: python tools/ozaki_model.py [L] [N_r]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xframe_b200 import tables  # noqa: E402


def slices(x, axis, bits, n):
    """error-free split of x into n integer slices of `bits` bits after scaling every vector along `axis` by a power of two."""
    mx = np.abs(x).max(axis=axis, keepdims=True)
    e = np.where(mx > 0, np.ceil(np.log2(np.where(mx > 0, mx, 1.0))), 0.0)
    r = x / 2.0 ** e                      # |r| <= 1
    out = []
    for _ in range(n):
        r = r * 2.0 ** bits
        q = np.trunc(r)
        out.append(q)
        r = r - q
    return out, e


def ozaki_matmul(A, W, bits=6, n=9, full=False):
    a_s, ea = slices(A, 1, bits, n)
    w_s, ew = slices(W, 0, bits, n)
    acc = np.zeros((A.shape[0], W.shape[1]))
    n_prod = 0
    for s in range(2 * n - 1 if full else n):          # group by weight 2^{-bits (i + j + 2)}
        g = np.zeros_like(acc)
        for i in range(n):
            j = s - i
            if 0 <= j < n:
                g += a_s[i] @ w_s[j]                   # exact: integers below 2^31
                n_prod += 1
        acc += g * 2.0 ** (-bits * (s + 2))
    return acc * 2.0 ** ea * 2.0 ** ew, n_prod


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 63
    n_r = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    W = tables.hankel_weights(L, n_r, 2.0, 'midpoint')            # [L+1][p][k]
    rng = np.random.default_rng(0)
    rs, _ = tables.radial_grids('midpoint', 0.322416, n_r, 2.0)
    print(f'L={L} N_r={n_r}; weights dynamic range per order: '
          + ', '.join(f'l={l}: {np.abs(W[l][W[l] != 0]).min():.1e}..{np.abs(W[l]).max():.1e}' for l in (0, L // 2, L)))
    for kind in ('white', 'physical'):
        worst = {}
        for l in (0, 8, 16, 32, 48, L):
            R = 2 * (2 * l + 1)
            A = rng.normal(size=(R, n_r))
            if kind == 'physical':                         # coefficients of a compact density behave like r^l near the origin
                A = A * (rs / rs.max())[None, :] ** l
            exact = (A.astype(np.longdouble) @ W[l].astype(np.longdouble))
            ref64 = A @ W[l]
            nrm = np.linalg.norm(exact.astype(float))
            e64 = np.linalg.norm((ref64 - exact).astype(float)) / nrm
            line = f'{kind:8s} l={l:3d}: fp64 {e64:.1e}'
            for n in (7, 8, 9, 10):
                oz, n_prod = ozaki_matmul(A, W[l], 6, n)
                eo = np.linalg.norm((oz - exact).astype(float)) / nrm
                rowerr = (np.linalg.norm((oz - exact).astype(float), axis=1) / np.maximum(np.linalg.norm(exact.astype(float), axis=1), 1e-300)).max()
                line += f' | {n} slices ({n_prod} int8 GEMMs): {eo:.1e} (worst row {rowerr:.1e})'
                worst[n] = max(worst.get(n, 0), eo)
            print(line)
        print(f'{kind}: worst rel-L2 per slice count: ' + ', '.join(f'{n}: {v:.1e}' for n, v in worst.items()))
    # cost model (per step of 128 runs: 3 Hankel applications)
    runs, apps = 128, 3
    macs = runs * apps * 2.0 * (L + 1) ** 2 * n_r * n_r            # real MACs: (re, im) rows x N_r x N_r
    print(f'per step: {macs / 1e9:.1f} G real MACs = {2 * macs / 1e9:.1f} GFLOP FP64')
    for n, n_prod in ((8, 36), (9, 45), (10, 55)):
        ops = 2 * macs * n_prod
        t_peak = ops / 4.5e15 * 1e3                                # dense int8 tcgen05 peak (nominal 4.5 POP/s)
        groups = n                                                # int32 accumulators converted + scaled + added in FP64 per output element
        epi_flops = runs * apps * 2.0 * (L + 1) ** 2 * n_r * groups * 3   # I2F + FMA (+ scale) per group and output element
        print(f'{n} slices: {n_prod} int8 GEMMs = {ops / 1e12:.2f} TOP -> {t_peak:.2f} ms at the nominal 4.5 POP/s (K = {n_r}: 4 MMA k-steps per tile, '
              f'epilogue bound in practice), FP64 epilogue {epi_flops / 1e9:.1f} GFLOP, operand slicing {n} x 1 B per 8 B element read')


if __name__ == '__main__':
    t0 = time.time()
    main()
    print(f'({time.time() - t0:.1f}s)')
