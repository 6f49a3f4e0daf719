// 2-D (polar) flavour of the fxs path: circular-harmonic transforms and the per-order invariant projection.
//
// Reference semantics
//   circularHarmonicTransform_complex_forward / _inverse : fft(x, axis=1)/n_phi , ifft(c*n_phi, axis=1)
//   circularHarmonicTransform_real_forward / _inverse    : rfft(Re x)/n_phi , irfft(c*n_phi, n_phi)
//                                                          (xframe/library/mathLibrary.py:469-496)
//   approximate_unknowns (2-D), mtip_projection (2-D)    : fxs_Projections.py:723-750, 792-830, 852-862
//   restated in oracle/mtip2d.py.
//
// n_phi = 2M+1 is odd (harmonic_transforms.py:44-47,60; 127 for max_order 63), often prime, so the transform is a
// direct O(n_phi^2) DFT per shell with a twiddle table -- 127 points x 16 B per shell are read once from HBM and the
// 4 n_phi^2 FP64 FMAs per shell stay in shared memory / registers.
//
// Layouts
//   grid [S][N] complex (phi contiguous; S = runs * N_r shells)
//   c2   [N][S] complex, j = DFT index (order m = j for j <= M, j - N for j > M): the orders of one m are contiguous
//        over the shells, which is the operand layout of the radial (Hankel) GEMM (gemm.cuh:hankel_kernel).
#pragma once
#include "common.cuh"

#define DFT_ROWS 16          // shells per CTA
#define DFT_THREADS 256
__host__ __device__ inline int dft_ld(int n) { return (n & 1) ? n + 2 : n + 1; }   // odd row stride: conflict-free 128-bit columns
static inline size_t dft_smem(int n) { return (size_t)(DFT_ROWS * dft_ld(n) + n) * sizeof(double2); }

// forward: c2[j][s] = scale * sum_phi x[s][phi] exp(-2 pi i j phi / N);  real_only: use Re x (rfft semantics, all j are
// still produced: c[N-j] = conj(c[j])).  `sub`, when given, is subtracted from the input first (linearity of the FT).
__global__ void __launch_bounds__(DFT_THREADS) dft2d_forward_kernel(SlotView grid, int shells_per_run, const double2* __restrict__ sub_flat,
                                                                   double2* __restrict__ c2, int S, int N, double scale, int real_only) {
    extern __shared__ double2 smem_dft[];
    const int ld = dft_ld(N);
    double2* xs = smem_dft;                     // [DFT_ROWS][ld]
    double2* tw = xs + DFT_ROWS * ld;           // [N]
    const int s0 = blockIdx.x * DFT_ROWS, tid = threadIdx.x;
    for (int k = tid; k < N; k += DFT_THREADS) {
        double sn, cs;
        sincospi(-2.0 * (double)k / (double)N, &sn, &cs);
        tw[k] = make_double2(cs, sn);
    }
    for (int idx = tid; idx < DFT_ROWS * N; idx += DFT_THREADS) {
        const int row = idx / N, ph = idx - row * N;
        const int s = s0 + row;
        double2 v = make_double2(0.0, 0.0);
        if (s < S) {
            const int b = s / shells_per_run, r = s - b * shells_per_run;
            v = slot_run_ptr(grid, b)[(long long)r * N + ph];
            if (sub_flat) { const double2 t = ldg2(sub_flat + (long long)s * N + ph); v.x -= t.x; v.y -= t.y; }
            if (real_only) v.y = 0.0;
        }
        xs[row * ld + ph] = v;
    }
    __syncthreads();
    for (int idx = tid; idx < DFT_ROWS * N; idx += DFT_THREADS) {
        const int j = idx / DFT_ROWS, row = idx - j * DFT_ROWS;       // consecutive threads -> consecutive shells of one order
        const double2* x = xs + row * ld;
        double ar = 0.0, ai = 0.0;
        int k = 0;
        for (int ph = 0; ph < N; ++ph) {
            const double2 w = tw[k], v = x[ph];
            ar += v.x * w.x - v.y * w.y;
            ai += v.x * w.y + v.y * w.x;
            k += j; if (k >= N) k -= N;
        }
        if (s0 + row < S) c2[(size_t)j * S + s0 + row] = make_double2(ar * scale, ai * scale);
    }
}

// inverse: x[s][phi] = sum_j c2[j][s] exp(+2 pi i j phi / N).  herm: irfft semantics -- only j <= N/2 is read, the other
// half is its conjugate mirror and the imaginary part of c[0] is ignored (numpy.fft.irfft with odd N).
__global__ void __launch_bounds__(DFT_THREADS) dft2d_inverse_kernel(const double2* __restrict__ c2, double2* __restrict__ grid, int S, int N,
                                                                   int herm) {
    extern __shared__ double2 smem_dft[];
    const int ld = dft_ld(N);
    double2* cs_ = smem_dft;                    // [DFT_ROWS][ld]  (row = shell, column = j)
    double2* tw = cs_ + DFT_ROWS * ld;
    const int s0 = blockIdx.x * DFT_ROWS, tid = threadIdx.x;
    for (int k = tid; k < N; k += DFT_THREADS) {
        double sn, cs;
        sincospi(2.0 * (double)k / (double)N, &sn, &cs);
        tw[k] = make_double2(cs, sn);
    }
    for (int idx = tid; idx < DFT_ROWS * N; idx += DFT_THREADS) {
        const int j = idx / DFT_ROWS, row = idx - j * DFT_ROWS;
        double2 v = make_double2(0.0, 0.0);
        if (s0 + row < S) {
            if (herm && j > N / 2) { v = ldg2(c2 + (size_t)(N - j) * S + s0 + row); v.y = -v.y; }
            else v = ldg2(c2 + (size_t)j * S + s0 + row);
            if (herm && j == 0) v.y = 0.0;
        }
        cs_[row * ld + j] = v;
    }
    __syncthreads();
    for (int idx = tid; idx < DFT_ROWS * N; idx += DFT_THREADS) {
        const int row = idx / N, ph = idx - row * N;                  // consecutive threads -> consecutive phi of one shell
        const double2* c = cs_ + row * ld;
        double ar = 0.0, ai = 0.0;
        int k = 0;
        for (int j = 0; j < N; ++j) {
            const double2 w = tw[k], v = c[j];
            ar += v.x * w.x - v.y * w.y;
            ai += v.x * w.y + v.y * w.x;
            k += ph; if (k >= N) k -= N;
        }
        if (s0 + row < S) grid[(size_t)(s0 + row) * N + ph] = make_double2(ar, ai);
    }
}

// Invariant projection, 2-D (one CTA per run, one warp per order m = 0..M):
//   u_m = s/|s|,  s = sum_q I_m(q) conj(V_m(q)) q   (1 if s == 0; 1 for the pinned SO order)      fxs_Projections.py:727-748
//   I'_m(q) = V_m(q) u_m on the radial mask, I'_0 = V_0 there, then I'_0 /= sqrt(N_particles) everywhere   :818-826,852-862
// c_in holds the full DFT of the real intensity; only m <= M is read (rfft half).  c_out gets m <= M and its conjugate
// mirror (the inverse transform of the loop uses irfft semantics and reads m <= M only).
__global__ void project2d_kernel(const double2* __restrict__ c_in, double2* __restrict__ c_out, const double2* __restrict__ V,
                                 const uint8_t* __restrict__ radial_mask, const double* __restrict__ q, double2* __restrict__ unknowns,
                                 int n_orders, int M, int N, int n_r, int S, double inv_sqrt_np, int so_order) {
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int m = warp; m <= M; m += nwarp) {
        const double2* src = c_in + (size_t)m * S + (size_t)b * n_r;
        double2* dst = c_out + (size_t)m * S + (size_t)b * n_r;
        double2* dstn = (m > 0) ? c_out + (size_t)(N - m) * S + (size_t)b * n_r : nullptr;
        double2 u = make_double2(1.0, 0.0);
        const bool used = m < n_orders;
        if (used) {
            double sx = 0.0, sy = 0.0;
            for (int k = lane; k < n_r; k += 32) {
                const double2 i = ldg2(src + k), v = V[(size_t)m * n_r + k];
                const double w = q[k];
                sx += (i.x * v.x + i.y * v.y) * w;          // I conj(V)
                sy += (i.y * v.x - i.x * v.y) * w;
            }
            sx = warp_sum(sx); sy = warp_sum(sy);
            sx = __shfl_sync(0xffffffffu, sx, 0); sy = __shfl_sync(0xffffffffu, sy, 0);
            if (sx != 0.0 || sy != 0.0) { const double a = hypot(sx, sy); u = make_double2(sx / a, sy / a); }
            if (m == so_order) u = make_double2(1.0, 0.0);
            if (lane == 0) unknowns[(size_t)b * n_orders + m] = u;
        }
        for (int k = lane; k < n_r; k += 32) {
            double2 o = ldg2(src + k);
            if (used && radial_mask[(size_t)m * n_r + k]) {
                const double2 v = V[(size_t)m * n_r + k];
                o = (m == 0) ? v : make_double2(v.x * u.x - v.y * u.y, v.x * u.y + v.y * u.x);
            }
            if (m == 0) { o.x *= inv_sqrt_np; o.y *= inv_sqrt_np; }
            dst[k] = o;
            if (dstn) dstn[k] = make_double2(o.x, -o.y);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Folded DFT of odd length N = 2M+1 (the circular-harmonic transform of the 2-D path) on the FP64 tensor cores.
//   cos(2 pi j k / N) is even and sin odd under j -> N - j and under k -> N - k, so with
//       e_0 = x_0, e_j = x_j + x_{N-j},  o_j = x_j - x_{N-j}            (j = 1..M)
//       A_k = sum_j e_j cos(2 pi j k / N),  B_k = sum_j o_j sin(2 pi j k / N)      (k = 0..M)
//   the forward transform is Y_k = A_k - i B_k, Y_{N-k} = A_k + i B_k and the inverse x_j = A'_j + i B'_j, x_{N-j} = A'_j - i B'_j
//   with (e', o') built from (c_k, c_{N-k}): two real [H x H] matrices (H = M+1 padded to 16) instead of two [N x N] ones --
//   a quarter of the flops of the dense product, and the transposition between the row layout [S][N] of the grids and the
//   order-major layout [N][S] of the coefficients is done by the fold / combine kernels themselves.
//   Row buffers: eo [2][S][H] (block 0 = e rows, block 1 = o rows), ab [2][S][H] (A rows, B rows).
// ---------------------------------------------------------------------------------------------------------------------
// forward fold: grid rows (slot view) -> eo
__global__ void dft_fold_rows_kernel(SlotView grid, int shells_per_run, double2* __restrict__ eo, int S, int N, int H) {
    const int M = N / 2;
    const long long n = (long long)S * H;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int s = (int)(idx / H), j = (int)(idx - (long long)s * H);
        const int b = s / shells_per_run, r = s - b * shells_per_run;
        const double2* x = slot_run_ptr(grid, b) + (long long)r * N;
        double2 e = make_double2(0.0, 0.0), o = e;
        if (j == 0) e = x[0];
        else if (j <= M) {
            const double2 u = x[j], v = x[N - j];
            e = make_double2(u.x + v.x, u.y + v.y);
            o = make_double2(u.x - v.x, u.y - v.y);
        }
        eo[idx] = e;
        eo[n + idx] = o;
    }
}
// forward combine (+ transposition): ab -> c2[k][s] = scale (A - i B), c2[N-k][s] = scale (A + i B)
__global__ void dft_combine_to_orders_kernel(const double2* __restrict__ ab, double2* __restrict__ c2, int S, int N, int H, double scale) {
    __shared__ double2 ta[32][33], tb[32][33];
    const int M = N / 2;
    const int k0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
    const long long n = (long long)S * H;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {           // rows: k contiguous
        const int s = s0 + i, k = k0 + threadIdx.x;
        if (s < S && k <= M) { ta[i][threadIdx.x] = ab[(long long)s * H + k]; tb[i][threadIdx.x] = ab[n + (long long)s * H + k]; }
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {           // columns: s contiguous
        const int k = k0 + i, s = s0 + threadIdx.x;
        if (s < S && k <= M) {
            const double2 a = ta[threadIdx.x][i], b = tb[threadIdx.x][i];
            c2[(long long)k * S + s] = make_double2(scale * (a.x + b.y), scale * (a.y - b.x));                 // A - i B
            if (k > 0) c2[(long long)(N - k) * S + s] = make_double2(scale * (a.x - b.y), scale * (a.y + b.x));   // A + i B
        }
    }
}
// inverse fold (+ transposition): c2[k][s] -> eo;  herm: only k <= M is valid and c_{N-k} = conj c_k, Im c_0 = 0 (irfft)
__global__ void dft_fold_orders_kernel(const double2* __restrict__ c2, double2* __restrict__ eo, int S, int N, int H, int herm) {
    __shared__ double2 te[32][33], to_[32][33];
    const int M = N / 2;
    const int k0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
    const long long n = (long long)S * H;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {           // s contiguous
        const int k = k0 + i, s = s0 + threadIdx.x;
        double2 e = make_double2(0.0, 0.0), o = e;
        if (s < S && k <= M) {
            const double2 u = c2[(long long)k * S + s];
            if (k == 0) e = herm ? make_double2(u.x, 0.0) : u;
            else {
                const double2 v = herm ? make_double2(u.x, -u.y) : c2[(long long)(N - k) * S + s];
                e = make_double2(u.x + v.x, u.y + v.y);
                o = make_double2(u.x - v.x, u.y - v.y);
            }
        }
        te[i][threadIdx.x] = e; to_[i][threadIdx.x] = o;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {           // k contiguous
        const int s = s0 + i, k = k0 + threadIdx.x;
        if (s < S && k < H) { eo[(long long)s * H + k] = te[threadIdx.x][i]; eo[n + (long long)s * H + k] = to_[threadIdx.x][i]; }
    }
}
// inverse combine: ab -> grid rows x[s][j] = A + i B, x[s][N-j] = A - i B
__global__ void dft_combine_to_rows_kernel(const double2* __restrict__ ab, double2* __restrict__ rows, int S, int N, int H) {
    const int M = N / 2;
    const long long n = (long long)S * H, tot = (long long)S * (M + 1);
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < tot; idx += (long long)gridDim.x * blockDim.x) {
        const int s = (int)(idx / (M + 1)), j = (int)(idx - (long long)s * (M + 1));
        const double2 a = ab[(long long)s * H + j], b = ab[n + (long long)s * H + j];
        double2* x = rows + (long long)s * N;
        x[j] = make_double2(a.x - b.y, a.y + b.x);                                   // A + i B
        if (j > 0) x[N - j] = make_double2(a.x + b.y, a.y - b.x);                    // A - i B
    }
}

