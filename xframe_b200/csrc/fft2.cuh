// Two-stage register FFT along phi for N = N1*N2 (64 = 8x8, 128 = 8x16, 256 = 16x16).
//
// The generic kernel in fft.cuh is shared-memory bound (4 Stockham passes through smem); here each thread does a
// radix-N1 DFT on values it loads straight from global memory (coalesced: N2 consecutive threads read N2 consecutive
// complex numbers), the twiddled results cross shared memory ONCE, and a second thread mapping does the radix-N2 DFT
// and writes the [shell][m][theta] layout with theta contiguous across threads.
//
//   forward:  X[u + N1 k2] = sum_t W_N^{ut} W_N2^{k2 t} sum_j x[t + N2 j] W_N1^{uj}
//   inverse:  x[t + N2 j]  = sum_u W_N1^{-uj} W_N^{-ut} sum_k2 X[u + N1 k2] W_N2^{-k2 t}
#pragma once
#include "common.cuh"
#include "pointwise.cuh"

__device__ __constant__ double kCos16[16] = {1.0, 0.9238795325112867, 0.7071067811865476, 0.38268343236508984, 0.0,
                                             -0.38268343236508984, -0.7071067811865476, -0.9238795325112867, -1.0,
                                             -0.9238795325112867, -0.7071067811865476, -0.38268343236508984, 0.0,
                                             0.38268343236508984, 0.7071067811865476, 0.9238795325112867};
__device__ __constant__ double kSin16[16] = {0.0, 0.38268343236508984, 0.7071067811865476, 0.9238795325112867, 1.0,
                                             0.9238795325112867, 0.7071067811865476, 0.38268343236508984, 0.0,
                                             -0.38268343236508984, -0.7071067811865476, -0.9238795325112867, -1.0,
                                             -0.9238795325112867, -0.7071067811865476, -0.38268343236508984};

// In-register DFT of R (power of two <= 16) points, natural order in and out, decimation in time.
// SIGN = -1: forward kernel exp(-2 pi i jk/R); +1: inverse (unnormalised).
template <int R, int SIGN>
__device__ __forceinline__ void dft_reg(double2 (&x)[R]) {
    if constexpr (R == 2) {
        const double2 a = x[0], b = x[1];
        x[0] = cadd(a, b);
        x[1] = csub(a, b);
    } else if constexpr (R > 2) {
        double2 e[R / 2], o[R / 2];
#pragma unroll
        for (int i = 0; i < R / 2; ++i) { e[i] = x[2 * i]; o[i] = x[2 * i + 1]; }
        dft_reg<R / 2, SIGN>(e);
        dft_reg<R / 2, SIGN>(o);
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
            const int idx = k * (16 / R);            // twiddle exp(SIGN 2 pi i k / R) = (cos, SIGN sin)(2 pi idx / 16)
            double2 t;
            if (idx == 0) t = o[k];
            else if (idx == 4) t = (SIGN < 0) ? make_double2(o[k].y, -o[k].x) : make_double2(-o[k].y, o[k].x);
            else {
                const double c = kCos16[idx], s = (SIGN < 0) ? -kSin16[idx] : kSin16[idx];
                t = make_double2(o[k].x * c - o[k].y * s, o[k].x * s + o[k].y * c);
            }
            x[k] = cadd(e[k], t);
            x[k + R / 2] = csub(e[k], t);
        }
    }
}

template <int N1, int N2>
struct Fft2Cfg {
    static constexpr int N = N1 * N2;
    static constexpr int TH = 256 / N1;              // theta rows per CTA (stage B: TH x N1 threads)
    static constexpr int ROWLEN = N + 1;             // padded smem row (complex): conflict-free transposed access
    static constexpr int PASSES = (TH * N2) / 256;   // stage A passes (rows per pass = 256 / N2)
    static constexpr size_t SMEM = (size_t)TH * ROWLEN * sizeof(double2);
};

// grid [S][n_theta][N] -> a [S][M2][n_theta]
template <int N1, int N2>
// minimum CTAs per SM of the launch bounds (= register budget).  Measured per step of 128 runs, phi-FFT group: forward 2 /
// fused-epilogue inverse 2 / plain inverse 2 -> 4.57 ms;  3 / 3 / 2 -> 4.99;  2 / 3 / 2 -> 4.73;  3 / 2 / 2 -> 4.90;
// 4 / 3 / 2 -> 5.57 (spills);  2 / 2 / 1 -> 5.18;  1 / 1 / 1 -> 6.03.
#ifndef FFT2_FWD_MINB
#define FFT2_FWD_MINB 2
#endif
#ifndef FFT2_MOD_MINB
#define FFT2_MOD_MINB 2
#endif
#ifndef FFT2_INV_MINB
#define FFT2_INV_MINB 2
#endif
__global__ void __launch_bounds__(256, FFT2_FWD_MINB) fft2_forward_kernel(SlotView grid, int shells_per_run, const double2* __restrict__ sub_flat,
                                                           double2* __restrict__ a, const double2* __restrict__ tw_g, int n_theta,
                                                           int l_max, int flags) {
    // flags: bit 0 = real input, write only m >= 0;  bit 1 = transform |x|^2 instead of x (square_grid fused, misk.py:159-168)
    using C = Fft2Cfg<N1, N2>;
    extern __shared__ double2 smem_f2[];
    const int pos_only = flags & 1, square = flags & 2;
    const int s = blockIdx.x, theta0 = blockIdx.y * C::TH, tid = threadIdx.x;
    const int M2 = 2 * l_max + 1;
    const int run = s / shells_per_run, shell_in_run = s - run * shells_per_run;
    const double2* src = slot_run_ptr(grid, run) + ((size_t)shell_in_run * n_theta + theta0) * C::N;
    const double2* sub = sub_flat ? sub_flat + ((size_t)s * n_theta + theta0) * C::N : nullptr;
    // ---- stage A: thread (row, t), radix-N1 over j of x[t + N2 j]
    {
        const int t = tid % N2, row0 = tid / N2;
#pragma unroll
        for (int ps = 0; ps < C::PASSES; ++ps) {
            const int row = row0 + ps * (256 / N2);
            double2 x[N1];
#pragma unroll
            for (int j = 0; j < N1; ++j) {
                x[j] = src[(size_t)row * C::N + t + N2 * j];
                if (sub) { const double2 w = ldg2(sub + (size_t)row * C::N + t + N2 * j); x[j].x -= w.x; x[j].y -= w.y; }
                if (square) x[j] = make_double2(__dadd_rn(__dmul_rn(x[j].x, x[j].x), __dmul_rn(x[j].y, x[j].y)), 0.0);
            }
            dft_reg<N1, -1>(x);
#pragma unroll
            for (int u = 0; u < N1; ++u) {
                double2 v = x[u];
                if (u > 0) v = cmul(v, ldg2(tw_g + ((t * u) & (C::N - 1))));          // W_N^{tu}, forward table
                smem_f2[row * C::ROWLEN + u * N2 + t] = v;
            }
        }
    }
    __syncthreads();
    // ---- stage B: thread (theta_local, u), radix-N2 over t, outputs k = u + N1 k2 with theta contiguous over threads
    {
        const int th = tid % C::TH, u = tid / C::TH;
        double2 z[N2];
#pragma unroll
        for (int t = 0; t < N2; ++t) z[t] = smem_f2[th * C::ROWLEN + u * N2 + t];
        dft_reg<N2, -1>(z);
        double2* dst = a + (size_t)s * M2 * n_theta + theta0 + th;
#pragma unroll
        for (int k2 = 0; k2 < N2; ++k2) {
            const int k = u + N1 * k2;
            const int m = (k <= C::N / 2) ? k : k - C::N;
            if (m >= (pos_only ? 0 : -l_max) && m <= l_max) {     // pos_only: real input, the m<0 half is redundant
                const int mm = (m >= 0) ? m : M2 + m;
                dst[(size_t)mm * n_theta] = z[k2];
            }
        }
    }
}

// a [S][M2][n_theta] -> grid [S][n_theta][N] (unnormalised inverse DFT)
template <int N1, int N2, bool MOD>
// mod_rho_hat != nullptr: the transform output is I_proj and the kernel writes the modified-intensity density instead
// (project_to_modified_intensity fused, fxs_Projections.py:899-909): mod_out[x] = rho_hat[x] sqrt(Re I_proj[x] / |rho_hat[x]|^2)
__global__ void __launch_bounds__(256, MOD ? FFT2_MOD_MINB : FFT2_INV_MINB) fft2_inverse_kernel(const double2* __restrict__ a, double2* __restrict__ grid,
                                                           const double2* __restrict__ tw_g, int n_theta, int l_max, int herm,
                                                           const double2* __restrict__ mod_rho_hat, SlotView mod_out, int shells_per_run) {
    using C = Fft2Cfg<N1, N2>;
    extern __shared__ double2 smem_f2[];
    const int s = blockIdx.x, theta0 = blockIdx.y * C::TH, tid = threadIdx.x;
    const int M2 = 2 * l_max + 1;
    // ---- stage A': thread (theta_local, u): X[u + N1 k2] -> radix-N2 inverse over k2 -> Y[u][t] * W_N^{-ut}
    {
        const int th = tid % C::TH, u = tid / C::TH;
        const double2* src = a + (size_t)s * M2 * n_theta + theta0 + th;
        double2 z[N2];
#pragma unroll
        for (int k2 = 0; k2 < N2; ++k2) {
            const int k = u + N1 * k2;
            const int m = (k <= C::N / 2) ? k : k - C::N;
            double2 v = make_double2(0.0, 0.0);
            if (m >= -l_max && m <= l_max) {
                if (herm && m < 0) { v = ldg2(src + (size_t)(-m) * n_theta); v.y = -v.y; }   // X[-m] = conj(X[m]): real output
                else v = ldg2(src + (size_t)((m >= 0) ? m : M2 + m) * n_theta);
            }
            z[k2] = v;
        }
        dft_reg<N2, +1>(z);
#pragma unroll
        for (int t = 0; t < N2; ++t) {
            double2 v = z[t];
            if (u > 0) {
                double2 w = ldg2(tw_g + ((t * u) & (C::N - 1)));
                w.y = -w.y;                                                          // conj: W_N^{-ut}
                v = cmul(v, w);
            }
            smem_f2[th * C::ROWLEN + u * N2 + t] = v;
        }
    }
    __syncthreads();
    // ---- stage B': thread (row, t): radix-N1 inverse over u -> x[t + N2 j]
    {
        const int t = tid % N2, row0 = tid / N2;
        double2* dst = grid + ((size_t)s * n_theta + theta0) * C::N;
        const double2* rh = nullptr;
        if constexpr (MOD) {
            const int run = s / shells_per_run, shell_in_run = s - run * shells_per_run;
            rh = mod_rho_hat + ((size_t)s * n_theta + theta0) * C::N;
            dst = slot_run_ptr(mod_out, run) + ((size_t)shell_in_run * n_theta + theta0) * C::N;
        }
#pragma unroll
        for (int ps = 0; ps < C::PASSES; ++ps) {
            const int row = row0 + ps * (256 / N2);
            double2 y[N1];
#pragma unroll
            for (int u = 0; u < N1; ++u) y[u] = smem_f2[row * C::ROWLEN + u * N2 + t];
            dft_reg<N1, +1>(y);
#pragma unroll
            for (int j = 0; j < N1; ++j) {
                double2 o = y[j];
                if constexpr (MOD) {
                    const double2 v = ldg2(rh + (size_t)row * C::N + t + N2 * j);
                    const double sq = __dadd_rn(__dmul_rn(v.x, v.x), __dmul_rn(v.y, v.y));
                    const double mult = mod_intensity_multiplier(o.x, sq);
                    o = make_double2(v.x * mult, v.y * mult);
                }
                dst[(size_t)row * C::N + t + N2 * j] = o;
            }
        }
    }
}

template <int N1, int N2>
static int launch_fft2(bool forward, SlotView in, int shells_per_run, const double2* sub, double2* out, const double2* tw, int n_shells,
                       int n_theta, int l_max, cudaStream_t st, int half, const double2* mod_rho_hat, SlotView mod_out) {
    using C = Fft2Cfg<N1, N2>;
    static XfbPerDeviceOnce attr_once;
    if (xfb_first_on_device(attr_once)) {
        XFB_CUDA(cudaFuncSetAttribute(fft2_forward_kernel<N1, N2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        XFB_CUDA(cudaFuncSetAttribute(fft2_inverse_kernel<N1, N2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        XFB_CUDA(cudaFuncSetAttribute(fft2_inverse_kernel<N1, N2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    }
    dim3 g(n_shells, n_theta / C::TH);
    if (forward)
        fft2_forward_kernel<N1, N2><<<g, 256, C::SMEM, st>>>(in, shells_per_run, sub, out, tw, n_theta, l_max, half);
    else if (mod_rho_hat)
        fft2_inverse_kernel<N1, N2, true><<<g, 256, C::SMEM, st>>>(in.base, out, tw, n_theta, l_max, half, mod_rho_hat, mod_out, shells_per_run);
    else
        fft2_inverse_kernel<N1, N2, false><<<g, 256, C::SMEM, st>>>(in.base, out, tw, n_theta, l_max, half, mod_rho_hat, mod_out, shells_per_run);
    XFB_CUDA(cudaGetLastError());
    return 0;
}

// returns -1 when (n_phi, n_theta) is not covered by the register FFT (caller falls back to fft.cuh)
static inline bool fft2_covers(int n_phi, int n_theta) {
    return (n_phi == 64 && n_theta % Fft2Cfg<8, 8>::TH == 0) || (n_phi == 128 && n_theta % Fft2Cfg<8, 16>::TH == 0) ||
           (n_phi == 256 && n_theta % Fft2Cfg<16, 16>::TH == 0);
}
static int launch_fft2_any(bool forward, int n_phi, SlotView in, int shells_per_run, const double2* sub, double2* out, const double2* tw,
                           int n_shells, int n_theta, int l_max, cudaStream_t st, int half, const double2* mod_rho_hat = nullptr,
                           SlotView mod_out = SlotView{}) {
    if (n_phi == 64 && n_theta % Fft2Cfg<8, 8>::TH == 0) return launch_fft2<8, 8>(forward, in, shells_per_run, sub, out, tw, n_shells, n_theta, l_max, st, half, mod_rho_hat, mod_out);
    if (n_phi == 128 && n_theta % Fft2Cfg<8, 16>::TH == 0) return launch_fft2<8, 16>(forward, in, shells_per_run, sub, out, tw, n_shells, n_theta, l_max, st, half, mod_rho_hat, mod_out);
    if (n_phi == 256 && n_theta % Fft2Cfg<16, 16>::TH == 0) return launch_fft2<16, 16>(forward, in, shells_per_run, sub, out, tw, n_shells, n_theta, l_max, st, half, mod_rho_hat, mod_out);
    return -1;
}
