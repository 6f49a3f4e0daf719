// phi-FFT stage of the spherical harmonic transform (shtns-style: FFT along phi,
// then a per-m Legendre contraction).  Bandwidth-bound: every grid row is read
// once with 128-bit accesses, transformed in shared memory by one warp
// (Stockham radix-4/2, autosort) and written once, transposed to the
// [shell][m][theta] layout the Legendre GEMM consumes.
//
// Reference semantics: the FFT inside shtns' analys_cplx/synth_cplx
// (shtns_plugin.py:222,233) -- restated in oracle/sht.py (np.fft.fft / ifft*n).
#pragma once
#include "common.cuh"
#include "pointwise.cuh"
#include "fft2.cuh"

// padded index inside a shared-memory row: one pad element every 4 keeps the
// strided Stockham writes (stride 4 / 16 / 64 complex) off the same banks.
#define XFB_PHYS(i) ((i) + ((i) >> 2))
__host__ __device__ inline int xfb_fft_rowlen(int n) { return n + (n >> 2) + 1; }

// One Stockham pass (radix R chosen at compile time) and the recursion over passes.
// Formulation: for butterfly j, k = j mod Ns, inputs x[j + r*N/R] * w^(r*k*N/(Ns*R)),
// outputs y[(j-k)*R + k + r*Ns]  (autosort, mixed radix 4,4,..,[2]).
template <int N, int SIGN, int Ns>
__device__ __forceinline__ void fft_passes(double2* __restrict__ row, const double2* __restrict__ tw, int lane) {
    if constexpr (Ns < N) {
        constexpr int R = ((N / Ns) % 4 == 0) ? 4 : 2;
        constexpr int nbf = N / R;
        constexpr int U = (nbf + 31) / 32;  // butterflies per lane
        constexpr int ts = N / (Ns * R);
        double2 v[U * R];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = lane + 32 * u;
            if (j < nbf) {
#pragma unroll
                for (int r = 0; r < R; ++r) v[u * R + r] = row[XFB_PHYS(j + r * nbf)];
            }
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = lane + 32 * u;
            if (j < nbf) {
                const int k = j & (Ns - 1);
                double2 x[R];
                x[0] = v[u * R];
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    double2 w = tw[(r * k * ts) & (N - 1)];
                    if (SIGN > 0) w.y = -w.y;
                    x[r] = cmul(v[u * R + r], w);
                }
                const int j0 = (j - k) * R + k;
                if constexpr (R == 4) {
                    const double2 t0 = cadd(x[0], x[2]), t1 = csub(x[0], x[2]);
                    const double2 t2 = cadd(x[1], x[3]);
                    const double2 d = csub(x[1], x[3]);
                    // forward: multiply by -i ; inverse: by +i
                    const double2 t3 = (SIGN < 0) ? make_double2(d.y, -d.x) : make_double2(-d.y, d.x);
                    row[XFB_PHYS(j0)] = cadd(t0, t2);
                    row[XFB_PHYS(j0 + Ns)] = cadd(t1, t3);
                    row[XFB_PHYS(j0 + 2 * Ns)] = csub(t0, t2);
                    row[XFB_PHYS(j0 + 3 * Ns)] = csub(t1, t3);
                } else {
                    row[XFB_PHYS(j0)] = cadd(x[0], x[1]);
                    row[XFB_PHYS(j0 + Ns)] = csub(x[0], x[1]);
                }
            }
        }
        __syncwarp();
        fft_passes<N, SIGN, Ns * R>(row, tw, lane);
    }
}

// One warp transforms one row of N complex numbers in place (row is in padded layout).
template <int N, int SIGN>
__device__ __forceinline__ void warp_fft_row(double2* __restrict__ row, const double2* __restrict__ tw, int lane) {
    fft_passes<N, SIGN, 1>(row, tw, lane);
}

// grid [S][n_theta][N] -> a [S][M2][n_theta], M2 = 2L+1, mm = m (m>=0) or M2+m (m<0)
template <int N>
__global__ void __launch_bounds__(256) fft_phi_forward_kernel(SlotView grid, int shells_per_run, const double2* __restrict__ sub_flat,
                                                              double2* __restrict__ a, const double2* __restrict__ tw_g, int n_theta,
                                                              int l_max, int th, int pos_only) {
    extern __shared__ double2 smem_fft[];
    const int rowlen = xfb_fft_rowlen(N);
    double2* tw = smem_fft;
    double2* buf = smem_fft + N;
    const int s = blockIdx.x;
    const int theta0 = blockIdx.y * th;
    const int tid = threadIdx.x;
    const int M2 = 2 * l_max + 1;
    for (int i = tid; i < N; i += blockDim.x) tw[i] = tw_g[i];
    const int run = s / shells_per_run, shell_in_run = s - run * shells_per_run;
    const double2* src = slot_run_ptr(grid, run) + ((size_t)shell_in_run * n_theta + theta0) * N;
    const double2* sub = sub_flat ? sub_flat + ((size_t)s * n_theta + theta0) * N : nullptr;   // transform of (grid - sub)
    for (int idx = tid; idx < th * N; idx += blockDim.x) {
        const int t = idx / N, i = idx - t * N;
        double2 v = src[idx];
        if (sub) { const double2 w = ldg2(sub + idx); v.x -= w.x; v.y -= w.y; }
        buf[t * rowlen + XFB_PHYS(i)] = v;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
    for (int t = warp; t < th; t += nwarp) warp_fft_row<N, -1>(buf + t * rowlen, tw, lane);
    __syncthreads();
    double2* dst = a + (size_t)s * M2 * n_theta + theta0;
    for (int idx = tid; idx < M2 * th; idx += blockDim.x) {
        const int mm = idx / th, t = idx - mm * th;
        if (pos_only && mm > l_max) continue;                       // real input: the m<0 half is redundant
        const int mi = (mm <= l_max) ? mm : N - (M2 - mm);
        dst[(size_t)mm * n_theta + t] = buf[t * rowlen + XFB_PHYS(mi)];
    }
}

// a [S][M2][n_theta] -> grid [S][n_theta][N]  (unnormalised inverse DFT = synthesis sum over m)
template <int N>
__global__ void __launch_bounds__(256) fft_phi_inverse_kernel(const double2* __restrict__ a, double2* __restrict__ grid,
                                                              const double2* __restrict__ tw_g, int n_theta, int l_max, int th,
                                                              int herm) {
    extern __shared__ double2 smem_fft[];
    const int rowlen = xfb_fft_rowlen(N);
    double2* tw = smem_fft;
    double2* buf = smem_fft + N;
    const int s = blockIdx.x;
    const int theta0 = blockIdx.y * th;
    const int tid = threadIdx.x;
    const int M2 = 2 * l_max + 1;
    for (int i = tid; i < N; i += blockDim.x) tw[i] = tw_g[i];
    const double2* src = a + (size_t)s * M2 * n_theta + theta0;
    for (int idx = tid; idx < N * th; idx += blockDim.x) {
        const int i = idx / th, t = idx - i * th;
        const int m = (i <= N / 2) ? i : i - N;
        double2 val = make_double2(0.0, 0.0);
        if (m >= -l_max && m <= l_max) {
            if (herm && m < 0) { val = ldg2(src + (size_t)(-m) * n_theta + t); val.y = -val.y; }   // X[-m] = conj(X[m])
            else val = ldg2(src + (size_t)((m >= 0) ? m : M2 + m) * n_theta + t);
        }
        buf[t * rowlen + XFB_PHYS(i)] = val;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
    for (int t = warp; t < th; t += nwarp) warp_fft_row<N, +1>(buf + t * rowlen, tw, lane);
    __syncthreads();
    double2* dst = grid + ((size_t)s * n_theta + theta0) * N;
    for (int idx = tid; idx < th * N; idx += blockDim.x) {
        const int t = idx / N, i = idx - t * N;
        dst[idx] = buf[t * rowlen + XFB_PHYS(i)];
    }
}

template <int N>
static int launch_fft_n(bool forward, SlotView in, int shells_per_run, const double2* sub, double2* out, const double2* tw, int n_shells, int n_theta,
                        int l_max, cudaStream_t st, int half) {
    const int th = (n_theta % 16 == 0) ? 16 : 8;
    const size_t smem = (size_t)(N + th * xfb_fft_rowlen(N)) * sizeof(double2);
    dim3 g(n_shells, n_theta / th);
    // dynamic smem can exceed 48 KB: opt in once per instantiation, sized for the largest theta block
    const int smem_max = (int)((size_t)(N + 16 * xfb_fft_rowlen(N)) * sizeof(double2));
    static XfbPerDeviceOnce attr_once;
    if (xfb_first_on_device(attr_once)) {
        XFB_CUDA(cudaFuncSetAttribute(fft_phi_forward_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        XFB_CUDA(cudaFuncSetAttribute(fft_phi_inverse_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    }
    if (forward)
        fft_phi_forward_kernel<N><<<g, 256, smem, st>>>(in, shells_per_run, sub, out, tw, n_theta, l_max, th, half);
    else
        fft_phi_inverse_kernel<N><<<g, 256, smem, st>>>(in.base, out, tw, n_theta, l_max, th, half);
    XFB_CUDA(cudaGetLastError());
    return 0;
}

// half: forward -> write only m >= 0 (real input); inverse -> synthesise from m >= 0 with conjugate symmetry (real output)
// half: bit 0 = real field (forward: write m >= 0 only; inverse: conjugate-symmetric synthesis); bit 1 (forward, register
// FFT only) = transform |x|^2.  mod_rho_hat / mod_out (inverse, register FFT only): fused modified-intensity epilogue.
static int launch_fft(bool forward, int n_phi, SlotView in, int shells_per_run, const double2* sub, double2* out, const double2* tw, int n_shells,
                      int n_theta, int l_max, cudaStream_t st, int half = 0, const double2* mod_rho_hat = nullptr,
                      SlotView mod_out = SlotView{}) {
    {   // register two-stage FFT where it applies (64 / 128 / 256 points), generic Stockham kernel otherwise
        const int rc = launch_fft2_any(forward, n_phi, in, shells_per_run, sub, out, tw, n_shells, n_theta, l_max, st, half, mod_rho_hat, mod_out);
        if (rc >= 0) return rc;
    }
    if ((half & 2) || mod_rho_hat) XFB_FAIL("fused square / modified-intensity FFT variants need the register FFT (n_phi 64/128/256)");
    half &= 1;
    switch (n_phi) {
        case 16: return launch_fft_n<16>(forward, in, shells_per_run, sub, out, tw, n_shells, n_theta, l_max, st, half);
        case 32: return launch_fft_n<32>(forward, in, shells_per_run, sub, out, tw, n_shells, n_theta, l_max, st, half);
        case 64: return launch_fft_n<64>(forward, in, shells_per_run, sub, out, tw, n_shells, n_theta, l_max, st, half);
        case 128: return launch_fft_n<128>(forward, in, shells_per_run, sub, out, tw, n_shells, n_theta, l_max, st, half);
        case 256: return launch_fft_n<256>(forward, in, shells_per_run, sub, out, tw, n_shells, n_theta, l_max, st, half);
        case 512: return launch_fft_n<512>(forward, in, shells_per_run, sub, out, tw, n_shells, n_theta, l_max, st, half);
        default: XFB_FAIL("n_phi=%d unsupported (power of two in [16,512])", n_phi);
    }
}
