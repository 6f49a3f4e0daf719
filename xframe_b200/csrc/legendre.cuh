// Per-m associated-Legendre contraction of the SHT as FP64 tensor-core GEMMs
// (DMMA.8x8x4), with the north/south mirror symmetry P_l^m(-x) = (-1)^(l+m) P_l^m(x)
// folded in: the theta contraction runs over n_theta/2 nodes, separately for
// (l-m) even and odd.
//
// Reference semantics: analys_cplx / synth_cplx of shtns (shtns_plugin.py:218-261),
// restated in oracle/sht.py:ShtCore.analys_batch / synth_batch.
//
// Data layouts
//   a : [S][M2][n_theta] complex, M2 = 2L+1, mm = m (m>=0) / M2+m (m<0)   (phi-Fourier space)
//   c : [(L+1)^2][S] complex, row index l*(l+1)+m, S shells contiguous        (internal coefficient layout)
//   tables: FE/FO [L+1][K2][NP] (forward: w_j * 2pi/n_phi * P), IE/IO [L+1][NP][K2] (inverse: P)
//     K2 = n_theta/2, NP = number of same-parity degrees rounded up to 8, zero padded.
#pragma once
#include "common.cuh"

#define LEG_ROWS 32      // 16 shells x 2 signs of m
#define LEG_NB 32        // output columns per pass (4 warps x 8)
#define LEG_LDB (LEG_NB + 4)

#define LEG_THREADS 256  // 8 warps: warp w -> output columns (w&3)*8.., rows (w>>2)*16..
// acc[mb][0..1] += A[16 x K] (rows r0 + mb*8.., smem [row*lda + k]) * B[K x 8] (smem [k*ldb + n0 + n])
__device__ __forceinline__ void leg_mma_cplx(const double* __restrict__ Are, const double* __restrict__ Aim, int lda,
                                             const double* __restrict__ Bs, int ldb, int n0, int r0, int K, int lane,
                                             double (&cre)[2][2], double (&cim)[2][2]) {
    const int ar = lane >> 2, ak = lane & 3;
    for (int k0 = 0; k0 < K; k0 += 4) {
        const double b = Bs[(k0 + ak) * ldb + n0 + ar];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            const int off = (r0 + mb * 8 + ar) * lda + k0 + ak;
            dmma884(cre[mb][0], cre[mb][1], Are[off], b);
            dmma884(cim[mb][0], cim[mb][1], Aim[off], b);
        }
    }
}

__global__ void __launch_bounds__(LEG_THREADS, 4) legendre_forward_kernel(const double2* __restrict__ a, double2* __restrict__ c,
                                                               const double* __restrict__ FE, const double* __restrict__ FO,
                                                               int S, int l_max, int n_theta, int NP) {
    extern __shared__ double smem_leg[];
    const int K2 = n_theta >> 1;
    const int lda = K2 + 4;
    double* Ae_re = smem_leg;
    double* Ae_im = Ae_re + LEG_ROWS * lda;
    double* Ao_re = Ae_im + LEG_ROWS * lda;
    double* Ao_im = Ao_re + LEG_ROWS * lda;
    double* Be = Ao_im + LEG_ROWS * lda;
    double* Bo = Be + K2 * LEG_LDB;
    const int m = blockIdx.y;
    const int sh0 = blockIdx.x * 16;
    const int M2 = 2 * l_max + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- stage A: fold north/south, apply (-1)^m for the -m rows
    for (int item = tid; item < LEG_ROWS * K2; item += LEG_THREADS) {
        const int row = item / K2, j = item - row * K2;
        const int sh = sh0 + (row & 15), sign = row >> 4;
        double2 e = make_double2(0, 0), o = make_double2(0, 0);
        if (sh < S && (sign == 0 || m > 0)) {
            const int mm = sign ? (M2 - m) : m;
            const double2* src = a + ((size_t)sh * M2 + mm) * n_theta;
            const double2 x = ldg2(src + j), y = ldg2(src + (n_theta - 1 - j));
            e = cadd(x, y);
            o = csub(x, y);
            if (sign && (m & 1)) { e.x = -e.x; e.y = -e.y; o.x = -o.x; o.y = -o.y; }
        }
        Ae_re[row * lda + j] = e.x; Ae_im[row * lda + j] = e.y;
        Ao_re[row * lda + j] = o.x; Ao_im[row * lda + j] = o.y;
    }
    const double* FEm = FE + (size_t)m * K2 * NP;
    const double* FOm = FO + (size_t)m * K2 * NP;
    const int ne = (l_max - m) / 2 + 1;  // degrees l = m, m+2, ...
    for (int nb0 = 0; nb0 < ne; nb0 += LEG_NB) {
        __syncthreads();
        for (int item = tid; item < K2 * LEG_NB; item += LEG_THREADS) {
            const int j = item / LEG_NB, cc = item - j * LEG_NB;
            const bool ok = (nb0 + cc) < NP;
            Be[j * LEG_LDB + cc] = ok ? FEm[(size_t)j * NP + nb0 + cc] : 0.0;
            Bo[j * LEG_LDB + cc] = ok ? FOm[(size_t)j * NP + nb0 + cc] : 0.0;
        }
        __syncthreads();
        const int wn = warp & 3, r0 = (warp >> 2) * 16;
        const int no = (l_max - m + 1) / 2;              // degrees l = m+1, m+3, ...
        const bool do_e = nb0 + wn * 8 < ne, do_o = nb0 + wn * 8 < no;   // skip column blocks that are pure padding
        double ere[2][2] = {}, eim[2][2] = {}, ore_[2][2] = {}, oim[2][2] = {};
        if (do_e) leg_mma_cplx(Ae_re, Ae_im, lda, Be, LEG_LDB, wn * 8, r0, K2, lane, ere, eim);
        if (do_o) leg_mma_cplx(Ao_re, Ao_im, lda, Bo, LEG_LDB, wn * 8, r0, K2, lane, ore_, oim);
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            const int row = r0 + mb * 8 + (lane >> 2);
            const int sh = sh0 + (row & 15), sign = row >> 4;
            if (sh >= S || (sign && m == 0)) continue;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int col = nb0 + wn * 8 + 2 * (lane & 3) + cc;
                const int le = m + 2 * col, lo = le + 1;
                const int ms = sign ? -m : m;
                if (do_e && le <= l_max) c[(size_t)(le * (le + 1) + ms) * S + sh] = make_double2(ere[mb][cc], eim[mb][cc]);
                if (do_o && lo <= l_max) c[(size_t)(lo * (lo + 1) + ms) * S + sh] = make_double2(ore_[mb][cc], oim[mb][cc]);
            }
        }
    }
}

// pos_only: the field is real (c_{l,-m} = (-1)^m conj c_{l,m}), only the m >= 0 rows of `a` are produced (the phi-FFT
// completes the spectrum by conjugate symmetry) and a CTA takes 32 shells instead of 16 shells x (+m, -m).
__global__ void __launch_bounds__(LEG_THREADS, 4) legendre_inverse_kernel(const double2* __restrict__ c, double2* __restrict__ a,
                                                               const double* __restrict__ IE, const double* __restrict__ IO,
                                                               int S, int l_max, int n_theta, int NP, int pos_only) {
    extern __shared__ double smem_leg[];
    const int K2 = n_theta >> 1;
    const int lda = NP + 4;
    double* Ce_re = smem_leg;
    double* Ce_im = Ce_re + LEG_ROWS * lda;
    double* Co_re = Ce_im + LEG_ROWS * lda;
    double* Co_im = Co_re + LEG_ROWS * lda;
    double* Be = Co_im + LEG_ROWS * lda;
    double* Bo = Be + NP * LEG_LDB;
    const int m = blockIdx.y;
    const int sh0 = blockIdx.x * (pos_only ? 32 : 16);
    const int M2 = 2 * l_max + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- stage A: gather coefficients of order +-m, (-1)^m on the -m rows; shells fastest for coalescing
    for (int item = tid; item < 2 * NP * LEG_ROWS; item += LEG_THREADS) {
        const int rr = pos_only ? (item & 31) : (item & 15);
        int rest = item >> 4;
        const int sign = pos_only ? 0 : (rest & 1);
        rest >>= 1;
        const int i = rest % NP, par = rest / NP;
        const int l = m + par + 2 * i;
        const int sh = sh0 + rr;
        double2 v = make_double2(0, 0);
        if (l <= l_max && sh < S && (sign == 0 || m > 0)) {
            v = ldg2(c + (size_t)(l * (l + 1) + (sign ? -m : m)) * S + sh);
            if (sign && (m & 1)) { v.x = -v.x; v.y = -v.y; }
        }
        const int row = pos_only ? rr : sign * 16 + rr;
        if (par == 0) { Ce_re[row * lda + i] = v.x; Ce_im[row * lda + i] = v.y; }
        else          { Co_re[row * lda + i] = v.x; Co_im[row * lda + i] = v.y; }
    }
    const double* IEm = IE + (size_t)m * NP * K2;
    const double* IOm = IO + (size_t)m * NP * K2;
    for (int j0 = 0; j0 < K2; j0 += LEG_NB) {
        __syncthreads();
        for (int item = tid; item < NP * LEG_NB; item += LEG_THREADS) {
            const int i = item / LEG_NB, cc = item - i * LEG_NB;
            const bool ok = (j0 + cc) < K2;
            Be[i * LEG_LDB + cc] = ok ? IEm[(size_t)i * K2 + j0 + cc] : 0.0;
            Bo[i * LEG_LDB + cc] = ok ? IOm[(size_t)i * K2 + j0 + cc] : 0.0;
        }
        __syncthreads();
        const int wn = warp & 3, r0 = (warp >> 2) * 16;
        const int ne = (l_max - m) / 2 + 1, no = (l_max - m + 1) / 2;
        const int Ke = (ne + 3) & ~3, Ko = (no + 3) & ~3;   // contraction only over existing degrees (A is zero padded)
        double ere[2][2] = {}, eim[2][2] = {}, ore_[2][2] = {}, oim[2][2] = {};
        leg_mma_cplx(Ce_re, Ce_im, lda, Be, LEG_LDB, wn * 8, r0, Ke, lane, ere, eim);
        leg_mma_cplx(Co_re, Co_im, lda, Bo, LEG_LDB, wn * 8, r0, Ko, lane, ore_, oim);
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            const int row = r0 + mb * 8 + (lane >> 2);
            const int sh = sh0 + (pos_only ? row : (row & 15)), sign = pos_only ? 0 : (row >> 4);
            if (sh >= S || (sign && m == 0)) continue;
            const int mm = sign ? (M2 - m) : m;
            double2* dst = a + ((size_t)sh * M2 + mm) * n_theta;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int j = j0 + wn * 8 + 2 * (lane & 3) + cc;
                if (j < K2) {
                    dst[j] = make_double2(ere[mb][cc] + ore_[mb][cc], eim[mb][cc] + oim[mb][cc]);
                    dst[n_theta - 1 - j] = make_double2(ere[mb][cc] - ore_[mb][cc], eim[mb][cc] - oim[mb][cc]);
                }
            }
        }
    }
}

static inline size_t legendre_fwd_smem(int n_theta) {
    const int K2 = n_theta / 2;
    return (size_t)(4 * LEG_ROWS * (K2 + 4) + 2 * K2 * LEG_LDB) * sizeof(double);
}
static inline size_t legendre_inv_smem(int n_theta, int NP) {
    (void)n_theta;
    return (size_t)(4 * LEG_ROWS * (NP + 4) + 2 * NP * LEG_LDB) * sizeof(double);
}

// =====================================================================================================================
// Pipelined kernels (n_theta <= 128, NP <= 64: the L=63 / 64 x 128 workload of the bench and L=127 / 128 x 256).  One CTA
// works on ONE order m and walks over many groups of shells; the phi-Fourier rows (forward) / coefficient rows (inverse)
// of the NEXT groups are fetched with cp.async while the current group is multiplied, so HBM loads stay in flight all
// the time (the v1 kernels above stall on their one load phase per CTA: ncu long-scoreboard 6.9 of 16 cycles per
// issue).  The north/south fold is done on the fly when the A fragments are read (2 x 128-bit shared loads give
// e = x + y and o = x - y for re and im at once).
//   rows of a group: 8 shells x (+m, -m)   or, pos_only (real field: c_{l,-m} = (-1)^m conj c_{l,m} is redundant) and
//   for m = 0, 16 shells x (+m).
// =====================================================================================================================
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;                     // src-size 0: zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

#define LEG2_FR 16          // forward: rows per group (8 shells x +-m, or 16 shells)
#ifndef LEG2_FST
#define LEG2_FST 3          // forward: cp.async stages (2 groups in flight per CTA)
#endif
#define LEG2_IR 16          // inverse: rows per group
#ifndef LEG2_IST
#define LEG2_IST 3          // inverse: cp.async stages
#endif
#ifndef LEG3_MINB
#define LEG3_MINB 4         // CTAs per SM of the small instantiation (register budget 128)
#endif
#ifndef LEG3_WAVES
#define LEG3_WAVES 8        // CTAs per order m: this many waves of the resident CTAs over the whole grid (2 .. 32 measured: 3.72, 3.67, 3.62, 3.60, 3.62 ms)
#endif

// v3 forward: the table fragments live in REGISTERS and a warp owns two 8-column blocks.  The tables FE / FO of the CTA's order m are constant over its whole life, and an 8 x 4 x 8 DMMA takes
// its B operand one double per lane: K2 / 4 (<= 8) k-steps x 2 column blocks x 2 parities = 32 doubles per lane, loaded
// once from L2.  Per k-step a warp then reads only its two A rows (x at theta_j, y at the mirrored node: 2 x 128-bit
// shared loads) and issues 8 DMMAs, against 4 DMMAs per (2 x 128-bit + 2 x 64-bit) loads before: 1 instead of 3 shared
// wavefronts per DMMA, half the fold additions per DMMA, and no table staging in shared memory (52 KB per CTA).
//   64 NCG threads: warp = (row block of 8 rows) x (column-block pair cg < NCG: degrees m + 2 (16 cg + 0..15) [+1]);
//   KS = k-steps held in registers.  <KS 8, NCG 2>: n_theta <= 64, NP <= 32 (L = 63), 4 CTAs per SM;
//   <KS 16, NCG 4>: n_theta <= 128, NP <= 64 (L = 127), 64 table doubles per lane, 1 CTA of 8 warps per SM.
#ifndef LEG3_BIG_ST
#define LEG3_BIG_ST 3      // cp.async stages of the <KS 16, NCG 4> instantiation (one CTA per SM: shared memory allows up to 6)
#endif
static inline size_t legendre3_fwd_smem(int n_theta, int stages = LEG2_FST) { return (size_t)stages * LEG2_FR * (n_theta + 4) * sizeof(double2); }

template <int R, int ST, int KS, int NCG>
__global__ void __launch_bounds__(64 * NCG, KS <= 8 ? LEG3_MINB : 1) legendre3_forward_kernel(const double2* __restrict__ a, double2* __restrict__ c,
                                                                           const double* __restrict__ FE, const double* __restrict__ FO,
                                                                           int S, int l_max, int n_theta, int NP, int pos_only, long long c_stride) {
    // S: shells handled by this launch (rows of a);  c_stride: shells per coefficient row of c (>= S: a launch may cover
    // a chunk of the batch, c then points at the chunk's first shell)
    static_assert(R == 16, "two row blocks of 8 rows x NCG column-block pairs");
    constexpr int LEG3_THREADS = 64 * NCG;
    extern __shared__ __align__(16) unsigned char smem_leg2[];
    const int K2 = n_theta >> 1;
    const int RS = n_theta + 4;                            // row stride (double2): rows 64 B apart mod 128 -> conflict-free fragments
    double2* raw = reinterpret_cast<double2*>(smem_leg2);  // [ST][R][RS]
    const int m = blockIdx.y;
    const int M2 = 2 * l_max + 1;
    const bool both = (!pos_only) && m > 0;
    const int SH = both ? R / 2 : R;                       // shells per group
    const int n_groups = (S + SH - 1) / SH;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ar = lane >> 2, ak = lane & 3;
    const int cg = warp % NCG, r0 = (warp / NCG) * 8;
    const int ne = (l_max - m) / 2 + 1, no = (l_max - m + 1) / 2;
    bool do_e[2], do_o[2];
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) { do_e[nb] = (2 * cg + nb) * 8 < ne; do_o[nb] = (2 * cg + nb) * 8 < no; }

    // item = tid + q * LEG3_THREADS of the R x n_theta tile as (row, j): decomposed ONCE per thread, then advanced by the
    // constant step with a carry -- the runtime division per 16-byte copy was a third of the kernel's instructions (ncu:
    // issue slots 52 % busy with 4 % of the instructions being DMMAs)
    const int row_t0 = tid / n_theta, j_t0 = tid - row_t0 * n_theta;
    const int d_row = LEG3_THREADS / n_theta, d_j = LEG3_THREADS - d_row * n_theta;
    auto fetch = [&](int g, int buf) {
        if (g < n_groups) {
            double2* dst = raw + (size_t)buf * R * RS;
            const int sh0 = g * SH;
            int row = row_t0, j = j_t0;
            while (row < R) {
                const int sh = sh0 + (both ? (row % (R / 2)) : row), sign = both ? (row / (R / 2)) : 0;
                const bool ok = sh < S;
                const int mm = sign ? (M2 - m) : m;
                cp_async16(dst + row * RS + j, a + ((size_t)(ok ? sh : 0) * M2 + mm) * n_theta + j, ok);
                row += d_row; j += d_j;
                if (j >= n_theta) { j -= n_theta; ++row; }
            }
        }
        cp_async_commit();                                 // (possibly empty) group: keeps the wait count uniform
    };

    const int per_cta = (n_groups + gridDim.x - 1) / gridDim.x;
    int g = blockIdx.x * per_cta;
    const int g_end = min(n_groups, g + per_cta);
    if (g >= g_end) return;
#pragma unroll
    for (int s = 0; s < ST - 1; ++s) fetch(g + s < g_end ? g + s : n_groups, s);
    // table fragments: B[k][n] of k-step t is FE[m][4 t + ak][8 (2 cg + nb) + ar]
    double be[2][KS], bo[2][KS];
    {
        const double* FEm = FE + (size_t)m * K2 * NP;
        const double* FOm = FO + (size_t)m * K2 * NP;
#pragma unroll
        for (int t = 0; t < KS; ++t)
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) {
                const int col = (2 * cg + nb) * 8 + ar, j = 4 * t + ak;
                const bool ok = (j < K2) && (col < NP);
                be[nb][t] = (ok && do_e[nb]) ? __ldg(FEm + (size_t)j * NP + col) : 0.0;
                bo[nb][t] = (ok && do_o[nb]) ? __ldg(FOm + (size_t)j * NP + col) : 0.0;
            }
    }
    int buf = 0;
    for (; g < g_end; ++g) {
        cp_async_wait<ST - 2>();                           // the oldest outstanding group (this one) has landed
        __syncthreads();                                   // ... for every thread; and everyone is done with the buffer refilled next
        fetch(g + ST - 1 < g_end ? g + ST - 1 : n_groups, (buf + ST - 1) % ST);
        const double2* rr = raw + (size_t)buf * R * RS + (r0 + ar) * RS;
        double ere[2][2] = {}, eim[2][2] = {}, ore_[2][2] = {}, oim[2][2] = {};
        if (do_e[0]) {                                     // warp-uniform: a column-block pair that is pure padding has nothing to do
#pragma unroll
            for (int t = 0; t < KS; ++t) {
                if (4 * t < K2) {
                    const double2 x = rr[4 * t + ak], y = rr[n_theta - 1 - 4 * t - ak];
                    const double er = x.x + y.x, ei = x.y + y.y, orr = x.x - y.x, oi = x.y - y.y;
#pragma unroll
                    for (int nb = 0; nb < 2; ++nb) {
                        if (do_e[nb]) { dmma884(ere[nb][0], ere[nb][1], er, be[nb][t]); dmma884(eim[nb][0], eim[nb][1], ei, be[nb][t]); }
                        if (do_o[nb]) { dmma884(ore_[nb][0], ore_[nb][1], orr, bo[nb][t]); dmma884(oim[nb][0], oim[nb][1], oi, bo[nb][t]); }
                    }
                }
            }
            const int row = r0 + ar;
            const int sh = g * SH + (both ? (row % (R / 2)) : row), sign = both ? (row / (R / 2)) : 0;
            if (sh < S) {
                const double sg = (sign && (m & 1)) ? -1.0 : 1.0;      // (-1)^m on the -m rows
                const int ms = sign ? -m : m;
#pragma unroll
                for (int nb = 0; nb < 2; ++nb)
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        const int col = (2 * cg + nb) * 8 + 2 * ak + cc;
                        const int le = m + 2 * col, lo = le + 1;
                        if (do_e[nb] && le <= l_max) c[(size_t)(le * (le + 1) + ms) * c_stride + sh] = make_double2(sg * ere[nb][cc], sg * eim[nb][cc]);
                        if (do_o[nb] && lo <= l_max) c[(size_t)(lo * (lo + 1) + ms) * c_stride + sh] = make_double2(sg * ore_[nb][cc], sg * oim[nb][cc]);
                    }
            }
        }
        buf = (buf + 1) % ST;
    }
}


// v3 inverse (synthesis) counterpart: coefficients c [(L+1)^2][S] -> phi-Fourier rows a [S][M2][n_theta] for one order m per
// CTA; the coefficient rows of the next shell groups are gathered with cp.async (8 / 16 consecutive shells of one (l, +-m) row
// are contiguous).  Table fragments in registers, two 8-node blocks per warp (see legendre3_forward_kernel): per k-step one 128-bit
// shared load feeds 4 DMMAs.   128 threads: warp = (row block of 8 rows) x (node-block pair cg: theta_j, j = 16 cg + 0..15)
static inline size_t legendre3_inv_smem(int NP, int stages = LEG2_IST) { return (size_t)stages * 2 * LEG2_IR * (NP + 4) * sizeof(double2); }

template <int R, int ST, int KS, int NCG>
__global__ void __launch_bounds__(64 * NCG, KS <= 8 ? LEG3_MINB : 1) legendre3_inverse_kernel(const double2* __restrict__ c, double2* __restrict__ a,
                                                                           const double* __restrict__ IE, const double* __restrict__ IO,
                                                                           int S, int l_max, int n_theta, int NP, int pos_only, long long c_stride) {
    static_assert(R == 16, "two row blocks of 8 rows x NCG node-block pairs");
    constexpr int LEG3_THREADS = 64 * NCG;
    extern __shared__ __align__(16) unsigned char smem_leg2[];
    const int K2 = n_theta >> 1;
    const int RS = NP + 4;                                 // row stride (double2): rows 64 B apart mod 128
    double2* raw = reinterpret_cast<double2*>(smem_leg2);  // [ST][2 parities][R][RS]
    const int m = blockIdx.y;
    const int M2 = 2 * l_max + 1;
    const bool both = (!pos_only) && m > 0;
    const int SH = both ? R / 2 : R;
    const int n_groups = (S + SH - 1) / SH;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ar = lane >> 2, ak = lane & 3;
    const int cg = warp % NCG, r0 = (warp / NCG) * 8;
    const int ne = (l_max - m) / 2 + 1, no = (l_max - m + 1) / 2;
    const int Ke = (ne + 3) & ~3, Ko = (no + 3) & ~3;      // contraction only over existing degrees (zero padded)
    bool jok[2];
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) jok[nb] = (2 * cg + nb) * 8 < K2;

    // item = tid + q * LEG3_THREADS of the (parity, degree index, [sign], shell) tile, shells fastest (contiguous 16-byte elements of
    // one coefficient row).  SH and 2 SH divide the thread count, so the shell and the sign of a thread never change; the pair
    // (degree index i, parity) is decomposed once and advanced with a carry (no runtime division per copy).
    const int shl_t = tid % SH;
    const int rest_t = tid / SH;
    const int sign_t = both ? (rest_t & 1) : 0;
    const int rest_t2 = both ? (rest_t >> 1) : rest_t;
    const int i_t0 = rest_t2 % NP, par_t0 = rest_t2 / NP;
    const int d_rest = (LEG3_THREADS / SH) >> (both ? 1 : 0);
    const int d_i = d_rest % NP, d_par = d_rest / NP;
    const int row_t = sign_t * (R / 2) + shl_t;
    const int ms_t = sign_t ? -m : m;
    auto fetch = [&](int g, int buf) {
        if (g < n_groups) {
            double2* dst = raw + (size_t)buf * 2 * R * RS;
            const int sh = g * SH + shl_t;
            const bool sh_ok = sh < S;
            int i = i_t0, par = par_t0;
            while (par < 2) {
                const int l = m + par + 2 * i;
                const bool ok = sh_ok && (l <= l_max);
                cp_async16(dst + ((size_t)par * R + row_t) * RS + i, c + (size_t)(ok ? l * (l + 1) + ms_t : 0) * c_stride + (ok ? sh : 0), ok);
                i += d_i; par += d_par;
                if (i >= NP) { i -= NP; ++par; }
            }
        }
        cp_async_commit();
    };

    const int per_cta = (n_groups + gridDim.x - 1) / gridDim.x;
    int g = blockIdx.x * per_cta;
    const int g_end = min(n_groups, g + per_cta);
    if (g >= g_end) return;
#pragma unroll
    for (int s_ = 0; s_ < ST - 1; ++s_) fetch(g + s_ < g_end ? g + s_ : n_groups, s_);
    // table fragments: B[k][n] of k-step t is IE[m][4 t + ak][8 (2 cg + nb) + ar]   (degree index x northern node)
    double be[2][KS], bo[2][KS];
    {
        const double* IEm = IE + (size_t)m * NP * K2;
        const double* IOm = IO + (size_t)m * NP * K2;
#pragma unroll
        for (int t = 0; t < KS; ++t)
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) {
                const int j = (2 * cg + nb) * 8 + ar, i = 4 * t + ak;
                be[nb][t] = (i < Ke && i < NP && j < K2) ? __ldg(IEm + (size_t)i * K2 + j) : 0.0;
                bo[nb][t] = (i < Ko && i < NP && j < K2) ? __ldg(IOm + (size_t)i * K2 + j) : 0.0;
            }
    }
    int buf = 0;
    for (; g < g_end; ++g) {
        cp_async_wait<ST - 2>();
        __syncthreads();
        fetch(g + ST - 1 < g_end ? g + ST - 1 : n_groups, (buf + ST - 1) % ST);
        const double2* ce = raw + (size_t)buf * 2 * R * RS + (size_t)(r0 + ar) * RS;
        const double2* co = ce + (size_t)R * RS;
        double ere[2][2] = {}, eim[2][2] = {}, ore_[2][2] = {}, oim[2][2] = {};
        if (jok[0]) {
#pragma unroll
            for (int t = 0; t < KS; ++t) {
                if (4 * t < Ke) {
                    const double2 x = ce[4 * t + ak];
#pragma unroll
                    for (int nb = 0; nb < 2; ++nb)
                        if (jok[nb]) { dmma884(ere[nb][0], ere[nb][1], x.x, be[nb][t]); dmma884(eim[nb][0], eim[nb][1], x.y, be[nb][t]); }
                }
                if (4 * t < Ko) {
                    const double2 x = co[4 * t + ak];
#pragma unroll
                    for (int nb = 0; nb < 2; ++nb)
                        if (jok[nb]) { dmma884(ore_[nb][0], ore_[nb][1], x.x, bo[nb][t]); dmma884(oim[nb][0], oim[nb][1], x.y, bo[nb][t]); }
                }
            }
            const int row = r0 + ar;
            const int sh = g * SH + (both ? (row % (R / 2)) : row), sign = both ? (row / (R / 2)) : 0;
            if (sh < S) {
                const double sg = (sign && (m & 1)) ? -1.0 : 1.0;      // (-1)^m on the -m rows
                const int mm = sign ? (M2 - m) : m;
                double2* dst = a + ((size_t)sh * M2 + mm) * n_theta;
                // a lane holds the nodes j, j + 1 (j even): one 256-bit store for the northern pair and one for the mirrored
                // southern pair (n_theta - 2 - j, n_theta - 1 - j); n_theta is a multiple of 8, so both are 32-byte aligned
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) {
                    const int j = (2 * cg + nb) * 8 + 2 * ak;
                    if (jok[nb] && j < K2) {
                        st_global_256(dst + j, make_double2(sg * (ere[nb][0] + ore_[nb][0]), sg * (eim[nb][0] + oim[nb][0])),
                                      make_double2(sg * (ere[nb][1] + ore_[nb][1]), sg * (eim[nb][1] + oim[nb][1])));
                        st_global_256(dst + n_theta - 2 - j, make_double2(sg * (ere[nb][1] - ore_[nb][1]), sg * (eim[nb][1] - oim[nb][1])),
                                      make_double2(sg * (ere[nb][0] - ore_[nb][0]), sg * (eim[nb][0] - oim[nb][0])));
                    }
                }
            }
        }
        buf = (buf + 1) % ST;
    }
}
