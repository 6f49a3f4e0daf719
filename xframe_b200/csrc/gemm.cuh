// FP64 tensor-core (DMMA.8x8x4) GEMM kernels:
//   hankel_kernel      -- the radial (Hankel) contraction: complex rows x real per-order matrix
//                         out[row][k] = (-+i)^l * scale * sum_p in[row][p+skip] * W_l[p][k]
//                         reference: apply_weights OpenCL kernel / generate_spherical_ht
//                         (hankel_transforms.py:642-766), restated in oracle/mtip.py:generate_spherical_ht_direct
//   grouped_gemm_kernel-- small real GEMMs with arbitrary strides for the Procrustes step
//                         (PD_l @ I_l and V_l @ unk_l, fxs_Projections.py:752-767,835-841)
#pragma once
#include "common.cuh"

#define HK_BM 64
#define HK_BN 64
#define HK_BK 16
#define HK_LDA (HK_BK + 4)
#define HK_LDB (HK_BN + 4)

struct HankelTile {
    int l;        // index of the weight matrix (3-D: order l; 2-D: DFT index j of the order m)
    int row0;     // first flat row (lm*nb + b) of this tile
    int row_end;  // one past the last flat row of this order
    int ph;       // order mod 4 (non-negative): the prefactor is (-i)^ph forward, (+i)^ph inverse
};

__global__ void __launch_bounds__(256) hankel_kernel(const double2* __restrict__ in, double2* __restrict__ out,
                                                     const double* __restrict__ W, const HankelTile* __restrict__ tiles,
                                                     int n_r, int n_sum, int skip, double scale, int inverse) {
    __shared__ double As_re[HK_BM * HK_LDA];
    __shared__ double As_im[HK_BM * HK_LDA];
    __shared__ double Bs[HK_BK * HK_LDB];
    const HankelTile t = tiles[blockIdx.x];
    const int k_tile0 = blockIdx.y * HK_BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps, warp tile 32 rows x 16 cols
    const double* Wl = W + (size_t)t.l * n_sum * n_r;
    double cre[4][2][2] = {}, cim[4][2][2] = {};
    const int ar = lane >> 2, ak = lane & 3;

    for (int p0 = 0; p0 < n_sum; p0 += HK_BK) {
        __syncthreads();
        // A chunk: 64 rows x 16 p (complex), 16 consecutive threads read 256 contiguous bytes
#pragma unroll
        for (int q = 0; q < (HK_BM * HK_BK) / 256; ++q) {
            const int item = tid + q * 256;
            const int row = item >> 4, pp = item & 15;
            const int grow = t.row0 + row;
            double2 v = make_double2(0, 0);
            if (grow < t.row_end && (p0 + pp) < n_sum) v = ldg2(in + (size_t)grow * n_r + skip + p0 + pp);
            As_re[row * HK_LDA + pp] = v.x;
            As_im[row * HK_LDA + pp] = v.y;
        }
        // B chunk: 16 p x 64 k
#pragma unroll
        for (int q = 0; q < (HK_BK * HK_BN) / 256; ++q) {
            const int item = tid + q * 256;
            const int pp = item >> 6, kk = item & 63;
            double v = 0.0;
            if ((p0 + pp) < n_sum && (k_tile0 + kk) < n_r) v = __ldg(Wl + (size_t)(p0 + pp) * n_r + k_tile0 + kk);
            Bs[pp * HK_LDB + kk] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k0 = 0; k0 < HK_BK; k0 += 4) {
            double b[2];
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) b[nb] = Bs[(k0 + ak) * HK_LDB + wn * 16 + nb * 8 + ar];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) {
                const int off = (wm * 32 + mb * 8 + ar) * HK_LDA + k0 + ak;
                const double xr = As_re[off], xi = As_im[off];
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) {
                    dmma884(cre[mb][nb][0], cre[mb][nb][1], xr, b[nb]);
                    dmma884(cim[mb][nb][0], cim[mb][nb][1], xi, b[nb]);
                }
            }
        }
    }
    // epilogue: multiply by scale * (-i)^l (forward) or (+i)^l (inverse)
    const int ph = t.ph;
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
        const int grow = t.row0 + wm * 32 + mb * 8 + ar;
        if (grow >= t.row_end) continue;
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
            const int k = k_tile0 + wn * 16 + nb * 8 + 2 * ak;
            double2 o[2];
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const double x = cre[mb][nb][cc] * scale, y = cim[mb][nb][cc] * scale;
                double2 r;
                if (ph == 0) r = make_double2(x, y);
                else if (ph == 2) r = make_double2(-x, -y);
                else if ((ph == 1) != (inverse != 0)) r = make_double2(y, -x);   // multiply by -i
                else r = make_double2(-y, x);                                     // multiply by +i
                o[cc] = r;
            }
            if (k < n_r) out[(size_t)grow * n_r + k] = o[0];
            if (k + 1 < n_r) out[(size_t)grow * n_r + k + 1] = o[1];
        }
    }
}

// v2: same contraction, operands fetched with cp.async into a 3-stage ring (A rows as interleaved complex: one 128-bit
// shared load gives re and im of a fragment element), so the DMMA pipe is not idle while the next K chunk is loaded.
// Needs N_r even (16-byte chunks of the real weight rows); the v1 kernel covers odd sizes.
#define HK2_ST 3
#define HK2_LDA (HK_BK + 4)      // double2 units: rows 64 B apart mod 128 -> conflict-free fragment loads
#define HK2_LDB (HK_BN + 4)      // doubles
static inline size_t hankel2_smem() { return (size_t)HK2_ST * (HK_BM * HK2_LDA * sizeof(double2) + HK_BK * HK2_LDB * sizeof(double)); }
__device__ __forceinline__ void hk_cp_async16(void* smem, const void* gmem, bool valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(bytes));
}
// ldw: leading dimension of the weight rows (>= n_r, even); accumulate: out += result (second pass of a complex matrix,
// used by the 2-D DFT: X (C -+ i S) = X C -+ i (X S)).
#ifndef HK2_MINB
#define HK2_MINB 2           // CTAs per SM (register budget): 1 / 2 / 3 measured at 3.21 / 2.42 / 2.78 ms per step
#endif
__global__ void __launch_bounds__(256, HK2_MINB) hankel2_kernel(const double2* __restrict__ in, double2* __restrict__ out,
                                                         const double* __restrict__ W, const HankelTile* __restrict__ tiles,
                                                         int n_r, int n_sum, int skip, double scale, int inverse, int ldw, int accumulate) {
    extern __shared__ __align__(16) unsigned char smem_hk[];
    double2* As = reinterpret_cast<double2*>(smem_hk);                                     // [ST][HK_BM][HK2_LDA]
    double* Bs = reinterpret_cast<double*>(As + HK2_ST * HK_BM * HK2_LDA);                 // [ST][HK_BK][HK2_LDB]
    const HankelTile t = tiles[blockIdx.x];
    const int k_tile0 = blockIdx.y * HK_BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps, warp tile 32 rows x 16 cols
    const double* Wl = W + (size_t)t.l * n_sum * ldw;
    double cre[4][2][2] = {}, cim[4][2][2] = {};
    const int ar = lane >> 2, ak = lane & 3;
    const int n_chunks = (n_sum + HK_BK - 1) / HK_BK;

    auto fetch = [&](int ch, int st) {
        if (ch < n_chunks) {
            const int p0 = ch * HK_BK;
            double2* a = As + (size_t)st * HK_BM * HK2_LDA;
            double* b = Bs + (size_t)st * HK_BK * HK2_LDB;
#pragma unroll
            for (int q = 0; q < (HK_BM * HK_BK) / 256; ++q) {      // A: 64 rows x 16 p complex, 16 consecutive threads = 256 contiguous bytes
                const int item = tid + q * 256;
                const int row = item >> 4, pp = item & 15;
                const int grow = t.row0 + row;
                const bool ok = grow < t.row_end && (p0 + pp) < n_sum;
                hk_cp_async16(a + row * HK2_LDA + pp, in + (size_t)(ok ? grow : t.row0) * n_r + skip + (ok ? p0 + pp : 0), ok);
            }
#pragma unroll
            for (int q = 0; q < (HK_BK * HK_BN / 2) / 256; ++q) {  // B: 16 p x 64 k doubles as 16-byte pairs
                const int item = tid + q * 256;
                const int pp = item >> 5, kk = (item & 31) * 2;
                const bool ok = (p0 + pp) < n_sum && (k_tile0 + kk) < n_r;
                hk_cp_async16(b + pp * HK2_LDB + kk, Wl + (size_t)(ok ? p0 + pp : 0) * ldw + (ok ? k_tile0 + kk : 0), ok);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
#pragma unroll
    for (int s_ = 0; s_ < HK2_ST - 1; ++s_) fetch(s_, s_);
    int st = 0;
    for (int ch = 0; ch < n_chunks; ++ch) {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(HK2_ST - 2));
        __syncthreads();
        fetch(ch + HK2_ST - 1, (st + HK2_ST - 1) % HK2_ST);
        const double2* a = As + (size_t)st * HK_BM * HK2_LDA;
        const double* b_ = Bs + (size_t)st * HK_BK * HK2_LDB;
#pragma unroll
        for (int k0 = 0; k0 < HK_BK; k0 += 4) {
            double b[2];
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) b[nb] = b_[(k0 + ak) * HK2_LDB + wn * 16 + nb * 8 + ar];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) {
                const double2 x = a[(wm * 32 + mb * 8 + ar) * HK2_LDA + k0 + ak];
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) {
                    dmma884(cre[mb][nb][0], cre[mb][nb][1], x.x, b[nb]);
                    dmma884(cim[mb][nb][0], cim[mb][nb][1], x.y, b[nb]);
                }
            }
        }
        st = (st + 1) % HK2_ST;
    }
    const int ph = t.ph;
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
        const int grow = t.row0 + wm * 32 + mb * 8 + ar;
        if (grow >= t.row_end) continue;
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
            const int k = k_tile0 + wn * 16 + nb * 8 + 2 * ak;
            double2 o[2];
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const double x = cre[mb][nb][cc] * scale, y = cim[mb][nb][cc] * scale;
                double2 r;
                if (ph == 0) r = make_double2(x, y);
                else if (ph == 2) r = make_double2(-x, -y);
                else if ((ph == 1) != (inverse != 0)) r = make_double2(y, -x);   // multiply by -i
                else r = make_double2(-y, x);                                     // multiply by +i
                o[cc] = r;
            }
            if (accumulate) {
                if (k < n_r) { const double2 c = out[(size_t)grow * n_r + k]; o[0].x += c.x; o[0].y += c.y; }
                if (k + 1 < n_r) { const double2 c = out[(size_t)grow * n_r + k + 1]; o[1].x += c.x; o[1].y += c.y; }
            }
            if (k < n_r) out[(size_t)grow * n_r + k] = o[0];
            if (k + 1 < n_r) out[(size_t)grow * n_r + k + 1] = o[1];
        }
    }
}

// v3: the same contraction with the operand tiles fed by the TMA engine (bulk asynchronous copies, cp.async.bulk ->
// SASS UBLKCP) that signal an mbarrier per pipeline stage: ONE warp issues 64 + 16 row copies per K chunk (a 256-byte run of an
// input row, a 512-byte run of a weight row, each to its padded shared-memory row) instead of 256 threads issuing six 16-byte
// cp.async each, and the consumers wait on the stage's mbarrier phase instead of cp.async.wait_group.  Bulk copies cannot
// zero-fill, so this kernel takes only full K chunks and full column tiles (n_sum % 16 == 0, N_r % 64 == 0: the L=63 / N_r=128
// and L=127 / N_r=256 workloads); rows past the end of an order are not copied -- a GEMM row only feeds its own output row,
// which is not stored.  Same accumulation order as hankel2_kernel: bit-identical results.
__device__ __forceinline__ unsigned hk_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hk_mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hk_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void hk_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hk_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hk_mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                     : "=r"(done) : "r"(hk_smem_addr(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void hk_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(hk_smem_addr(smem_dst)), "l"(gmem_src),
                 "r"(bytes), "r"(hk_smem_addr(bar)) : "memory");
}
static inline size_t hankel3_smem() { return hankel2_smem() + 64; }
__global__ void __launch_bounds__(256, HK2_MINB) hankel3_tma_kernel(const double2* __restrict__ in, double2* __restrict__ out,
                                                             const double* __restrict__ W, const HankelTile* __restrict__ tiles,
                                                             int n_r, int n_sum, int skip, double scale, int inverse, int ldw, int accumulate) {
    extern __shared__ __align__(16) unsigned char smem_hk[];
    double2* As = reinterpret_cast<double2*>(smem_hk);                                     // [ST][HK_BM][HK2_LDA]
    double* Bs = reinterpret_cast<double*>(As + HK2_ST * HK_BM * HK2_LDA);                 // [ST][HK_BK][HK2_LDB]
    uint64_t* full = reinterpret_cast<uint64_t*>(Bs + HK2_ST * HK_BK * HK2_LDB);           // [ST] one mbarrier per stage
    const HankelTile t = tiles[blockIdx.x];
    const int k_tile0 = blockIdx.y * HK_BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps, warp tile 32 rows x 16 cols
    const double* Wl = W + (size_t)t.l * n_sum * ldw;
    double cre[4][2][2] = {}, cim[4][2][2] = {};
    const int ar = lane >> 2, ak = lane & 3;
    const int n_chunks = n_sum / HK_BK;
    const int rows_valid = min(HK_BM, t.row_end - t.row0);
    if (tid == 0) {
#pragma unroll
        for (int s_ = 0; s_ < HK2_ST; ++s_) hk_mbar_init(full + s_, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int ch, int st) {            // warp 0: arm the stage's barrier with the byte count, then the row copies
        if (warp == 0 && ch < n_chunks) {
            const int p0 = ch * HK_BK;
            double2* a = As + (size_t)st * HK_BM * HK2_LDA;
            double* b = Bs + (size_t)st * HK_BK * HK2_LDB;
            if (lane == 0) hk_mbar_expect_tx(full + st, (unsigned)(rows_valid * HK_BK * sizeof(double2) + HK_BK * HK_BN * sizeof(double)));
            __syncwarp();
            for (int i = lane; i < HK_BM + HK_BK; i += 32) {
                if (i < HK_BM) {
                    if (i < rows_valid)
                        hk_bulk_g2s(a + i * HK2_LDA, in + (size_t)(t.row0 + i) * n_r + skip + p0, HK_BK * sizeof(double2), full + st);
                } else {
                    const int pp = i - HK_BM;
                    hk_bulk_g2s(b + pp * HK2_LDB, Wl + (size_t)(p0 + pp) * ldw + k_tile0, HK_BN * sizeof(double), full + st);
                }
            }
        }
    };
#pragma unroll
    for (int s_ = 0; s_ < HK2_ST - 1; ++s_) issue(s_, s_);
    int st = 0;
    for (int ch = 0; ch < n_chunks; ++ch) {
        hk_mbar_wait(full + st, (unsigned)((ch / HK2_ST) & 1));      // this chunk's bytes have landed
        __syncthreads();                                             // everyone is done with the stage refilled next
        issue(ch + HK2_ST - 1, (st + HK2_ST - 1) % HK2_ST);
        const double2* a = As + (size_t)st * HK_BM * HK2_LDA;
        const double* b_ = Bs + (size_t)st * HK_BK * HK2_LDB;
#pragma unroll
        for (int k0 = 0; k0 < HK_BK; k0 += 4) {
            double b[2];
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) b[nb] = b_[(k0 + ak) * HK2_LDB + wn * 16 + nb * 8 + ar];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) {
                const double2 x = a[(wm * 32 + mb * 8 + ar) * HK2_LDA + k0 + ak];
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) {
                    dmma884(cre[mb][nb][0], cre[mb][nb][1], x.x, b[nb]);
                    dmma884(cim[mb][nb][0], cim[mb][nb][1], x.y, b[nb]);
                }
            }
        }
        st = (st + 1) % HK2_ST;
    }
    const int ph = t.ph;
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
        const int grow = t.row0 + wm * 32 + mb * 8 + ar;
        if (grow >= t.row_end) continue;
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
            const int k = k_tile0 + wn * 16 + nb * 8 + 2 * ak;
            double2 o[2];
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const double x = cre[mb][nb][cc] * scale, y = cim[mb][nb][cc] * scale;
                double2 r;
                if (ph == 0) r = make_double2(x, y);
                else if (ph == 2) r = make_double2(-x, -y);
                else if ((ph == 1) != (inverse != 0)) r = make_double2(y, -x);   // multiply by -i
                else r = make_double2(-y, x);                                     // multiply by +i
                o[cc] = r;
            }
            if (accumulate) {
                const double2 c0 = out[(size_t)grow * n_r + k], c1 = out[(size_t)grow * n_r + k + 1];
                o[0].x += c0.x; o[0].y += c0.y; o[1].x += c1.x; o[1].y += c1.y;
            }
            st_global_256(out + (size_t)grow * n_r + k, o[0], o[1]);            // k even, N_r % 64 == 0: 32-byte aligned pair
        }
    }
}

// Hankel transform evaluated at the first output radius only: out0[row] = (-+i)^l scale sum_p in[row][p+skip] W_l[p][0].
// Used by the fused ft_stab step: only shell 0 of IFT(rho_hat) is needed.  One warp per row.
__global__ void hankel_row0_kernel(const double2* __restrict__ in, double2* __restrict__ out0, const double* __restrict__ W, int n_rows,
                                   int nb, int n_r, int n_sum, int skip, double scale, int inverse) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const int lm = row / nb;
    int l = (int)sqrt((double)lm);
    while ((l + 1) * (l + 1) <= lm) ++l;
    while (l * l > lm) --l;
    const double* Wl = W + (size_t)l * n_sum * n_r;
    double sx = 0.0, sy = 0.0;
    for (int p = lane; p < n_sum; p += 32) {
        const double2 v = ldg2(in + (size_t)row * n_r + skip + p);
        const double w = __ldg(Wl + (size_t)p * n_r);
        sx += v.x * w; sy += v.y * w;
    }
    sx = warp_sum(sx) * scale; sy = warp_sum(sy) * scale;
    if (lane == 0) {
        const int ph = l & 3;
        double2 r;
        if (ph == 0) r = make_double2(sx, sy);
        else if (ph == 2) r = make_double2(-sx, -sy);
        else if ((ph == 1) != (inverse != 0)) r = make_double2(sy, -sx);
        else r = make_double2(-sy, sx);
        out0[row] = r;
    }
}

// ---------------------------------------------------------------------------
// grouped real GEMM  C = alpha * A * B, arbitrary element strides, one
// descriptor per problem; 64x64 block tiles, 4 warps (2x2), warp tile 32x32.
// ---------------------------------------------------------------------------
struct GemmProblem {
    const double* A;
    const double* B;
    double* C;
    int M, N, K;
    long long a_rs, a_cs;  // element strides of A[m][k]
    long long b_rs, b_cs;  // B[k][n]
    long long c_rs, c_cs;  // C[m][n]
    double alpha;
    int tile0;  // first tile id of this problem in the launch
    int tiles_n;
};

#define GG_BM 64
#define GG_BN 64
#define GG_BK 16
#define GG_LDA (GG_BK + 4)
#define GG_LDB (GG_BN + 4)

// tile_base: first tile of this launch in the (run-major) tile list -- a launch may cover the runs [b0, b0 + n) of the batch
__global__ void __launch_bounds__(128) grouped_gemm_kernel(const GemmProblem* __restrict__ probs, const int* __restrict__ tile_prob, int tile_base) {
    __shared__ double As[GG_BM * GG_LDA];
    __shared__ double Bs[GG_BK * GG_LDB];
    const int tile = blockIdx.x + tile_base;
    const GemmProblem pr = probs[tile_prob[tile]];
    const int tl = tile - pr.tile0;
    const int m0 = (tl / pr.tiles_n) * GG_BM, n0 = (tl % pr.tiles_n) * GG_BN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    const int ar = lane >> 2, ak = lane & 3;
    double acc[4][4][2] = {};
    const bool a_kfast = (pr.a_cs == 1);
    const bool b_nfast = (pr.b_cs == 1);
    for (int k0 = 0; k0 < pr.K; k0 += GG_BK) {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < (GG_BM * GG_BK) / 128; ++q) {
            const int item = tid + q * 128;
            int mm, kk;
            if (a_kfast) { mm = item >> 4; kk = item & 15; } else { kk = item >> 6; mm = item & 63; }
            double v = 0.0;
            if ((m0 + mm) < pr.M && (k0 + kk) < pr.K) v = __ldg(pr.A + (m0 + mm) * pr.a_rs + (k0 + kk) * pr.a_cs);
            As[mm * GG_LDA + kk] = v;
        }
#pragma unroll
        for (int q = 0; q < (GG_BK * GG_BN) / 128; ++q) {
            const int item = tid + q * 128;
            int kk, nn;
            if (b_nfast) { kk = item >> 6; nn = item & 63; } else { nn = item >> 4; kk = item & 15; }
            double v = 0.0;
            if ((k0 + kk) < pr.K && (n0 + nn) < pr.N) v = __ldg(pr.B + (k0 + kk) * pr.b_rs + (n0 + nn) * pr.b_cs);
            Bs[kk * GG_LDB + nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kq = 0; kq < GG_BK; kq += 4) {
            double a[4], b[4];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) a[mb] = As[(wm * 32 + mb * 8 + ar) * GG_LDA + kq + ak];
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) b[nb] = Bs[(kq + ak) * GG_LDB + wn * 32 + nb * 8 + ar];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
        }
    }
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
        const int m = m0 + wm * 32 + mb * 8 + ar;
        if (m >= pr.M) continue;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int n = n0 + wn * 32 + nb * 8 + 2 * ak + cc;
                if (n < pr.N) pr.C[m * pr.c_rs + n * pr.c_cs] = pr.alpha * acc[mb][nb][cc];
            }
        }
    }
}
