// Per-order projection of the intensity harmonic coefficients I_l onto the
// measured invariants (orthogonal Procrustes):  I'_l = V_l * polar(V_l^T D^2 I_l)
// reference: approximate_unknowns + mtip_projection (fxs_Projections.py:752-872),
// restated in oracle/mtip.py:ReciprocalProjection.
//
// Formulation used on the device ("Procrustes"):
//  * |rho_hat|^2 is real, so I_{l,-m} = (-1)^m conj(I_{l,m}).  With the unitary change of
//    basis to real spherical harmonics the problem becomes REAL: X = I_l T,
//    M = PD_l X (n_l x (2l+1)), polar(M_complex) = polar(M) T^H.
//  * polar(M) = U V^T is computed with a one-sided Jacobi SVD of G = M^T held in shared
//    memory (rotating the n_l columns of G, each of length 2l+1).  The accumulated rotations
//    are applied to V_l instead of the identity, so the kernel directly yields V_l*U and
//    V*Sigma; then T = (V_l U)(V)^T is one more small GEMM.
//  * Columns whose norm falls below sv_cutoff * max norm are singular directions the
//    reference's LAPACK SVD cannot resolve either (it returns an arbitrary orthonormal
//    completion there); they are dropped (zero completion).
#pragma once
#include "common.cuh"

struct ProcOrder {
    int l;        // harmonic order
    int n_cols;   // n_l: columns of V_l
    int n_c;      // 2l+1
    long long pd_off;   // into pd / vt constant arrays (n_cols * N_r doubles)
    long long xt_off;   // per-run offset of Xt / Tt block (n_c * N_r)
    long long g_off;    // per-run offset of G / Gn block (n_cols * jacobi_stride(n_c), zero padded columns)
    long long vw_off;   // per-run offset of VW block (n_cols * jacobi_wstride(N_r), zero padded columns)
};

enum { ORD_PASS = 0, ORD_ZERO = 1, ORD_ZEROTH = 2, ORD_ACTIVE = 3 };

// c [(L+1)^2][S] complex -> Xt[b][order][m'][k] real, m'=0: Re c_{l0}; 2m-1: sqrt2 Re c_{lm}; 2m: sqrt2 Im c_{lm}
__global__ void procrustes_pack_kernel(const double2* __restrict__ c, double* __restrict__ xt, const ProcOrder* __restrict__ orders,
                                       int n_r, int S, long long xt_run_stride) {
    const ProcOrder o = orders[blockIdx.x];
    const int b = blockIdx.y;
    const double s2 = 1.4142135623730951;
    double* dst = xt + (size_t)b * xt_run_stride + o.xt_off;
    for (int idx = threadIdx.x; idx < (o.l + 1) * n_r; idx += blockDim.x) {
        const int m = idx / n_r, k = idx - m * n_r;
        const double2 v = ldg2(c + (size_t)(o.l * (o.l + 1) + m) * S + (size_t)b * n_r + k);
        if (m == 0) {
            dst[k] = v.x;
        } else {
            dst[(size_t)(2 * m - 1) * n_r + k] = s2 * v.x;
            dst[(size_t)(2 * m) * n_r + k] = s2 * v.y;
        }
    }
}

// One-sided Jacobi SVD, persistent CTAs over the (order, run) problems (largest orders first).
//
// Storage (all padded with zeros so every access is a 128-bit one):
//   G : n columns (column i of G = M^T = row i of M) of length len <= ldg, ldg = 64 / 128 / 256.  Input g [n][ldg]
//       (written by the grouped GEMM, pads zeroed once at allocation), working copy in shared memory when it fits,
//       otherwise in the output buffer gn [n][ldg] (global memory, L2 resident).
//   W : accumulator, n columns of length wld = 128 / 256 (>= N_r), vw [n][wld] in global memory, initialised from
//       vt (V_l^T, or the identity for xfb_get_unknowns).  At the start of every sweep the ACTIVE columns are staged
//       into shared memory when they fit (dead columns are dropped as the sweeps proceed) and written back afterwards.
//
// Register-blocked sweep: the active columns are grouped in blocks of 4; blocks are paired by the round-robin (circle)
// method and one WARP owns a block pair (A, B) for a whole block round.  The 8 columns are loaded ONCE, the 16 cross
// pairs (A_g, B_(g+s)%4), s = 0..3, are rotated in registers (8 lanes per pair, 4 pairs in flight per warp, the B
// columns travel between the lane groups by shuffles) and the columns are stored ONCE: shared-memory traffic per
// rotation is 1/4 of the pair-at-a-time scheme and there is one __syncthreads per block round (n/4 per sweep) instead
// of one per pair round.  In block round 0 the 6 + 6 pairs inside A and inside B are rotated too, so every pair is
// visited exactly once per sweep.  The rotation parameters are kept per warp and replayed on the accumulator columns
// (second pass: W needs no dot products).
#define JG 8                 // lanes per column pair
#define JROT_STEPS 10        // 3 (inside A) + 3 (inside B) + 4 (cross)
__host__ __device__ inline int jacobi_stride(int len) { return len <= 64 ? 64 : (len <= 128 ? 128 : 256); }
__host__ __device__ inline int jacobi_wstride(int n_r) { return n_r <= 128 ? 128 : 256; }

__device__ __forceinline__ double2 shfl_d2(double2 v, int src_lane) {
    v.x = __shfl_sync(0xffffffffu, v.x, src_lane);
    v.y = __shfl_sync(0xffffffffu, v.y, src_lane);
    return v;
}

// rotation of the column pair (x, y) held by one 8-lane group: (cs, sn) with x' = cs x - sn y, y' = sn x + cs y, and
// tapq = tan(theta) * (x.y), the change of the squared norms (|x'|^2 = |x|^2 - tapq, |y'|^2 = |y|^2 + tapq);
// (1, 0, 0) when the pair is skipped.  The squared norms app, aqq are CACHED by the caller (computed when the columns
// are loaded, updated after every rotation), so a pair costs one dot product instead of three.  Symmetric under
// exchanging the roles of x and y (sn and tapq change sign), so two lane groups that hold the same pair with swapped
// roles take bitwise-consistent decisions.
struct JRot { double cs, sn, tapq; };
template <int NV>
__device__ __forceinline__ JRot jacobi_pair_params(const double2 (&x)[NV], const double2 (&y)[NV], double app, double aqq, bool valid,
                                                   double thr, double tol) {
    double apq = 0.0, apq2 = 0.0;
#pragma unroll
    for (int t = 0; t < NV; ++t) { apq += x[t].x * y[t].x; apq2 += x[t].y * y[t].y; }
    apq += apq2;
#pragma unroll
    for (int off = JG / 2; off > 0; off >>= 1) apq += __shfl_xor_sync(0xffffffffu, apq, off);
    bool rot = valid && (app > thr) && (aqq > thr);
    if (rot) rot = apq * apq > (tol * tol) * (app * aqq);
    if (!rot) return JRot{1.0, 0.0, 0.0};
    // tan(2 theta) = 2 apq / (aqq - app), |theta| <= pi/4 (same rotation as the textbook zeta/t form)
    const double d = aqq - app, s2 = 2.0 * apq;
    const double rh = rsqrt(d * d + s2 * s2);
    const double u = 0.5 + 0.5 * fabs(d) * rh;            // cos^2(theta)
    const double rc = rsqrt(u);                            // 1 / cos(theta)
    const double sn = copysign(0.5 * s2 * rh, d * s2) * rc;   // sin(2 theta) / (2 cos theta)
    return JRot{u * rc, sn, sn * rc * apq};
}

template <int NV>
__device__ __forceinline__ double jacobi_col_norm2(const double2 (&x)[NV]) {
    double a = 0.0, a2 = 0.0;
#pragma unroll
    for (int t = 0; t < NV; ++t) { a += x[t].x * x[t].x; a2 += x[t].y * x[t].y; }
    a += a2;
#pragma unroll
    for (int off = JG / 2; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    return a;
}

template <int NV>
__device__ __forceinline__ void jacobi_load_col(double2 (&v)[NV], const double* base, long long col_off, int sub, bool valid) {
    const double2* p = reinterpret_cast<const double2*>(base + col_off) + sub;
#pragma unroll
    for (int t = 0; t < NV; ++t) v[t] = valid ? p[JG * t] : make_double2(0.0, 0.0);
}
template <int NV>
__device__ __forceinline__ void jacobi_store_col(const double2 (&v)[NV], double* base, long long col_off, int sub, bool valid) {
    if (!valid) return;
    double2* p = reinterpret_cast<double2*>(base + col_off) + sub;
#pragma unroll
    for (int t = 0; t < NV; ++t) p[JG * t] = v[t];
}

// Block pair i of block round r (circle method on nblkp blocks: position 0 is fixed, position k >= 1 holds block
// 1 + ((k-1-r) mod (nblkp-1))); pairs are (position i, position nblkp-1-i).
struct JBlockPair { int ba, bb; };
__device__ __forceinline__ JBlockPair jacobi_block_pair(int i, int r, int nblkp) {
    // 0 <= i < nblkp / 2 and 0 <= r <= mod, so both differences lie in (-2 mod, mod): two conditional additions instead of
    // an integer modulo (the division sat on the critical path of every round)
    const int mod = nblkp - 1;
    JBlockPair bp;
    bp.ba = 0;
    if (i != 0) { int t_ = i - 1 - r; if (t_ < 0) t_ += mod; if (t_ < 0) t_ += mod; bp.ba = 1 + t_; }
    int tb = mod - 1 - i - r; if (tb < 0) tb += mod; if (tb < 0) tb += mod;
    bp.bb = 1 + tb;
    return bp;
}

// Pass 1 of a block pair (one warp): load the 8 G columns, rotate the 16 cross pairs (and in round 0 the 6 + 6 pairs
// inside A and B) in registers, store the columns, leave the rotation parameters in rb[JROT_STEPS][4] and
// rb[JROT_STEPS*4].x = 1 if anything rotated.
template <int NV2>
__device__ __forceinline__ void jacobi_g_pass(double* Gs, int ldg, const int* __restrict__ list, int nact, JBlockPair bp, bool intra,
                                              double thr, double tol, double2* rb, double* nrm2) {
    const int lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & (JG - 1);
    const int ba = bp.ba, bb = bp.bb;
    const int pa = ba * 4 + grp;                                   // position of A_g in the active list
    const bool va = pa < nact;
    const int ca = va ? list[pa] : 0;
    const int pb0 = bb * 4 + grp, pb3 = bb * 4 + ((grp + 3) & 3);
    const bool vb0 = pb0 < nact, vb3 = pb3 < nact;
    bool any_rot = false;
    double2 x[NV2], y[NV2];
    jacobi_load_col<NV2>(x, Gs, (long long)ca * ldg, sub, va);
    jacobi_load_col<NV2>(y, Gs, (long long)(vb0 ? list[pb0] : 0) * ldg, sub, vb0);
    // squared norms: computed once per sweep (jacobi_active_list), then carried in nrm2 [column] and updated by every
    // rotation (|x'|^2 = |x|^2 - tan(theta) x.y), as LAPACK's one-sided Jacobi does; they travel with the columns
    const int cb0 = vb0 ? list[pb0] : 0;
    double nx = va ? nrm2[ca] : 0.0, ny = vb0 ? nrm2[cb0] : 0.0;
    if (intra) {
#pragma unroll
        for (int t = 0; t < 3; ++t) {                             // pairs inside A: partner group = grp ^ (t+1)
            const bool vp = (ba * 4 + (grp ^ (t + 1))) < nact;
            double2 z[NV2];
#pragma unroll
            for (int k = 0; k < NV2; ++k) z[k] = shfl_d2(x[k], lane ^ (8 * (t + 1)));
            const double nz = __shfl_sync(0xffffffffu, nx, lane ^ (8 * (t + 1)));
            const JRot R = jacobi_pair_params<NV2>(x, z, nx, nz, va && vp, thr, tol);
            if (R.sn != 0.0) {
                any_rot = true;
#pragma unroll
                for (int k = 0; k < NV2; ++k) x[k] = make_double2(R.cs * x[k].x - R.sn * z[k].x, R.cs * x[k].y - R.sn * z[k].y);
                nx -= R.tapq;
            }
            if (sub == 0) rb[t * 4 + grp] = make_double2(R.cs, R.sn);
        }
#pragma unroll
        for (int t = 0; t < 3; ++t) {                             // pairs inside B
            const bool vp = (bb * 4 + (grp ^ (t + 1))) < nact;
            double2 z[NV2];
#pragma unroll
            for (int k = 0; k < NV2; ++k) z[k] = shfl_d2(y[k], lane ^ (8 * (t + 1)));
            const double nz = __shfl_sync(0xffffffffu, ny, lane ^ (8 * (t + 1)));
            const JRot R = jacobi_pair_params<NV2>(y, z, ny, nz, vb0 && vp, thr, tol);
            if (R.sn != 0.0) {
                any_rot = true;
#pragma unroll
                for (int k = 0; k < NV2; ++k) y[k] = make_double2(R.cs * y[k].x - R.sn * z[k].x, R.cs * y[k].y - R.sn * z[k].y);
                ny -= R.tapq;
            }
            if (sub == 0) rb[(3 + t) * 4 + grp] = make_double2(R.cs, R.sn);
        }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {                                 // cross pairs (A_g, B_(g+s)%4)
        const bool vb = (bb * 4 + ((grp + s) & 3)) < nact;
        const JRot R = jacobi_pair_params<NV2>(x, y, nx, ny, va && vb, thr, tol);
        if (R.sn != 0.0) {
            any_rot = true;
#pragma unroll
            for (int k = 0; k < NV2; ++k) {
                const double2 a = x[k], b = y[k];
                x[k] = make_double2(R.cs * a.x - R.sn * b.x, R.cs * a.y - R.sn * b.y);
                y[k] = make_double2(R.sn * a.x + R.cs * b.x, R.sn * a.y + R.cs * b.y);
            }
            nx -= R.tapq; ny += R.tapq;
        }
        if (sub == 0) rb[(6 + s) * 4 + grp] = make_double2(R.cs, R.sn);
        if (s < 3) {
#pragma unroll
            for (int k = 0; k < NV2; ++k) y[k] = shfl_d2(y[k], (lane + 8) & 31);       // B columns move to the previous group
            ny = __shfl_sync(0xffffffffu, ny, (lane + 8) & 31);
        }
    }
    jacobi_store_col<NV2>(x, Gs, (long long)ca * ldg, sub, va);
    const int cb3 = vb3 ? list[pb3] : 0;
    jacobi_store_col<NV2>(y, Gs, (long long)cb3 * ldg, sub, vb3);
    if (sub == 0) {
        if (va) nrm2[ca] = nx;
        if (vb3) nrm2[cb3] = ny;
    }
    any_rot = __any_sync(0xffffffffu, any_rot);
    if (lane == 0) rb[JROT_STEPS * 4] = make_double2(any_rot ? 1.0 : 0.0, 0.0);
}

// Pass 2 of a block pair (one warp): replay the recorded rotations on the accumulator columns.  Returns true if
// anything was rotated.  Wb columns are indexed by the POSITION in the active list when w_compact (shared-memory
// staging) or by the column id otherwise.
template <int WV2>
__device__ __forceinline__ bool jacobi_w_pass(double* Wb, int wld, bool w_compact, const int* __restrict__ list, int nact, JBlockPair bp,
                                              bool intra, const double2* rb) {
    if (rb[JROT_STEPS * 4].x == 0.0) return false;                // warp-uniform
    const int lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & (JG - 1);
    const int pa = bp.ba * 4 + grp, pb0 = bp.bb * 4 + grp, pb3 = bp.bb * 4 + ((grp + 3) & 3);
    const bool va = pa < nact, vb0 = pb0 < nact, vb3 = pb3 < nact;
    const long long wa = (long long)(w_compact ? pa : (va ? list[pa] : 0)) * wld;
    const long long wb0 = (long long)(w_compact ? pb0 : (vb0 ? list[pb0] : 0)) * wld;
    const long long wb3 = (long long)(w_compact ? pb3 : (vb3 ? list[pb3] : 0)) * wld;
    double2 x[WV2], y[WV2];
    jacobi_load_col<WV2>(x, Wb, wa, sub, va);
    jacobi_load_col<WV2>(y, Wb, wb0, sub, vb0);
    if (intra) {
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const double2 cs = rb[t * 4 + grp];
            double2 z[WV2];
#pragma unroll
            for (int k = 0; k < WV2; ++k) z[k] = shfl_d2(x[k], lane ^ (8 * (t + 1)));
            if (cs.y != 0.0) {
#pragma unroll
                for (int k = 0; k < WV2; ++k) x[k] = make_double2(cs.x * x[k].x - cs.y * z[k].x, cs.x * x[k].y - cs.y * z[k].y);
            }
        }
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const double2 cs = rb[(3 + t) * 4 + grp];
            double2 z[WV2];
#pragma unroll
            for (int k = 0; k < WV2; ++k) z[k] = shfl_d2(y[k], lane ^ (8 * (t + 1)));
            if (cs.y != 0.0) {
#pragma unroll
                for (int k = 0; k < WV2; ++k) y[k] = make_double2(cs.x * y[k].x - cs.y * z[k].x, cs.x * y[k].y - cs.y * z[k].y);
            }
        }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const double2 cs = rb[(6 + s) * 4 + grp];
        if (cs.y != 0.0) {
#pragma unroll
            for (int k = 0; k < WV2; ++k) {
                const double2 a = x[k], b = y[k];
                x[k] = make_double2(cs.x * a.x - cs.y * b.x, cs.x * a.y - cs.y * b.y);
                y[k] = make_double2(cs.y * a.x + cs.x * b.x, cs.y * a.y + cs.x * b.y);
            }
        }
        if (s < 3) {
#pragma unroll
            for (int k = 0; k < WV2; ++k) y[k] = shfl_d2(y[k], (lane + 8) & 31);
        }
    }
    jacobi_store_col<WV2>(x, Wb, wa, sub, va);
    jacobi_store_col<WV2>(y, Wb, wb3, sub, vb3);
    return true;
}

// One sweep over all pairs of the `nact` active columns list[0..nact).  rotbuf: [2][JROT_SLOTS][JROT_RB] double2.
//  * enough block pairs for every warp: each warp runs pass 1 then pass 2 of its block pairs, one barrier per round;
//  * fewer block pairs than warps: the warps split into a G team (pass 1 of round r) and a W team (pass 2 of round
//    r-1, from the double-buffered rotation parameters) that run concurrently -- G and W are independent data.
#define JROT_RB (JROT_STEPS * 4 + 1)
// block pairs per round that need a parameter slot (kernel template parameter SLOTS): n <= 128 columns -> 16, n <= 256 -> 32,
// n <= 64 (the two-CTAs-per-SM variant for the small orders) -> 8
template <int NV2, int WV2>
__device__ __forceinline__ void jacobi_sweep_blocked(double* Gs, int ldg, double* Wb, int wld, bool w_compact, const int* __restrict__ list,
                                                     int nact, double thr, double tol, double2* rotbuf, int* s_rot, double* nrm2, int JROT_SLOTS) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int nblk = (nact + 3) >> 2;
    const int nblkp = max(2, nblk + (nblk & 1));
    const int half = nblkp >> 1, rounds = nblkp - 1;
    bool rotated = false;
    if (4 * half > 3 * nwarp || nwarp < 4) {
        for (int r = 0; r < rounds; ++r) {
            for (int i = warp; i < half; i += nwarp) {                     // warp-uniform
                const JBlockPair bp = jacobi_block_pair(i, r, nblkp);
                double2* rb = rotbuf + (size_t)(i % JROT_SLOTS) * JROT_RB;
                jacobi_g_pass<NV2>(Gs, ldg, list, nact, bp, r == 0, thr, tol, rb, nrm2);
                __syncwarp();
                rotated |= jacobi_w_pass<WV2>(Wb, wld, w_compact, list, nact, bp, r == 0, rb);
                __syncwarp();
            }
            __syncthreads();
        }
    } else {
        const int ng = half, nw = nwarp - half;                            // ng <= 3/4 nwarp
        for (int r = 0; r <= rounds; ++r) {
            if (warp < ng) {
                if (r < rounds)
                    jacobi_g_pass<NV2>(Gs, ldg, list, nact, jacobi_block_pair(warp, r, nblkp), r == 0, thr, tol,
                                       rotbuf + (size_t)((r & 1) * JROT_SLOTS + warp) * JROT_RB, nrm2);
            } else if (r > 0) {
                for (int i = warp - ng; i < half; i += nw)
                    rotated |= jacobi_w_pass<WV2>(Wb, wld, w_compact, list, nact, jacobi_block_pair(i, r - 1, nblkp), r == 1,
                                                  rotbuf + (size_t)(((r - 1) & 1) * JROT_SLOTS + i) * JROT_RB);
            }
            __syncthreads();
        }
    }
    if (rotated && lane == 0) *s_rot = 1;
}

// runtime column stride -> compile-time register tile size.  The accumulator columns have the same length as the G
// columns (the accumulator starts as the identity or as Q2, see the kernel); MAXV2 bounds the instantiations.
// `len` <= ld is the number of leading column entries that can be non-zero (the rest is zero padding): the register tiles
// cover ceil(len / 16) double2 per lane only.
template <int MAXV2>
__device__ __forceinline__ void jacobi_sweep_dispatch(double* Gs, double* Wb, int ld, int len, bool w_compact, const int* list, int nact,
                                                      double thr, double tol, double2* rotbuf, int* s_rot, double* nrm2, int slots) {
    if constexpr (MAXV2 >= 16) {
        if (len > 128) { jacobi_sweep_blocked<16, 16>(Gs, ld, Wb, ld, w_compact, list, nact, thr, tol, rotbuf, s_rot, nrm2, slots); return; }
    }
    if (len <= 32) jacobi_sweep_blocked<2, 2>(Gs, ld, Wb, ld, w_compact, list, nact, thr, tol, rotbuf, s_rot, nrm2, slots);
    else if (len <= 48) jacobi_sweep_blocked<3, 3>(Gs, ld, Wb, ld, w_compact, list, nact, thr, tol, rotbuf, s_rot, nrm2, slots);
    else if (len <= 64) jacobi_sweep_blocked<4, 4>(Gs, ld, Wb, ld, w_compact, list, nact, thr, tol, rotbuf, s_rot, nrm2, slots);
    else if constexpr (MAXV2 >= 8) {
        if (len <= 80) jacobi_sweep_blocked<5, 5>(Gs, ld, Wb, ld, w_compact, list, nact, thr, tol, rotbuf, s_rot, nrm2, slots);
        else if (len <= 96) jacobi_sweep_blocked<6, 6>(Gs, ld, Wb, ld, w_compact, list, nact, thr, tol, rotbuf, s_rot, nrm2, slots);
        else jacobi_sweep_blocked<8, 8>(Gs, ld, Wb, ld, w_compact, list, nact, thr, tol, rotbuf, s_rot, nrm2, slots);
    }
}

__device__ __forceinline__ void jacobi_col_norms(const double* Gs, int ldg, int n, double* nrm2);

// QR factorisation by modified Gram-Schmidt with re-orthogonalisation ("twice is enough"), whole CTA.
//   A [r][lda]: r columns (zero padded), overwritten by the orthonormal Q (a column that vanishes becomes 0);
//   Rt [r][ldr]: Rt[j][k] = R[j][k] (k >= j), i.e. column j of R^T -- the column storage of the NEXT factorisation step.
// Step j: q_j is final; every later column k gets a_k -= (q_j.a_k) q_j (one 8-lane group per column, on registers); the
// group that owns column j+1 normalises it in the same step, so there is one barrier per step.
// Selective re-orthogonalisation (Daniel-Gragg-Kaufman-Stewart criterion per projection): the projection is repeated only
// where the first pass removed more than half of the column's squared norm.  The squared norms are cached in n2c [r]
// and downdated (|a - c q|^2 = |a|^2 - c^2); nref [r] keeps the high word of the last freshly computed value, and a column
// whose downdated norm fell below 1e-6 of it is recomputed (the downdate has lost its digits by then).  On the fxs
// problems 1 - 20 % of the projections take the second pass; the factors agree with the always-twice variant to 1e-16
// (tests/jacobi_model.py:mgs2).
template <int NV>
__device__ __forceinline__ void mgs2_qr(double* A, int lda, int r, double* Rt, int ldr, double* n2c, int* nref) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int grp = lane >> 3, sub = lane & (JG - 1);
    for (int i = tid; i < r * ldr; i += blockDim.x) Rt[i] = 0.0;
    jacobi_col_norms(A, lda, r, n2c);
    __syncthreads();
    for (int i = tid; i < r; i += blockDim.x) nref[i] = __double2hiint(n2c[i]);
    if (warp == 0) {                                       // normalise column 0 (group 0 does the work, all lanes shuffle)
        double2 x[NV];
        jacobi_load_col<NV>(x, A, 0, sub, grp == 0);
        const double n2 = jacobi_col_norm2<NV>(x), inv = (n2 > 0.0) ? rsqrt(n2) : 0.0, nrm = n2 * inv;
#pragma unroll
        for (int t = 0; t < NV; ++t) { x[t].x *= inv; x[t].y *= inv; }
        jacobi_store_col<NV>(x, A, 0, sub, grp == 0);
        if (lane == 0) Rt[0] = nrm;
    }
    for (int j = 0; j < r; ++j) {
        __syncthreads();                                   // q_j is final
        for (int k0 = j + 1 + warp * 4; k0 < r; k0 += nwarp * 4) {      // warp-uniform trip count
            const int k = k0 + grp;
            const bool v = k < r;
            const bool first = (k0 == j + 1);              // warp-uniform: this task holds the next pivot column (group 0)
            double2 q[NV], a[NV];
            jacobi_load_col<NV>(q, A, (long long)j * lda, sub, true);
            jacobi_load_col<NV>(a, A, (long long)(v ? k : j) * lda, sub, v);
            const double n2k = v ? n2c[k] : 0.0;
            const double n2ref = v ? __hiloint2double(nref[k], 0) : 0.0;
            double c = 0.0, c2 = 0.0;
#pragma unroll
            for (int t = 0; t < NV; ++t) { c += q[t].x * a[t].x; c2 += q[t].y * a[t].y; }
            c += c2;
#pragma unroll
            for (int off = JG / 2; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
#pragma unroll
            for (int t = 0; t < NV; ++t) { a[t].x -= c * q[t].x; a[t].y -= c * q[t].y; }
            double c_tot = c;
            double n2n = n2k - c * c;                      // |a - c q|^2 for a unit q
            // second pass only where the first one cancelled more than half of the squared norm (or the downdated norm
            // has lost its digits); decided per warp, a superfluous second pass is harmless
            const bool slow = __any_sync(0xffffffffu, v && (c * c > 0.5 * n2k || n2n < 1e-6 * n2ref));
            if (slow) {
                c = 0.0; c2 = 0.0;
#pragma unroll
                for (int t = 0; t < NV; ++t) { c += q[t].x * a[t].x; c2 += q[t].y * a[t].y; }
                c += c2;
#pragma unroll
                for (int off = JG / 2; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
#pragma unroll
                for (int t = 0; t < NV; ++t) { a[t].x -= c * q[t].x; a[t].y -= c * q[t].y; }
                c_tot += c;
            }
            const bool fresh = slow || first;
            if (fresh) n2n = jacobi_col_norm2<NV>(a);
            if (first && grp == 0) {                       // the first trailing column becomes q_{j+1} right away
                const double inv = (n2n > 0.0) ? rsqrt(n2n) : 0.0, nrm = n2n * inv;     // one MUFU + Newton instead of sqrt and a division
#pragma unroll
                for (int t = 0; t < NV; ++t) { a[t].x *= inv; a[t].y *= inv; }
                if (sub == 0) Rt[(size_t)(j + 1) * ldr + j + 1] = nrm;
            }
            jacobi_store_col<NV>(a, A, (long long)k * lda, sub, v);
            if (v && sub == 0) {
                Rt[(size_t)j * ldr + k] = c_tot;
                n2c[k] = n2n;
                if (fresh) nref[k] = __double2hiint(n2n);
            }
        }
    }
    __syncthreads();
}
template <int MAXV2>
__device__ __forceinline__ void mgs2_qr_dispatch(double* A, int lda, int len, int r, double* Rt, int ldr, double* n2c, int* nref) {
    if (len <= 32) mgs2_qr<2>(A, lda, r, Rt, ldr, n2c, nref);
    else if (len <= 48) mgs2_qr<3>(A, lda, r, Rt, ldr, n2c, nref);
    else if (len <= 64) mgs2_qr<4>(A, lda, r, Rt, ldr, n2c, nref);
    else if constexpr (MAXV2 >= 8) {
        if (len <= 80) mgs2_qr<5>(A, lda, r, Rt, ldr, n2c, nref);
        else if (len <= 96) mgs2_qr<6>(A, lda, r, Rt, ldr, n2c, nref);
        else if (len <= 112) mgs2_qr<7>(A, lda, r, Rt, ldr, n2c, nref);
        else mgs2_qr<8>(A, lda, r, Rt, ldr, n2c, nref);
    }
}

// column norms of Gs -> nrm2, ordered list of the columns above the cut-off -> list, their number -> *s_nact
__device__ __forceinline__ void jacobi_active_list(const double* Gs, int ldg, int n, double* nrm2, int* list, double sv_cutoff, int* s_nact,
                                                   double* s_thr, int* s_rot);

// squared norms of the n columns (one 8-lane group per column)
__device__ __forceinline__ void jacobi_col_norms(const double* Gs, int ldg, int n, double* nrm2) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int grp = lane >> 3, sub = lane & (JG - 1);
    for (int c0 = warp * 4; c0 < n; c0 += nwarp * 4) {                     // warp-uniform trip count (shuffles need all lanes)
        const int c = c0 + grp;
        double s = 0.0;
        if (c < n) {
            const double2* pc = reinterpret_cast<const double2*>(Gs + (size_t)c * ldg) + sub;
            for (int t = 0; t < ldg / (2 * JG); ++t) { const double2 v = pc[JG * t]; s += v.x * v.x + v.y * v.y; }
        }
#pragma unroll
        for (int off = JG / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (sub == 0 && c < n) nrm2[c] = s;
    }
}

__device__ __forceinline__ void jacobi_active_list(const double* Gs, int ldg, int n, double* nrm2, int* list, double sv_cutoff, int* s_nact,
                                                   double* s_thr, int* s_rot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    jacobi_col_norms(Gs, ldg, n, nrm2);
    __syncthreads();
    if (warp == 0) {
        double mx = 0.0;
        for (int c = lane; c < n; c += 32) mx = fmax(mx, nrm2[c]);
        mx = warp_max(mx);
        const double thr = sv_cutoff * sv_cutoff * mx;
        int base = 0;                              // ordered compaction of the active columns
        for (int c0 = 0; c0 < n; c0 += 32) {
            const int c = c0 + lane;
            const bool act = (c < n) && (nrm2[c] > thr);
            const unsigned bal = __ballot_sync(0xffffffffu, act);
            if (act) list[base + __popc(bal & ((1u << lane) - 1u))] = c;
            base += __popc(bal);
        }
        if (lane == 0) { *s_nact = base; *s_thr = thr; *s_rot = 0; }
    }
    __syncthreads();
}

// Phase timing of the QR-preconditioned path (diagnostics, compiled in with -DJAC_TIMING only): cycles summed over all problems
//   0 active list + gather, 1 first QR, 2 second QR, 3 sweeps, 4 final norms + polar product, 5 problems, 6 sweeps run
__device__ unsigned long long g_jac_phase[8];
#ifdef JAC_TIMING
#define JAC_T0() long long jt_ = clock64()
#define JAC_TICK(i) do { __syncthreads(); if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_jac_phase[i], (unsigned long long)(t_ - jt_)); jt_ = t_; } } while (0)
#define JAC_COUNT(i, v) do { if (threadIdx.x == 0) atomicAdd(&g_jac_phase[i], (unsigned long long)(v)); } while (0)
#else
#define JAC_T0()
#define JAC_TICK(i)
#define JAC_COUNT(i, v)
#endif

// QR-preconditioned polar factor of one problem (Drmac-Veselic style preconditioning of the one-sided Jacobi SVD):
//   G_a = Q1 R1 (active columns only, len x r);  R1^T = Q2 R2;  L = R2^T;  Jacobi on the columns of L with the same
//   rotations applied to the columns of Q2:  L J = U~ Sigma,  W = Q2 J  =>  polar(R1) = U~ W^T =: P,  polar(G_a) = Q1 P.
//   Outputs:  gn[list[a]] = column a of Q1 (rows of inactive columns are zero),  pp[list[i]][list[a]] = P[i][a].
// The Jacobi then works on an r x r lower-triangular, well-graded matrix: ~7 sweeps instead of ~11 and ~2.5x fewer
// rotations on columns of length r instead of 2l+1 (numpy model: tests/jacobi_model.py:qr_polar).
// Shared memory: A = Q1 [r][ldg], B [r][ldl], C [r][ldl].
template <int MAXV2>
__device__ __forceinline__ int jacobi_qr_problem(const double* __restrict__ g, int n, int ldg, int len_g, double* __restrict__ gn, double* __restrict__ pp,
                                                 double* nrm2, int* list, int r, int ldl, double* region, double2* rotbuf, double sv_cutoff,
                                                 double tol, int max_sweeps, int* s_nact, double* s_thr, int* s_rot,
                                                 double* __restrict__ sigma, int slots) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    double* A = region;                 // [r][ldg]
    double* B = A + (size_t)r * ldg;    // [r][ldl]
    double* C = B + (size_t)r * ldl;    // [r][ldl]
    int* list2 = list + n;              // active columns of L (the caller reserved 2n ints)
    JAC_T0();
    // ---- gather the active columns, zero the outputs (rows / entries of inactive columns stay zero)
    for (int i = tid; i < r * (ldg / 2); i += nthr) {
        const int a = i / (ldg / 2), e = i - a * (ldg / 2);
        reinterpret_cast<double2*>(A)[i] = __ldg(reinterpret_cast<const double2*>(g + (size_t)list[a] * ldg) + e);
    }
    for (int i = tid; i < n * (ldg / 2); i += nthr) {
        reinterpret_cast<double2*>(gn)[i] = make_double2(0.0, 0.0);
        reinterpret_cast<double2*>(pp)[i] = make_double2(0.0, 0.0);
    }
    __syncthreads();
    JAC_TICK(0);
    mgs2_qr_dispatch<MAXV2>(A, ldg, len_g, r, B, ldl, nrm2, list2);     // A = Q1, B = columns of R1^T (nrm2 / list2 are free until the sweeps)
    for (int i = tid; i < r * (ldg / 2); i += nthr) {     // Q1 is final: rows list[a] of gn
        const int a = i / (ldg / 2), e = i - a * (ldg / 2);
        reinterpret_cast<double2*>(gn + (size_t)list[a] * ldg)[e] = reinterpret_cast<const double2*>(A)[i];
    }
    JAC_TICK(1);
    mgs2_qr_dispatch<MAXV2>(B, ldl, r, r, C, ldl, nrm2, list2);         // B = Q2, C = columns of R2^T = L
    JAC_TICK(2);
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        jacobi_active_list(C, ldl, r, nrm2, list2, sv_cutoff, s_nact, s_thr, s_rot);
        const int nact = *s_nact;
        const double thr = *s_thr;
        if (nact < 2) break;
        jacobi_sweep_dispatch<MAXV2>(C, B, ldl, r, false, list2, nact, thr, tol, rotbuf, s_rot, nrm2, slots);
        __syncthreads();
        const int rotated = *s_rot;
        __syncthreads();
        if (!rotated) { ++sweep; break; }
    }
    __syncthreads();
    JAC_TICK(3);
    jacobi_active_list(C, ldl, r, nrm2, list2, sv_cutoff, s_nact, s_thr, s_rot);     // final norms
    const double thr_f = *s_thr;
    for (int i = tid; i < r; i += nthr) sigma[i] = sqrt(nrm2[i]);
    // ---- U~ = C / sigma on the resolved directions (0 on the dropped ones)
    for (int i = tid; i < r * ldl; i += nthr) {
        const int c = i / ldl;
        const double s2 = nrm2[c];
        C[i] = (s2 > thr_f && s2 > 0.0) ? C[i] * rsqrt(s2) : 0.0;
    }
    __syncthreads();
    // ---- P[i][a] = sum_c U~[c][i] W[c][a]  ->  pp[list[i]][list[a]];  each thread a 2 x 4 tile (a fastest); rows / columns
    //      beyond r read the zero padding (ldl >= r rounded up to 16)
    const int aq = (r + 3) / 4, iq = (r + 1) / 2;
    for (int item = tid; item < iq * aq; item += nthr) {
        const int i = (item / aq) * 2, a4 = (item - (item / aq) * aq) * 4;
        double acc00 = 0.0, acc01 = 0.0, acc02 = 0.0, acc03 = 0.0, acc10 = 0.0, acc11 = 0.0, acc12 = 0.0, acc13 = 0.0;
        for (int c = 0; c < r; ++c) {
            const double2 u = *reinterpret_cast<const double2*>(C + (size_t)c * ldl + i);
            const double2 b01 = *reinterpret_cast<const double2*>(B + (size_t)c * ldl + a4);
            const double2 b23 = *reinterpret_cast<const double2*>(B + (size_t)c * ldl + a4 + 2);
            acc00 += u.x * b01.x; acc01 += u.x * b01.y; acc02 += u.x * b23.x; acc03 += u.x * b23.y;
            acc10 += u.y * b01.x; acc11 += u.y * b01.y; acc12 += u.y * b23.x; acc13 += u.y * b23.y;
        }
        double* dst = pp + (size_t)list[i] * ldg;
        if (a4 < r) dst[list[a4]] = acc00;
        if (a4 + 1 < r) dst[list[a4 + 1]] = acc01;
        if (a4 + 2 < r) dst[list[a4 + 2]] = acc02;
        if (a4 + 3 < r) dst[list[a4 + 3]] = acc03;
        if (i + 1 < r) {
            dst = pp + (size_t)list[i + 1] * ldg;
            if (a4 < r) dst[list[a4]] = acc10;
            if (a4 + 1 < r) dst[list[a4 + 1]] = acc11;
            if (a4 + 2 < r) dst[list[a4 + 2]] = acc12;
            if (a4 + 3 < r) dst[list[a4 + 3]] = acc13;
        }
    }
    __syncthreads();
    JAC_TICK(4);
    JAC_COUNT(5, 1);
    JAC_COUNT(6, sweep);
    return sweep;
}

// Kernel outputs, per (run, order) problem:  gn [n][ldg], pp [n][ldg] with  polar(G) = sum_c gn[c] (x) pp[c]
// (len x n); the host then forms  vw = pp V_l^T  and  T^T = gn^T vw  with two grouped DMMA GEMMs.
//   QR path:        gn = Q1 (rows of the active columns), pp = P scattered to the active rows / columns
//   direct path:    gn = U~ (normalised rotated columns of G), pp = J^T (accumulator started from the identity)
// MAXV2 = 8, THREADS = 512: column length <= 128 (the L=63 configuration);  MAXV2 = 16, THREADS = 256: up to 256
// (L=127), 255 registers per thread for the 2 x 16 double2 register tiles.
// MAXV2 = 4, THREADS = 256, MINB = 2, SLOTS = 8: the orders with 2l+1 <= 64 (column length and rank <= 64) run TWO problems per SM
// (half the shared memory each): the kernel is bound by the latency of its sequential rotation steps with a third of the
// issue slots used, so a second resident problem fills the idle cycles.
template <int MAXV2, int THREADS, bool QR, int MINB, int SLOTS>
__global__ void __launch_bounds__(THREADS, MINB) procrustes_jacobi_kernel(const double* __restrict__ g_in, double* __restrict__ gn_out,
                                                                       double* __restrict__ pp_out, double* __restrict__ sigma_out,
                                                                       const ProcOrder* __restrict__ orders, int n_orders, int n_batch,
                                                                       int sig_ld, long long g_run_stride, long long sig_run_stride,
                                                                       double sv_cutoff, double tol, int max_sweeps,
                                                                       int* __restrict__ sweeps_out, int smem_doubles,
                                                                       int* __restrict__ work_counter, int order0, int n_orders_all) {
    // orders [order0, order0 + n_orders) of the plan's list (largest first) are this launch's problems; sigma / sweeps are
    // indexed by the position in the whole list (n_orders_all entries per run)
    extern __shared__ __align__(16) double smem_j[];
    __shared__ int s_nact, s_rot, s_prob;
    __shared__ double s_thr;
    const int tid = threadIdx.x;
    double2* rotbuf = reinterpret_cast<double2*>(smem_j);                 // [2][slots][JROT_RB]
    const int fixed = 2 * 2 * SLOTS * JROT_RB;                            // doubles

    // dynamic work queue (largest problems first): a CTA fetches the next problem when it is done with the previous one
    for (;;) {
        __syncthreads();
        if (tid == 0) s_prob = atomicAdd(work_counter, 1);
        __syncthreads();
        const int prob = s_prob;
        if (prob >= n_orders * n_batch) break;
        const int oi = order0 + prob / n_batch;            // orders are sorted largest first
        const int b = prob - (oi - order0) * n_batch;
        const ProcOrder o = orders[oi];
        const int n = o.n_cols;
        const int ldg = jacobi_stride(o.n_c);
        double* nrm2 = smem_j + fixed;                     // [n]
        int* list = (int*)(nrm2 + n);                      // [2n]
        const int var0 = (fixed + 2 * n + 2) & ~1;                        // 16-byte aligned
        const double* g = g_in + (size_t)b * g_run_stride + o.g_off;
        double* gn = gn_out + (size_t)b * g_run_stride + o.g_off;
        double* pp = pp_out + (size_t)b * g_run_stride + o.g_off;
        double* sg = sigma_out + (size_t)b * sig_run_stride + (size_t)oi * sig_ld;
        if constexpr (QR) {
            // QR-preconditioned path when the three r-sized arrays fit in shared memory
            jacobi_active_list(g, ldg, n, nrm2, list, sv_cutoff, &s_nact, &s_thr, &s_rot);
            const int r = s_nact;
            const int ldl = r <= 64 ? 64 : (r <= 96 ? 96 : 128);
            if (r >= 2 && var0 + r * ldg + 2 * r * ldl <= smem_doubles) {
                __syncthreads();
                const int sw = jacobi_qr_problem<MAXV2>(g, n, ldg, o.n_c, gn, pp, nrm2, list, r, ldl, smem_j + var0, rotbuf, sv_cutoff, tol, max_sweeps,
                                                        &s_nact, &s_thr, &s_rot, sg, SLOTS);
                if (tid == 0 && sweeps_out) sweeps_out[b * n_orders_all + oi] = sw;
                continue;
            }
            __syncthreads();
        }
        // ---- direct path: Jacobi on G itself, accumulator W = J^T in pp (global), staged per sweep for the active columns
        const bool g_smem = var0 + n * ldg <= smem_doubles;
        double* Gs = g_smem ? smem_j + var0 : gn;          // [n][ldg]
        double* Ws = smem_j + var0 + (g_smem ? n * ldg : 0);              // [cap][ldg]
        const int cap = (smem_doubles - (int)(Ws - smem_j)) / ldg;
        {
            const double2* src = reinterpret_cast<const double2*>(g);
            double2* dst = reinterpret_cast<double2*>(Gs);
            for (int i = tid; i < n * ldg / 2; i += THREADS) dst[i] = src[i];
        }
        for (int i = tid; i < n * ldg; i += THREADS) {
            const int c = i / ldg, e = i - c * ldg;
            pp[i] = (e == c) ? 1.0 : 0.0;
        }
        __syncthreads();
        int sweep = 0;
        for (; sweep < max_sweeps; ++sweep) {
            jacobi_active_list(Gs, ldg, n, nrm2, list, sv_cutoff, &s_nact, &s_thr, &s_rot);
            const int nact = s_nact;
            const double thr = s_thr;
            if (nact < 2) break;
            const bool w_smem = nact <= cap;
            if (w_smem) {
                for (int i = tid; i < nact * (ldg / 2); i += THREADS) {
                    const int a = i / (ldg / 2), e = i - a * (ldg / 2);
                    reinterpret_cast<double2*>(Ws)[i] = __ldcg(reinterpret_cast<const double2*>(pp + (size_t)list[a] * ldg) + e);
                }
                __syncthreads();
            }
            jacobi_sweep_dispatch<MAXV2>(Gs, w_smem ? Ws : pp, ldg, ldg, w_smem, list, nact, thr, tol, rotbuf, &s_rot, nrm2, SLOTS);
            if (w_smem) {
                for (int i = tid; i < nact * (ldg / 2); i += THREADS) {
                    const int a = i / (ldg / 2), e = i - a * (ldg / 2);
                    __stcg(reinterpret_cast<double2*>(pp + (size_t)list[a] * ldg) + e, reinterpret_cast<const double2*>(Ws)[i]);
                }
            }
            __syncthreads();
            const int rotated = s_rot;
            __syncthreads();
            if (!rotated) { ++sweep; break; }
        }
        __syncthreads();
        // ---- final norms -> sigma, normalised columns
        jacobi_active_list(Gs, ldg, n, nrm2, list, sv_cutoff, &s_nact, &s_thr, &s_rot);
        const double thr_f = s_thr;
        for (int i = tid; i < n * ldg; i += THREADS) {
            const int c = i / ldg;
            const double s2 = nrm2[c];
            gn[i] = (s2 > thr_f && s2 > 0.0) ? Gs[i] / sqrt(s2) : 0.0;
        }
        for (int i = tid; i < n; i += THREADS) sg[i] = sqrt(nrm2[i]);
        if (tid == 0 && sweeps_out) sweeps_out[b * n_orders_all + oi] = sweep;
    }
}

// Tt[b][order][m'][k] (real) + pass-through rules -> c_out [(L+1)^2][S]
// kind[l]: ORD_PASS copy input, ORD_ZERO masked rows -> 0, ORD_ZEROTH masked rows -> v0[k], ORD_ACTIVE masked rows -> T
__global__ void procrustes_unpack_kernel(const double2* __restrict__ c_in, double2* __restrict__ c_out, const double* __restrict__ tt,
                                         const ProcOrder* __restrict__ orders, const int* __restrict__ kind,
                                         const int* __restrict__ act_index, const uint8_t* __restrict__ radial_mask,
                                         const double* __restrict__ v0, double inv_sqrt_np, int l_max, int n_r, int S,
                                         long long xt_run_stride, int half) {
    const int l = blockIdx.x;
    const int b = blockIdx.y;
    const int kd = kind[l];
    const double is2 = 0.7071067811865476;
    const double* T = nullptr;
    if (kd == ORD_ACTIVE) T = tt + (size_t)b * xt_run_stride + orders[act_index[l]].xt_off;
    const int n_c = 2 * l + 1;
    // half: the coefficients belong to a real field -- only m >= 0 exists in c_in and only m >= 0 is written
    for (int idx = threadIdx.x + (half ? l * n_r : 0); idx < n_c * n_r; idx += blockDim.x) {
        const int mi = idx / n_r, k = idx - mi * n_r;   // mi = m + l
        const int m = mi - l;
        const size_t pos = (size_t)(l * (l + 1) + m) * S + (size_t)b * n_r + k;
        double2 v = ldg2(c_in + pos);
        const bool masked = radial_mask[(size_t)l * n_r + k] != 0;
        if (masked) {
            if (kd == ORD_ZERO) v = make_double2(0, 0);
            else if (kd == ORD_ZEROTH) v = make_double2(v0[k], 0.0);
            else if (kd == ORD_ACTIVE) {
                if (m == 0) v = make_double2(T[k], 0.0);
                else {
                    const int am = m > 0 ? m : -m;
                    const double re = T[(size_t)(2 * am - 1) * n_r + k] * is2, im = T[(size_t)(2 * am) * n_r + k] * is2;
                    if (m > 0) v = make_double2(re, im);
                    else v = (am & 1) ? make_double2(-re, im) : make_double2(re, -im);   // (-1)^m conj
                }
            }
        }
        if (l == 0 && kd != ORD_PASS) { v.x *= inv_sqrt_np; v.y *= inv_sqrt_np; }
        c_out[pos] = v;
    }
}

// ---- fxs_unknowns on request (xfb_get_unknowns) ------------------------------------------------------------------
// polar(G) = sum_c gn[c] (x) pp[c] (see the Jacobi kernel), polar(M) = polar(G)^T (real basis) -> complex columns
// m = -l..l.  One block per row i of the unknown.
__global__ void unknown_assemble_kernel(const double* __restrict__ gn, const double* __restrict__ pp, int ldg, int n_cols, int n_c,
                                        int l, double2* __restrict__ out) {
    __shared__ double row[512];
    const int i = blockIdx.x;
    for (int e = threadIdx.x; e < n_c; e += blockDim.x) {
        double s = 0.0;
        for (int c = 0; c < n_cols; ++c) s += pp[(size_t)c * ldg + i] * gn[(size_t)c * ldg + e];
        row[e] = s;
    }
    __syncthreads();
    const double is2 = 0.7071067811865476;
    for (int m = threadIdx.x; m <= l; m += blockDim.x) {
        if (m == 0) { out[(size_t)i * n_c + l] = make_double2(row[0], 0.0); continue; }
        const double re = row[2 * m - 1] * is2, im = row[2 * m] * is2;
        out[(size_t)i * n_c + l + m] = make_double2(re, im);
        out[(size_t)i * n_c + l - m] = (m & 1) ? make_double2(-re, im) : make_double2(re, -im);
    }
}

// l = 0: PD_0 I_0 is the scalar sum_k V_0(q_k) q_k^2 I_00(q_k); U V^H = its phase (1 for a zero scalar)
__global__ void unknown_zeroth_kernel(const double2* __restrict__ i00, const double* __restrict__ v0, const double* __restrict__ q, int n_r,
                                      double2* __restrict__ out) {
    __shared__ double sx[128], sy[128];
    double ax = 0.0, ay = 0.0;
    for (int k = threadIdx.x; k < n_r; k += blockDim.x) {
        const double w = v0[k] * q[k] * q[k];
        ax += w * i00[k].x; ay += w * i00[k].y;
    }
    sx[threadIdx.x] = ax; sy[threadIdx.x] = ay;
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0.0, y = 0.0;
        for (int t = 0; t < blockDim.x; ++t) { x += sx[t]; y += sy[t]; }
        const double a = hypot(x, y);
        out[0] = (a > 0.0) ? make_double2(x / a, y / a) : make_double2(1.0, 0.0);
    }
}

__global__ void unknown_identity_kernel(double2* __restrict__ out, int n_l, int n_c) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n_l * n_c) out[idx] = make_double2((idx / n_c) == (idx % n_c) ? 1.0 : 0.0, 0.0);
}

// ---- degree-2 invariants  B_l = I_l I_l^H  (fxs_invariant_tools.py:915-923) and their distance to the data ----------------
// The coefficients belong to a REAL field (|rho_hat|^2), so with the real-harmonic columns X (see procrustes_pack_kernel)
// B_l = X_l X_l^T is real symmetric: one real [N_r x (2l+1)] . [(2l+1) x N_r] product per (run, order) on the grouped DMMA GEMM.
// c [(L+1)^2][S] complex (m >= 0 valid) -> x[b][l^2 + m'][k] real, ALL orders l = 0..L
__global__ void pack_real_all_kernel(const double2* __restrict__ c, double* __restrict__ x, int n_r, int S, long long run_stride) {
    const int l = blockIdx.x, b = blockIdx.y;
    const double s2 = 1.4142135623730951;
    double* dst = x + (size_t)b * run_stride + (size_t)l * l * n_r;
    for (int idx = threadIdx.x; idx < (l + 1) * n_r; idx += blockDim.x) {
        const int m = idx / n_r, k = idx - m * n_r;
        const double2 v = ldg2(c + (size_t)(l * (l + 1) + m) * S + (size_t)b * n_r + k);
        if (m == 0) {
            dst[k] = v.x;
        } else {
            dst[(size_t)(2 * m - 1) * n_r + k] = s2 * v.x;
            dst[(size_t)(2 * m) * n_r + k] = s2 * v.y;
        }
    }
}
// deg2_invariant_l2_diff (fxs_IO_methods.py:412-447): err[l] = sum_{q,q'} |Bref_l - mask B_l|^2 / sum |Bref_l|^2, -1 where the
// reference vanishes.  bref is the masked reference with order 0 already divided by the number of particles (:440), norm the
// sums of the masked reference BEFORE that division (:425-426).  One CTA per (order, run); fixed summation order.
__global__ void __launch_bounds__(256) deg2_diff_kernel(const double* __restrict__ bl, const double* __restrict__ bref, const double* __restrict__ norm,
                                                        const uint8_t* __restrict__ radial_mask, int n_r, int n_orders, long long bl_run_stride,
                                                        double* __restrict__ err_out, long long err_run_stride) {
    const int l = blockIdx.x, b = blockIdx.y;
    const double* B = bl + (size_t)b * bl_run_stride + (size_t)l * n_r * n_r;
    const double* R = bref + (size_t)l * n_r * n_r;
    const uint8_t* rm = radial_mask + (size_t)l * n_r;
    double acc = 0.0;
    for (int idx = threadIdx.x; idx < n_r * n_r; idx += blockDim.x) {
        const int q = idx / n_r, q2 = idx - q * n_r;
        const double v = (rm[q] && rm[q2]) ? B[idx] : 0.0;
        const double d = R[idx] - v;
        acc += d * d;
    }
    __shared__ double red[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        const double nl = norm[l];
        err_out[(size_t)b * err_run_stride + l] = (nl != 0.0) ? t / nl : -1.0;
    }
}

