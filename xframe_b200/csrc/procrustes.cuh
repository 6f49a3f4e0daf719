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
    long long g_off;    // per-run offset of G / Gn block (n_cols * n_c)
    long long vw_off;   // per-run offset of VW block (n_cols * N_r)
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

// One-sided Jacobi, persistent CTAs over the (order, run) problems (largest orders first).
//   G : [n_cols][n_c] in global (column i of G = M^T contiguous), copied to shared memory with the column
//       length padded to a multiple of 8 (zeros).
//   vw: [n_cols][N_r] accumulator initialised from vt (V_l^T).  At the start of every sweep the ACTIVE columns are
//       staged into shared memory (they fit once dead columns are dropped: cap columns) and written back at the
//       end of the sweep; only if more than `cap` columns are active are they rotated in global memory (L2).
//   A pair of columns is handled by JG = 8 lanes (4 pairs per warp, 64 pairs per 512-thread CTA = one whole
//   round-robin round in flight); dot products are reduced with 3 xor-shuffles inside the 8-lane group; the G
//   elements stay in registers between the dot and the rotation (one smem read + one write per element).
//   Outputs: gn = G~ / sigma (zero for dropped columns), sigma [n_cols].
#ifndef JG
#define JG 8                 // lanes per column pair (8, 16 or 32)
#endif
#define JGPW (32 / JG)       // pair groups per warp
#define JPASS (64 / (16 * JGPW))   // passes per round: 64 pair slots / (16 warps x groups per warp)
__host__ __device__ inline int jacobi_stride(int len) {
    return len <= 64 ? 64 : 128;    // columns are zero padded to 64 / 128 elements (NV2 = 4 / 8 double2 steps per lane)
}
#define JMAXE (128 / JG)     // max elements per lane (column length <= 128)

// All rounds of one sweep.  WSM: accumulator columns staged in shared memory (the normal case) -- 128-bit shared
// memory accesses, each lane owns the element pairs (2 sub, 2 sub + 1) + 16 t, NV2 = number of such double2 steps of
// a G column (4 for columns up to 64 long, 8 up to 128).  !WSM: rare fallback, accumulator rotated in global memory.
template <bool WSM, int NV2>
__device__ __forceinline__ void jacobi_sweep_rounds(double* __restrict__ Gs, int ldg, int ne, double* __restrict__ Wb, int wstride,
                                                    int n_r_grid, int wr_e, const int* __restrict__ list, int nact, double thr,
                                                    double tol, int slot0, int n_slots, int sub, int* s_rot) {
    const int npad = nact + (nact & 1);
    const int half = npad >> 1;
    const int mod = npad - 1;
    // round-robin (circle method): position 0 is fixed, the element at position k>=1 in round r is
    // 1 + ((k-1-r) mod (npad-1)); tracked incrementally (one decrement with wrap per round, no integer division).
    int pa_[JPASS], pb_[JPASS];
#pragma unroll
    for (int ps = 0; ps < JPASS; ++ps) {
        const int i = slot0 + ps * n_slots;
        const int ka = i, kb = npad - 1 - i;
        pa_[ps] = (ka == 0) ? 0 : 1 + (ka - 1) % mod;
        pb_[ps] = (i < half) ? 1 + (kb - 1) % mod : 1;
    }
    constexpr int WV2 = 128 / (2 * JG);                  // double2 steps of an accumulator column (N_r <= 128, zero padded)
    for (int r = 0; r < npad - 1; ++r) {
#pragma unroll
        for (int ps = 0; ps < JPASS; ++ps) {
            const int i = slot0 + ps * n_slots;
            const int warp_first = (slot0 / JGPW) * JGPW + ps * n_slots;     // first pair slot of this warp in this pass
            if (warp_first >= half) continue;                                // warp-uniform: nothing to do
            const bool has_slot = i < half;
            int pa = pa_[ps], pb = pb_[ps];
            bool valid = has_slot && (pa < nact) && (pb < nact);      // bye against the padding element
            int p = 0, q = 0, wpi = 0, wqi = 0;
            if (valid) {
                int sa = pa, sb = pb;
                p = list[pa]; q = list[pb];
                if (p > q) { int t_ = p; p = q; q = t_; sa = pb; sb = pa; }
                wpi = WSM ? sa : p;                                // smem: compact slot ; global: column index
                wqi = WSM ? sb : q;
            }
            if (i != 0) pa_[ps] = (pa == 1) ? mod : pa - 1;         // positions for the next round
            pb_[ps] = (pb == 1) ? mod : pb - 1;
            if constexpr (WSM) {
                double2* gp = reinterpret_cast<double2*>(Gs + (size_t)p * ldg) + sub;
                double2* gq = reinterpret_cast<double2*>(Gs + (size_t)q * ldg) + sub;
                double2 xg[NV2], yg[NV2];
                double app = 0.0, aqq = 0.0, apq = 0.0;
                if (valid) {
#pragma unroll
                    for (int t = 0; t < NV2; ++t) {
                        xg[t] = gp[JG * t]; yg[t] = gq[JG * t];
                        app += xg[t].x * xg[t].x; aqq += yg[t].x * yg[t].x; apq += xg[t].x * yg[t].x;
                        app += xg[t].y * xg[t].y; aqq += yg[t].y * yg[t].y; apq += xg[t].y * yg[t].y;
                    }
                }
#pragma unroll
                for (int off = JG / 2; off > 0; off >>= 1) {
                    app += __shfl_xor_sync(0xffffffffu, app, off);
                    aqq += __shfl_xor_sync(0xffffffffu, aqq, off);
                    apq += __shfl_xor_sync(0xffffffffu, apq, off);
                }
                bool rot = valid && (app > thr) && (aqq > thr);
                if (rot) rot = apq * apq > (tol * tol) * (app * aqq);
                if (rot) {
                    // tan(2 theta) = 2 apq / (aqq - app), |theta| <= pi/4 (same rotation as the textbook zeta/t form)
                    const double d = aqq - app, s2 = 2.0 * apq;
                    const double rh = rsqrt(d * d + s2 * s2);
                    const double u = 0.5 + 0.5 * fabs(d) * rh;            // cos^2(theta)
                    const double rc = rsqrt(u);
                    const double cs = u * rc;
                    const double sn = copysign(0.5 * s2 * rh, d * s2) * rc;   // sin(2 theta) / (2 cos theta)
#pragma unroll
                    for (int t = 0; t < NV2; ++t) {
                        gp[JG * t] = make_double2(cs * xg[t].x - sn * yg[t].x, cs * xg[t].y - sn * yg[t].y);
                        gq[JG * t] = make_double2(sn * xg[t].x + cs * yg[t].x, sn * xg[t].y + cs * yg[t].y);
                    }
                    double2* wp = reinterpret_cast<double2*>(Wb + (size_t)wpi * wstride) + sub;
                    double2* wq = reinterpret_cast<double2*>(Wb + (size_t)wqi * wstride) + sub;
#pragma unroll
                    for (int t = 0; t < WV2; ++t) {
                        const double2 x = wp[JG * t], y = wq[JG * t];
                        wp[JG * t] = make_double2(cs * x.x - sn * y.x, cs * x.y - sn * y.y);
                        wq[JG * t] = make_double2(sn * x.x + cs * y.x, sn * x.y + cs * y.y);
                    }
                    if (sub == 0) *s_rot = 1;
                }
            } else {
                double* gp = Gs + (size_t)p * ldg + sub;
                double* gq = Gs + (size_t)q * ldg + sub;
                double* wp = Wb + (size_t)wpi * wstride + sub;
                double* wq = Wb + (size_t)wqi * wstride + sub;
                double xw[JMAXE], yw[JMAXE];
                if (valid) {                                        // global accumulator: prefetch before the dot products
#pragma unroll
                    for (int t = 0; t < JMAXE; ++t)
                        if (t < wr_e && sub + JG * t < n_r_grid) { xw[t] = __ldcg(wp + JG * t); yw[t] = __ldcg(wq + JG * t); }
                }
                double app = 0.0, aqq = 0.0, apq = 0.0;
                if (valid) {
#pragma unroll 4
                    for (int t = 0; t < ne; ++t) {
                        const double x = gp[JG * t], y = gq[JG * t];
                        app += x * x; aqq += y * y; apq += x * y;
                    }
                }
#pragma unroll
                for (int off = JG / 2; off > 0; off >>= 1) {
                    app += __shfl_xor_sync(0xffffffffu, app, off);
                    aqq += __shfl_xor_sync(0xffffffffu, aqq, off);
                    apq += __shfl_xor_sync(0xffffffffu, apq, off);
                }
                bool rot = valid && (app > thr) && (aqq > thr);
                if (rot) rot = apq * apq > (tol * tol) * (app * aqq);
                if (rot) {
                    const double d = aqq - app, s2 = 2.0 * apq;
                    const double rh = rsqrt(d * d + s2 * s2);
                    const double u = 0.5 + 0.5 * fabs(d) * rh;
                    const double rc = rsqrt(u);
                    const double cs = u * rc;
                    const double sn = copysign(0.5 * s2 * rh, d * s2) * rc;
#pragma unroll 4
                    for (int tt = 0; tt < ne; ++tt) {
                        const double x = gp[JG * tt], y = gq[JG * tt];
                        gp[JG * tt] = cs * x - sn * y;
                        gq[JG * tt] = sn * x + cs * y;
                    }
#pragma unroll
                    for (int tt = 0; tt < JMAXE; ++tt)
                        if (tt < wr_e && sub + JG * tt < n_r_grid) {
                            __stcg(wp + JG * tt, cs * xw[tt] - sn * yw[tt]);
                            __stcg(wq + JG * tt, sn * xw[tt] + cs * yw[tt]);
                        }
                    if (sub == 0) *s_rot = 1;
                }
            }
        }
        __syncthreads();
    }
}

// rare fallback kept out of line so it does not inflate the register allocation of the shared-memory path
__device__ __noinline__ void jacobi_sweep_global(double* Gs, int ldg, int ne, double* W, int n_r_grid, int wr_e, const int* list, int nact,
                                                 double thr, double tol, int slot0, int n_slots, int sub, int* s_rot) {
    jacobi_sweep_rounds<false, 1>(Gs, ldg, ne, W, n_r_grid, n_r_grid, wr_e, list, nact, thr, tol, slot0, n_slots, sub, s_rot);
}

__global__ void __launch_bounds__(512, 1) procrustes_jacobi_kernel(const double* __restrict__ g_in, double* __restrict__ gn_out,
                                                                   double* __restrict__ vw, const double* __restrict__ vt,
                                                                   double* __restrict__ sigma_out, const ProcOrder* __restrict__ orders,
                                                                   int n_orders, int n_batch, int n_r_grid, long long g_run_stride,
                                                                   long long vw_run_stride, long long sig_run_stride, double sv_cutoff,
                                                                   double tol, int max_sweeps, int* __restrict__ sweeps_out,
                                                                   int smem_doubles) {
    extern __shared__ double smem_j[];
    __shared__ int s_nact, s_rot;
    __shared__ double s_thr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int grp = lane / JG, sub = lane & (JG - 1);     // JGPW groups of JG lanes per warp
    const int n_slots = nwarp * JGPW;
    const int slot0 = warp * JGPW + grp;
    const int wr_e = (n_r_grid + JG - 1) / JG;            // accumulator elements per lane
    const int wld = 128;                                  // smem accumulator column stride (N_r <= 128, zero padded)

    for (int prob = blockIdx.x; prob < n_orders * n_batch; prob += gridDim.x) {
        const int oi = prob / n_batch;                     // orders are sorted largest first
        const int b = prob - oi * n_batch;
        const ProcOrder o = orders[oi];
        const int n = o.n_cols, len = o.n_c;
        const int ne = (len + JG - 1) / JG;                // G elements per lane
        const int ldg = jacobi_stride(len);                // column stride == 8 (mod 16) doubles: the two pair groups of a
                                                           // half warp then fall on disjoint shared-memory banks
        double* Gs = smem_j;                               // [n][ldg]
        double* nrm2 = Gs + (size_t)n * ldg;               // [n]
        int* list = (int*)(nrm2 + n);                      // [n]
        const int ws_off = (n * ldg + n + (n + 1) / 2 + 1) & ~1;      // 16-byte aligned (double2 accesses)
        double* Ws = smem_j + ws_off;                      // [cap][wld]
        const int cap = (smem_doubles - ws_off) / wld;
        const double* g = g_in + (size_t)b * g_run_stride + o.g_off;
        double* W = vw + (size_t)b * vw_run_stride + o.vw_off;
        const double* V0 = vt + o.pd_off;
        __syncthreads();                                   // previous problem fully done with smem
        for (int i = tid; i < n * ldg; i += blockDim.x) {
            const int c = i / ldg, e = i - c * ldg;
            Gs[i] = (e < len) ? g[(size_t)c * len + e] : 0.0;
        }
        for (int i = tid; i < n * n_r_grid; i += blockDim.x) W[i] = V0[i];
        __syncthreads();
        int sweep = 0;
        for (; sweep < max_sweeps; ++sweep) {
            // ---- column norms (one 8-lane group per column)
            for (int c0 = warp * JGPW; c0 < n; c0 += n_slots) {      // warp-uniform trip count (shuffles need all lanes)
                const int c = c0 + grp;
                double s = 0.0;
                if (c < n)
                    for (int t = 0; t < ne; ++t) { const double v = Gs[c * ldg + sub + JG * t]; s += v * v; }
#pragma unroll
                for (int off = JG / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                if (sub == 0 && c < n) nrm2[c] = s;
            }
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
                if (lane == 0) { s_nact = base; s_thr = thr; s_rot = 0; }
            }
            __syncthreads();
            const int nact = s_nact;
            const double thr = s_thr;
            if (nact < 2) break;
            if (nact <= cap) {
                // stage the active accumulator columns in shared memory for this sweep
                for (int i = tid; i < nact * wld; i += blockDim.x) {
                    const int a = i / wld, e = i - a * wld;
                    Ws[i] = (e < n_r_grid) ? __ldcg(W + (size_t)list[a] * n_r_grid + e) : 0.0;
                }
                __syncthreads();
                if (ldg <= 64) jacobi_sweep_rounds<true, 64 / (2 * JG)>(Gs, ldg, ne, Ws, wld, n_r_grid, wr_e, list, nact, thr, tol, slot0, n_slots, sub, &s_rot);
                else jacobi_sweep_rounds<true, 128 / (2 * JG)>(Gs, ldg, ne, Ws, wld, n_r_grid, wr_e, list, nact, thr, tol, slot0, n_slots, sub, &s_rot);
                for (int i = tid; i < nact * n_r_grid; i += blockDim.x) {
                    const int a = i / n_r_grid, e = i - a * n_r_grid;
                    __stcg(W + (size_t)list[a] * n_r_grid + e, Ws[a * wld + e]);
                }
            } else {
                jacobi_sweep_global(Gs, ldg, ne, W, n_r_grid, wr_e, list, nact, thr, tol, slot0, n_slots, sub, &s_rot);
            }
            __syncthreads();
            const int rotated = s_rot;
            __syncthreads();
            if (!rotated) { ++sweep; break; }
        }
        __syncthreads();
        // ---- final norms -> sigma, normalised columns
        for (int c0 = warp * JGPW; c0 < n; c0 += n_slots) {
            const int c = c0 + grp;
            double s = 0.0;
            if (c < n)
                for (int t = 0; t < ne; ++t) { const double v = Gs[c * ldg + sub + JG * t]; s += v * v; }
#pragma unroll
            for (int off = JG / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (sub == 0 && c < n) nrm2[c] = s;
        }
        __syncthreads();
        if (warp == 0) {
            double mx = 0.0;
            for (int c = lane; c < n; c += 32) mx = fmax(mx, nrm2[c]);
            mx = warp_max(mx);
            if (lane == 0) s_thr = sv_cutoff * sv_cutoff * mx;
        }
        __syncthreads();
        const double thr_f = s_thr;
        double* gn = gn_out + (size_t)b * g_run_stride + o.g_off;
        for (int i = tid; i < n * len; i += blockDim.x) {
            const int c = i / len, e = i - c * len;
            const double s2 = nrm2[c];
            gn[i] = (s2 > thr_f && s2 > 0.0) ? Gs[c * ldg + e] / sqrt(s2) : 0.0;
        }
        double* sg = sigma_out + (size_t)b * sig_run_stride + (size_t)oi * n_r_grid;
        for (int i = tid; i < n; i += blockDim.x) sg[i] = sqrt(nrm2[i]);
        if (tid == 0 && sweeps_out) sweeps_out[b * n_orders + oi] = sweep;
    }
}

// Tt[b][order][m'][k] (real) + pass-through rules -> c_out [(L+1)^2][S]
// kind[l]: ORD_PASS copy input, ORD_ZERO masked rows -> 0, ORD_ZEROTH masked rows -> v0[k], ORD_ACTIVE masked rows -> T
__global__ void procrustes_unpack_kernel(const double2* __restrict__ c_in, double2* __restrict__ c_out, const double* __restrict__ tt,
                                         const ProcOrder* __restrict__ orders, const int* __restrict__ kind,
                                         const int* __restrict__ act_index, const uint8_t* __restrict__ radial_mask,
                                         const double* __restrict__ v0, double inv_sqrt_np, int l_max, int n_r, int S,
                                         long long xt_run_stride) {
    const int l = blockIdx.x;
    const int b = blockIdx.y;
    const int kd = kind[l];
    const double is2 = 0.7071067811865476;
    const double* T = nullptr;
    if (kd == ORD_ACTIVE) T = tt + (size_t)b * xt_run_stride + orders[act_index[l]].xt_off;
    const int n_c = 2 * l + 1;
    for (int idx = threadIdx.x; idx < n_c * n_r; idx += blockDim.x) {
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
// gn [n_cols][n_c] = U~^T (zero rows for dropped directions), vw [n_cols][n_r] = J^T (accumulator started from the
// identity).  polar(M) = J U~^T (real basis) -> complex columns m = -l..l.  One block per row i of the unknown.
__global__ void unknown_assemble_kernel(const double* __restrict__ gn, const double* __restrict__ vw, int n_r, int n_cols, int n_c,
                                        int l, double2* __restrict__ out) {
    __shared__ double row[512];
    const int i = blockIdx.x;
    for (int e = threadIdx.x; e < n_c; e += blockDim.x) {
        double s = 0.0;
        for (int c = 0; c < n_cols; ++c) s += vw[(size_t)c * n_r + i] * gn[(size_t)c * n_c + e];
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
