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

// One CTA per (order, run).  G: [n_cols][n_c] (column i of G^T contiguous) in global, copied to smem.
// vw: [n_cols][N_r] initialised from vt (V_l^T), rotated in global memory (L2 resident).
// Outputs: gn = G~ / sigma (zero for dropped columns), sigma [n_cols].
__global__ void __launch_bounds__(512) procrustes_jacobi_kernel(const double* __restrict__ g_in, double* __restrict__ gn_out,
                                                                double* __restrict__ vw, const double* __restrict__ vt,
                                                                double* __restrict__ sigma_out, const ProcOrder* __restrict__ orders,
                                                                int n_orders, int n_r_grid, long long g_run_stride,
                                                                long long vw_run_stride, long long sig_run_stride, double sv_cutoff,
                                                                double tol, int max_sweeps, int* __restrict__ sweeps_out) {
    extern __shared__ double smem_j[];
    const int oi = blockIdx.x % n_orders;      // orders are sorted largest first
    const int b = blockIdx.x / n_orders;
    const ProcOrder o = orders[oi];
    const int n = o.n_cols, len = o.n_c;
    double* Gs = smem_j;                       // [n][len]
    double* nrm2 = Gs + (size_t)n * len;       // [n]
    int* list = (int*)(nrm2 + n);              // [n]
    __shared__ int s_nact, s_rot;
    __shared__ double s_thr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const double* g = g_in + (size_t)b * g_run_stride + o.g_off;
    double* W = vw + (size_t)b * vw_run_stride + o.vw_off;
    const double* V0 = vt + o.pd_off;
    for (int i = tid; i < n * len; i += blockDim.x) Gs[i] = g[i];
    for (int i = tid; i < n * n_r_grid; i += blockDim.x) W[i] = V0[i];
    __syncthreads();
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        // column norms
        for (int cidx = warp; cidx < n; cidx += nwarp) {
            double s = 0.0;
            for (int e = lane; e < len; e += 32) { const double v = Gs[cidx * len + e]; s += v * v; }
            s = warp_sum(s);
            if (lane == 0) nrm2[cidx] = s;
        }
        __syncthreads();
        if (warp == 0) {
            double mx = 0.0;
            for (int cidx = lane; cidx < n; cidx += 32) mx = fmax(mx, nrm2[cidx]);
            mx = warp_max(mx);
            const double thr = sv_cutoff * sv_cutoff * mx;
            // ordered compaction of the active columns
            int base = 0;
            for (int c0 = 0; c0 < n; c0 += 32) {
                const int cidx = c0 + lane;
                const bool act = (cidx < n) && (nrm2[cidx] > thr);
                const unsigned bal = __ballot_sync(0xffffffffu, act);
                if (act) list[base + __popc(bal & ((1u << lane) - 1u))] = cidx;
                base += __popc(bal);
            }
            if (lane == 0) { s_nact = base; s_thr = thr; s_rot = 0; }
        }
        __syncthreads();
        const int nact = s_nact;
        const double thr = s_thr;
        if (nact < 2) break;
        const int npad = nact + (nact & 1);
        const int half = npad >> 1;
        for (int r = 0; r < npad - 1; ++r) {
            for (int i = warp; i < half; i += nwarp) {
                const int ka = i, kb = npad - 1 - i;
                const int pa = (ka == 0) ? 0 : 1 + ((ka - 1 - r) % (npad - 1) + (npad - 1)) % (npad - 1);
                const int pb = 1 + ((kb - 1 - r) % (npad - 1) + (npad - 1)) % (npad - 1);
                if (pa >= nact || pb >= nact) continue;  // bye
                int p = list[pa], q = list[pb];
                if (p > q) { const int t_ = p; p = q; q = t_; }
                double* gp = Gs + (size_t)p * len;
                double* gq = Gs + (size_t)q * len;
                double app = 0.0, aqq = 0.0, apq = 0.0;
                for (int e = lane; e < len; e += 32) {
                    const double x = gp[e], y = gq[e];
                    app += x * x; aqq += y * y; apq += x * y;
                }
                app = warp_sum(app); aqq = warp_sum(aqq); apq = warp_sum(apq);
                if (app <= thr || aqq <= thr) continue;
                if (fabs(apq) <= tol * sqrt(app * aqq)) continue;
                const double zeta = (aqq - app) / (2.0 * apq);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                for (int e = lane; e < len; e += 32) {
                    const double x = gp[e], y = gq[e];
                    gp[e] = cs * x - sn * y;
                    gq[e] = sn * x + cs * y;
                }
                double* wp = W + (size_t)p * n_r_grid;
                double* wq = W + (size_t)q * n_r_grid;
                for (int e = lane; e < n_r_grid; e += 32) {
                    const double x = wp[e], y = wq[e];
                    wp[e] = cs * x - sn * y;
                    wq[e] = sn * x + cs * y;
                }
                if (lane == 0) s_rot = 1;
            }
            __syncthreads();
        }
        const int rot = s_rot;
        __syncthreads();
        if (!rot) { ++sweep; break; }
    }
    __syncthreads();
    // final norms -> sigma, normalised columns
    for (int cidx = warp; cidx < n; cidx += nwarp) {
        double s = 0.0;
        for (int e = lane; e < len; e += 32) { const double v = Gs[cidx * len + e]; s += v * v; }
        s = warp_sum(s);
        if (lane == 0) nrm2[cidx] = s;
    }
    __syncthreads();
    if (warp == 0) {
        double mx = 0.0;
        for (int cidx = lane; cidx < n; cidx += 32) mx = fmax(mx, nrm2[cidx]);
        mx = warp_max(mx);
        if (lane == 0) s_thr = sv_cutoff * sv_cutoff * mx;
    }
    __syncthreads();
    const double thr = s_thr;
    double* gn = gn_out + (size_t)b * g_run_stride + o.g_off;
    for (int i = tid; i < n * len; i += blockDim.x) {
        const int cidx = i / len;
        const double s2 = nrm2[cidx];
        gn[i] = (s2 > thr && s2 > 0.0) ? Gs[i] / sqrt(s2) : 0.0;
    }
    double* sg = sigma_out + (size_t)b * sig_run_stride + (size_t)oi * n_r_grid;
    for (int i = tid; i < n; i += blockDim.x) sg[i] = sqrt(nrm2[i]);
    if (tid == 0 && sweeps_out) sweeps_out[b * n_orders + oi] = sweep;
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
