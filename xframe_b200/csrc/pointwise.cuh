// Bandwidth-bound elementwise / reduction kernels of the MTIP loop.  Every kernel
// touches each grid point once with 128-bit accesses; reductions are two-stage
// and deterministic (per-block partials summed in a fixed order).
#pragma once
#include "common.cuh"

// A batch of per-run grids that may live in one of several slots of a pool
// (history / best-density bookkeeping without copies, reconstruct.py:924-938).
struct SlotView {
    double2* base;
    const int* slot;        // per-run slot index or nullptr
    long long slot_stride;  // elements between slots
    long long run_stride;   // elements between runs
};
__device__ __forceinline__ double2* slot_run_ptr(const SlotView& v, int b) {
    return v.base + (v.slot ? (long long)v.slot[b] * v.slot_stride : 0ll) + (long long)b * v.run_stride;
}

// misk.py:159-168  square_grid: data * conj(data)
__global__ void square_kernel(const double2* __restrict__ in, double2* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double2 v = ldg2(in + i);
        out[i] = make_double2(__dadd_rn(__dmul_rn(v.x, v.x), __dmul_rn(v.y, v.y)), 0.0);
    }
}

// misk.py:221-225 abs_value
__global__ void abs_kernel(SlotView in, double2* __restrict__ out, long long per_run) {
    const int b = blockIdx.y;
    const double2* src = slot_run_ptr(in, b);
    double2* dst = out + (long long)b * per_run;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_run; i += (long long)gridDim.x * blockDim.x) {
        const double2 v = src[i];
        dst[i] = make_double2(sqrt(__dadd_rn(__dmul_rn(v.x, v.x), __dmul_rn(v.y, v.y))), 0.0);
    }
}

// fxs_Projections.py:899-909 project_to_modified_intensity
__global__ void modify_intensity_kernel(const double2* __restrict__ rho_hat, const double2* __restrict__ i_proj, SlotView out,
                                        long long per_run) {
    const int b = blockIdx.y;
    const double2* rh = rho_hat + (long long)b * per_run;
    const double2* ip = i_proj + (long long)b * per_run;
    double2* dst = slot_run_ptr(out, b);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_run; i += (long long)gridDim.x * blockDim.x) {
        const double2 v = ldg2(rh + i);
        const double sq = __dadd_rn(__dmul_rn(v.x, v.x), __dmul_rn(v.y, v.y));
        const double ni = ldg2(ip + i).x;
        const double mult = mod_intensity_multiplier(ni, sq);
        dst[i] = make_double2(v.x * mult, v.y * mult);
    }
}

// mathLibrary.py:616-624 (sic: q**4) and fxs_Projections.py:294-298 multiply_with_ft_gaussian
__global__ void mul_gauss_kernel(double2* __restrict__ data, const double* __restrict__ q, double sigma, int n_r, long long shell) {
    const int r = blockIdx.x % n_r;
    const long long base = (long long)blockIdx.x * shell;
    const double a = 1.0 / (2.0 * sigma * sigma);
    const double pi = 3.141592653589793;
    const double qq = q[r] * q[r];
    const double g = sqrt(pi / a) * exp(-(pi * pi) * (qq * qq) / a);
    for (long long i = blockIdx.y * (long long)blockDim.x + threadIdx.x; i < shell; i += (long long)gridDim.y * blockDim.x) {
        double2 v = data[base + i];
        v.x *= g; v.y *= g;
        data[base + i] = v;
    }
}

// Scalars of one iteration that change from step to step (HIO beta, error-history column, sub-loop iteration).  When an iteration is
// replayed from a CUDA graph they are read from device memory (written by iter_params_kernel right before the graph launch) instead of
// being kernel arguments frozen at capture time.
struct IterParams { double beta; int it; int outer; int pad; };
__global__ void iter_params_kernel(IterParams* ip, double beta, int it, int outer) { ip->beta = beta; ip->it = it; ip->outer = outer; ip->pad = 0; }

struct RealDesc {
    int n_ops;
    int ops[4];          // 1 support, 2 value_threshold, 3 limit_imag, 4 average_center
    int considered[4];
    int use_lo, use_hi;
    double lo, hi, imag_limit;
    int err_inside;
    int avg_shells;      // average_center: the first avg_shells radial shells are replaced by their angular mean (fxs_Projections.py:96-110)
};

// one step of the real projection chain on the value p of one grid point (fxs_Projections.py:72-130); returns "changed"
__device__ __forceinline__ bool real_chain_op(const RealDesc& rd, int k, bool outside, double2& p, const double2* avg_mean, long long i,
                                              long long shell) {
    const int op = rd.ops[k];
    if (op == 1) {
        if (outside) { p = make_double2(0.0, 0.0); return true; }
    } else if (op == 2) {
        const bool lo = rd.use_lo && (p.x < rd.lo);
        const bool hi = rd.use_hi && (p.x > rd.hi);
        if (lo) p.x = rd.lo;
        if (hi) p.x = rd.hi;
        return lo || hi;
    } else if (op == 3) {
        if (fabs(p.y) >= rd.imag_limit) { p.y = 0.0; return true; }
    } else if (op == 4) {
        if (avg_mean && i < (long long)rd.avg_shells * shell) p = avg_mean[i / shell];     // mask of this projection is False (:102,106)
    }
    return false;
}
// rho_new of one grid point (see real_update_kernel)
__device__ __forceinline__ double2 real_combine(double2 v, double2 prev, double2 t_rt, double2 t_rt0, bool has_rt, bool has_rt0, long long i,
                                                long long shell) {
    if (has_rt && i >= shell) { v.x += prev.x - t_rt.x; v.y += prev.y - t_rt.y; }
    if (has_rt0) {      // fused ft_stab: rho_ift holds IFT(rho_hat' - rho_hat); add rho (r>=1) or shell 0 of IFT(rho_hat)
        const double2 t = (i >= shell) ? prev : t_rt0;
        v.x += t.x; v.y += t.y;
    }
    return v;
}

#define RU_THREADS 256
// real_projection + HIO/ER + l2_projection_diff partial sums.
//   rho_new = rho_ift (+ (rho_prev - rho_rt) for radial index >= 1 when rho_rt != nullptr)   reconstruct.py:584-593, misk.py:325-329
//             or, fused form (rt0 != nullptr): rho_ift = IFT(rho_hat' - rho_hat) and rho_new = rho_ift + rho_prev (r >= 1),
//             rho_ift + IFT(rho_hat)[shell 0] (r = 0) -- the same quantity by linearity of IFT
//   projection chain                                                                         fxs_Projections.py:72-130
//   HIO: where(mask, rho_prev - beta (rho_new - proj), proj) ; ER: proj                      fxs_IO_methods.py:56-68
//   partial[b][block][0..1] = sum w |rho_new-proj|^2 , sum w |rho_new|^2 over the error region
#ifndef RU_MINB
#define RU_MINB 2
#endif
__global__ void __launch_bounds__(RU_THREADS, RU_MINB) real_update_kernel(const double2* __restrict__ rho_ift, const double2* __restrict__ rho_rt,
                                                                 SlotView rho_prev, SlotView rho_next, const uint8_t* __restrict__ support,
                                                                 const int* __restrict__ support_slot, long long support_slot_stride,
                                                                 const int* __restrict__ enforce, const uint8_t* __restrict__ init_support,
                                                                 const double* __restrict__ wt, RealDesc rd, int method, double beta,
                                                                 int n_theta, int n_phi, int wt_div, long long per_run, double* __restrict__ partial,
                                                                 const double2* __restrict__ rt0, const double2* __restrict__ avg_mean,
                                                                 const IterParams* __restrict__ ip) {
    if (ip) beta = ip->beta;
    const int b = blockIdx.y;
    const double2* ri = rho_ift + (long long)b * per_run;
    const double2* rt = rho_rt ? rho_rt + (long long)b * per_run : nullptr;
    const double2* rp = slot_run_ptr(rho_prev, b);
    double2* rn = slot_run_ptr(rho_next, b);
    const uint8_t* sup = support + (support_slot ? (long long)support_slot[b] * support_slot_stride : 0ll) + (long long)b * per_run;
    const bool enf = enforce ? (enforce[b] != 0) : true;
    const long long shell = (long long)n_theta * n_phi;
    double s_diff = 0.0, s_val = 0.0;
    // one grid point: combine, project, HIO/ER, error integrands
    const double2* avg_b = avg_mean ? avg_mean + (long long)b * rd.avg_shells : nullptr;
    auto point = [&](long long i, double2 v, double2 prev, double2 t_rt, double2 t_rt0, uint8_t sup_i, uint8_t init_i) -> double2 {
        v = real_combine(v, prev, t_rt, t_rt0, rt != nullptr, rt0 != nullptr, i, shell);
        const bool in_init = init_i != 0;
        const bool outside = enf ? (!in_init || sup_i == 0) : (sup_i == 0);
        double2 p = v;
        bool msel = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k >= rd.n_ops) break;
            const bool changed = real_chain_op(rd, k, outside, p, avg_b, i, shell);
            if (rd.considered[k]) msel = msel || changed;
        }
        double2 o = p;
        if (method == 0 && msel) o = make_double2(prev.x - beta * (v.x - p.x), prev.y - beta * (v.y - p.y));
        if (!rd.err_inside || in_init) {
            const double w = __ldg(wt + (unsigned)i / (unsigned)wt_div);    // per_run < 2^31: 32-bit division; wt_div = n_phi (3-D) or 1 (2-D)
            const double dx = v.x - p.x, dy = v.y - p.y;
            s_diff += w * (dx * dx + dy * dy);
            s_val += w * (v.x * v.x + v.y * v.y);
        }
        return o;
    };
    const double2 zero2 = make_double2(0.0, 0.0);
    if ((per_run & 3) == 0 && (shell & 3) == 0) {
        // 4 consecutive points per thread and iteration: 4 x 128-bit density loads per array, masks as 32-bit words
        const long long n4 = per_run >> 2;
        for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n4; q += (long long)gridDim.x * blockDim.x) {
            const long long i0 = q << 2;
            double2 v[4], pr[4], trt[4], trt0[4];
            // 256-bit accesses: a thread's 64 contiguous bytes of an array are two full 32-byte sectors per instruction
            ld_global_256_nc(ri + i0, v[0], v[1]); ld_global_256_nc(ri + i0 + 2, v[2], v[3]);
            ld_global_256(rp + i0, pr[0], pr[1]); ld_global_256(rp + i0 + 2, pr[2], pr[3]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                trt[u] = (rt && i0 >= shell) ? ldg2(rt + i0 + u) : zero2;
                trt0[u] = (rt0 && i0 < shell) ? ldg2(rt0 + (long long)b * shell + i0 + u) : zero2;
            }
            const uchar4 s4 = *reinterpret_cast<const uchar4*>(sup + i0), n4m = __ldg(reinterpret_cast<const uchar4*>(init_support + i0));
            const uint8_t sb[4] = {s4.x, s4.y, s4.z, s4.w}, ib[4] = {n4m.x, n4m.y, n4m.z, n4m.w};
            double2 o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) o[u] = point(i0 + u, v[u], pr[u], trt[u], trt0[u], sb[u], ib[u]);
            st_global_256(rn + i0, o[0], o[1]); st_global_256(rn + i0 + 2, o[2], o[3]);
        }
    } else {
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_run; i += (long long)gridDim.x * blockDim.x) {
            const double2 t_rt = (rt && i >= shell) ? ldg2(rt + i) : zero2;
            const double2 t_rt0 = (rt0 && i < shell) ? ldg2(rt0 + (long long)b * shell + i) : zero2;
            rn[i] = point(i, ldg2(ri + i), rp[i], t_rt, t_rt0, sup[i], init_support[i]);
        }
    }
    __shared__ double red[2][RU_THREADS / 32];
    s_diff = warp_sum(s_diff);
    s_val = warp_sum(s_val);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s_diff; red[1][threadIdx.x >> 5] = s_val; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int w = 0; w < RU_THREADS / 32; ++w) { a += red[0][w]; c += red[1][w]; }
        partial[((long long)b * gridDim.x + blockIdx.x) * 2 + 0] = a;
        partial[((long long)b * gridDim.x + blockIdx.x) * 2 + 1] = c;
    }
}

// average_center (fxs_Projections.py:96-110): angular mean of the chain value at the position of the 'average_center'
// entry (the operations before it applied) over each of the first rd.avg_shells shells; one CTA per (shell, run), fixed
// summation order.  The real update then substitutes the mean at those points.
__global__ void __launch_bounds__(256) average_center_kernel(const double2* __restrict__ rho_ift, const double2* __restrict__ rho_rt, SlotView rho_prev,
                                                             const uint8_t* __restrict__ support, const int* __restrict__ support_slot,
                                                             long long support_slot_stride, const int* __restrict__ enforce,
                                                             const uint8_t* __restrict__ init_support, RealDesc rd, int n_theta, int n_phi,
                                                             long long per_run, const double2* __restrict__ rt0, double2* __restrict__ avg_mean) {
    const int r = blockIdx.x, b = blockIdx.y;
    const long long shell = (long long)n_theta * n_phi;
    const double2* ri = rho_ift + (long long)b * per_run;
    const double2* rt = rho_rt ? rho_rt + (long long)b * per_run : nullptr;
    const double2* rp = slot_run_ptr(rho_prev, b);
    const uint8_t* sup = support + (support_slot ? (long long)support_slot[b] * support_slot_stride : 0ll) + (long long)b * per_run;
    const bool enf = enforce ? (enforce[b] != 0) : true;
    int k_avg = 0;
    while (k_avg < rd.n_ops && rd.ops[k_avg] != 4) ++k_avg;
    double sx = 0.0, sy = 0.0;
    const double2 zero2 = make_double2(0.0, 0.0);
    for (long long j = threadIdx.x; j < shell; j += blockDim.x) {
        const long long i = (long long)r * shell + j;
        const double2 t_rt = (rt && i >= shell) ? rt[i] : zero2;
        const double2 t_rt0 = (rt0 && i < shell) ? rt0[(long long)b * shell + i] : zero2;
        double2 p = real_combine(ri[i], rp[i], t_rt, t_rt0, rt != nullptr, rt0 != nullptr, i, shell);
        const bool outside = enf ? (init_support[i] == 0 || sup[i] == 0) : (sup[i] == 0);
        for (int k = 0; k < k_avg; ++k) real_chain_op(rd, k, outside, p, nullptr, i, shell);
        sx += p.x; sy += p.y;
    }
    __shared__ double red[2][8];
    sx = warp_sum(sx); sy = warp_sum(sy);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sx; red[1][threadIdx.x >> 5] = sy; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int w = 0; w < 8; ++w) { a += red[0][w]; c += red[1][w]; }
        avg_mean[(long long)b * rd.avg_shells + r] = make_double2(a / (double)shell, c / (double)shell);
    }
}

// second stage: fixed-order sum of partials -> err[b][2]
__global__ void reduce_pairs_kernel(const double* __restrict__ partial, int n_blocks, double* __restrict__ err) {
    const int b = blockIdx.x;
    double a = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < n_blocks; i += 32) {
        a += partial[((long long)b * n_blocks + i) * 2 + 0];
        c += partial[((long long)b * n_blocks + i) * 2 + 1];
    }
    a = warp_sum(a); c = warp_sum(c);
    if (threadIdx.x == 0) { err[b * 2 + 0] = a; err[b * 2 + 1] = c; }
}

// shrink wrap: min / max of max(Re c, 0) per run (fxs_Projections.py:245-258)
__global__ void __launch_bounds__(256) sw_minmax_kernel(const double2* __restrict__ conv, long long per_run, double* __restrict__ partial) {
    const int b = blockIdx.y;
    const double2* c = conv + (long long)b * per_run;
    double mn = INFINITY, mx = -INFINITY;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_run; i += (long long)gridDim.x * blockDim.x) {
        double v = ldg2(c + i).x;
        if (v < 0.0) v = 0.0;
        mn = fmin(mn, v); mx = fmax(mx, v);
    }
    __shared__ double red[2][8];
    mn = warp_min(mn); mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = mn; red[1][threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { mn = fmin(mn, red[0][w]); mx = fmax(mx, red[1][w]); }
        partial[((long long)b * gridDim.x + blockIdx.x) * 2 + 0] = mn;
        partial[((long long)b * gridDim.x + blockIdx.x) * 2 + 1] = mx;
    }
}
__global__ void sw_minmax_final_kernel(const double* __restrict__ partial, int n_blocks, double* __restrict__ mm) {
    const int b = blockIdx.x;
    double mn = INFINITY, mx = -INFINITY;
    for (int i = threadIdx.x; i < n_blocks; i += 32) {
        mn = fmin(mn, partial[((long long)b * n_blocks + i) * 2 + 0]);
        mx = fmax(mx, partial[((long long)b * n_blocks + i) * 2 + 1]);
    }
    mn = warp_min(mn); mx = warp_max(mx);
    if (threadIdx.x == 0) { mm[b * 2] = mn; mm[b * 2 + 1] = mx; }
}
__global__ void sw_mask_kernel(const double2* __restrict__ conv, const double* __restrict__ mm, double threshold,
                               uint8_t* __restrict__ out, const int* __restrict__ out_slot, long long out_slot_stride, long long per_run) {
    const int b = blockIdx.y;
    const double2* c = conv + (long long)b * per_run;
    uint8_t* dst = out + (out_slot ? (long long)out_slot[b] * out_slot_stride : 0ll) + (long long)b * per_run;
    const double mn = mm[b * 2], mx = mm[b * 2 + 1];
    const double lim = mn + threshold * (mx - mn);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_run; i += (long long)gridDim.x * blockDim.x) {
        double v = ldg2(c + i).x;
        if (v < 0.0) v = 0.0;
        dst[i] = (v >= lim) ? 1 : 0;
    }
}

// ---- coefficient layout conversion: direct [S][NLM] <-> internal [NLM][S] (complex), 32x32 tiles
__global__ void transpose_c128_kernel(const double2* __restrict__ in, double2* __restrict__ out, int rows, int cols) {
    __shared__ double2 tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = in[(size_t)r * cols + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[threadIdx.x][j];
    }
}

// 2-D helper: Hermitian completion of the 'real' coefficient rows before the complex inverse DFT (irfft semantics,
// numpy.fft.irfft with odd N): rows [S][N], c[N-j] = conj(c[j]) for j = 1..N/2, Im c[0] = 0
__global__ void hermitian_complete_rows_kernel(double2* __restrict__ c, int S, int N) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int half = N / 2 + 1;
    if (idx >= (long long)S * half) return;
    const int s = (int)(idx / half), j = (int)(idx - (long long)s * half);
    double2* row = c + (size_t)s * N;
    if (j == 0) row[0].y = 0.0;
    else { const double2 v = row[j]; row[N - j] = make_double2(v.x, -v.y); }
}

// ---- loop bookkeeping (reconstruct.py:927-939): error history, best-so-far, slot rotation
struct LoopState {
    int* rho_cur; int* rho_best; int* rho_next;       // slots in the rho pool
    int* rh_cur;  int* rh_best;  int* rh_next;        // slots in the rho_hat' pool
    int* mask_cur; int* mask_best; int* mask_next;    // slots in the support pool
    int* enforce_cur; int* enforce_best;              // enforce_initial_support at the time
    int* enforce_rep;                                 // enforce flag of the REPORTED current mask (state['mask'], reconstruct.py:883,948)
    int* best_iter;                                   // sub-loop iteration of the best error (state['best_iteration'], :938)
    int* nonfinite;                                   // iterations whose error was NaN / inf, per run (diagnostics)
    double* best_err; double* last_err; double* hist; int hist_cap;
};
__device__ __forceinline__ int free_slot(int a, int b) {   // smallest slot in {0,1,2} different from a and b
    for (int s = 0; s < 3; ++s) if (s != a && s != b) return s;
    return 0;
}
__global__ void loop_update_kernel(LoopState st, const double* __restrict__ err, int it, int n_batch, int outer_it, const IterParams* __restrict__ ip) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_batch) return;
    if (ip) { it = ip->it; outer_it = ip->outer; }
    const double num = err[b * 2], den = err[b * 2 + 1];
    const double e = (den != 0.0) ? num / den : INFINITY;
    if (it < st.hist_cap) st.hist[(long long)b * st.hist_cap + it] = e;
    st.last_err[b] = e;
    if (!isfinite(e)) st.nonfinite[b]++;
    st.rho_cur[b] = st.rho_next[b];
    st.rh_cur[b] = st.rh_next[b];
    if (st.best_err[b] > e) {
        st.best_err[b] = e;
        st.rho_best[b] = st.rho_cur[b];
        st.rh_best[b] = st.rh_cur[b];
        st.mask_best[b] = st.mask_cur[b];
        st.enforce_best[b] = st.enforce_rep[b];
        st.best_iter[b] = outer_it;
    }
    st.rho_next[b] = free_slot(st.rho_cur[b], st.rho_best[b]);
    st.rh_next[b] = free_slot(st.rh_cur[b], st.rh_best[b]);
}
// SW bookkeeping (reconstruct.py:877-885): enforce decision from the last main error, new mask becomes current
__global__ void sw_update_kernel(LoopState st, double error_limit, int have_error, int n_batch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_batch) return;
    st.enforce_cur[b] = (have_error && st.last_err[b] > error_limit) ? 1 : 0;
    st.enforce_rep[b] = st.enforce_cur[b];
    st.mask_cur[b] = st.mask_next[b];
    st.mask_next[b] = free_slot(st.mask_cur[b], st.mask_best[b]);
}
// End of a sub-loop with a finite best_density_not_in_first_n_iterations (reconstruct.py:945-949): runs whose best error came
// after iteration n_first continue from the best pair and the best mask.  real_pr.support = best_mask re-applies the CURRENT
// enforce flag on top of the (already effective) best mask, while the reported state['mask'] is the best mask itself.
__global__ void select_best_kernel(LoopState st, int n_first, int n_batch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_batch || st.best_iter[b] <= n_first) return;
    st.rho_cur[b] = st.rho_best[b];
    st.rh_cur[b] = st.rh_best[b];
    st.mask_cur[b] = st.mask_best[b];
    st.enforce_rep[b] = st.enforce_best[b];
    st.enforce_cur[b] = (st.enforce_cur[b] || st.enforce_best[b]) ? 1 : 0;
    st.rho_next[b] = free_slot(st.rho_cur[b], st.rho_best[b]);
    st.rh_next[b] = free_slot(st.rh_cur[b], st.rh_best[b]);
    st.mask_next[b] = free_slot(st.mask_cur[b], st.mask_best[b]);
}
// SW_center bookkeeping (reconstruct.py:886-897 with the reference's exchanged pair, see oracle/mtip.py): the freshly written
// next slots of both pools become current
__global__ void rotate_pair_kernel(LoopState st, int n_batch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_batch) return;
    st.rho_cur[b] = st.rho_next[b];
    st.rh_cur[b] = st.rh_next[b];
    st.rho_next[b] = free_slot(st.rho_cur[b], st.rho_best[b]);
    st.rh_next[b] = free_slot(st.rh_cur[b], st.rh_best[b]);
}
// |x| of a slot view -> real array [nb][per_run]  (np.abs(hist[-1][0]).real, reconstruct.py:901)
__global__ void abs_real_kernel(SlotView in, double* __restrict__ out, long long per_run) {
    const int b = blockIdx.y;
    const double2* src = slot_run_ptr(in, b);
    double* dst = out + (long long)b * per_run;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_run; i += (long long)gridDim.x * blockDim.x) {
        const double2 v = src[i];
        dst[i] = sqrt(__dadd_rn(__dmul_rn(v.x, v.x), __dmul_rn(v.y, v.y)));
    }
}
// project_to_fixed_intensity (fxs_Projections.py:911-922): rho_hat sqrt(fixed / |rho_hat|^2) where both are >= 0, else 0
__global__ void fixed_intensity_kernel(const double2* __restrict__ rho_hat, const double* __restrict__ fixed, SlotView out, long long per_run) {
    const int b = blockIdx.y;
    const double2* rh = rho_hat + (long long)b * per_run;
    const double* fx = fixed + (long long)b * per_run;
    double2* dst = slot_run_ptr(out, b);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_run; i += (long long)gridDim.x * blockDim.x) {
        const double2 v = ldg2(rh + i);
        const double sq = __dadd_rn(__dmul_rn(v.x, v.x), __dmul_rn(v.y, v.y));
        const double mult = mod_intensity_multiplier(fx[i], sq);
        dst[i] = make_double2(v.x * mult, v.y * mult);
    }
}
// effective support mask = ~_mask[0] (fxs_Projections.py:50-58)
__global__ void effective_support_kernel(const uint8_t* __restrict__ pool, const int* __restrict__ slot, const int* __restrict__ enforce,
                                         const uint8_t* __restrict__ init_support, long long slot_stride, long long per_run,
                                         uint8_t* __restrict__ out) {
    const int b = blockIdx.y;
    const uint8_t* src = pool + (long long)slot[b] * slot_stride + (long long)b * per_run;
    const bool enf = enforce[b] != 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_run; i += (long long)gridDim.x * blockDim.x)
        out[(long long)b * per_run + i] = (src[i] != 0 && (!enf || init_support[i] != 0)) ? 1 : 0;
}
__global__ void gather_slot_kernel(SlotView in, double2* __restrict__ out, long long per_run) {
    const int b = blockIdx.y;
    const double2* src = slot_run_ptr(in, b);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_run; i += (long long)gridDim.x * blockDim.x)
        out[(long long)b * per_run + i] = src[i];
}
__global__ void scatter_slot_kernel(const double2* __restrict__ in, SlotView out, long long per_run) {
    const int b = blockIdx.y;
    double2* dst = slot_run_ptr(out, b);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_run; i += (long long)gridDim.x * blockDim.x)
        dst[i] = in[(long long)b * per_run + i];
}
__global__ void copy_slot_kernel(SlotView in, SlotView out, long long per_run) {
    const int b = blockIdx.y;
    const double2* src = slot_run_ptr(in, b);
    double2* dst = slot_run_ptr(out, b);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_run; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void fill_u8_kernel(uint8_t* p, uint8_t v, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void fill_i32_kernel(int* p, int v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void fill_f64_kernel(double* p, double v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
