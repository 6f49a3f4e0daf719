// xfb200: plan object, C-ABI entry points (include/xfb200.h) and the device-resident MTIP loop driver.
#include "../../include/xfb200.h"
#include "common.cuh"
#include "pointwise.cuh"
#include "fft.cuh"
#include "legendre.cuh"
#include "gemm.cuh"
#include "procrustes.cuh"

#ifndef RU_WAVES
#define RU_WAVES 32   // real_update / shrink-wrap reductions: CTAs per launch = RU_WAVES x 148 (split over the runs)
#endif
#include "polar.cuh"

#include <algorithm>
#include <map>
#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: named ranges per stage of an iteration (visible in nsys / ncu timelines)

struct NvtxRange { explicit NvtxRange(const char* name) { nvtxRangePushA(name); } ~NvtxRange() { nvtxRangePop(); } };

thread_local std::string g_xfb_err;

enum ProfGroup { PG_FFT = 0, PG_LEGENDRE, PG_HANKEL, PG_PROC_GEMM, PG_PROC_JACOBI, PG_PROC_PACK, PG_POINTWISE, PG_REAL_UPDATE, PG_MISC, PG_SHT, PG_COUNT };
static const char* kProfNames[PG_COUNT] = {"fft_phi", "legendre", "hankel", "procrustes_gemm", "procrustes_jacobi",
                                           "procrustes_pack", "pointwise", "real_update", "misc", "sht"};

struct ProfEvent { cudaEvent_t a, b; int group; };

struct xfb_plan {
    int dims = 3;                                   // 3: spherical (SHT + spherical Hankel), 2: polar (circular harmonics + polar Hankel)
    int L = 0, n_r = 0, n_theta = 0, n_phi = 0, max_batch = 0, NLM = 0, M2 = 0, K2 = 0, NP = 0;
    int n_hankel = 0;                               // number of radial weight matrices: L+1 (3-D) or n_phi (2-D, one per DFT index)
    int wt_div = 1;                                 // quadrature weight index = point index / wt_div
    // 2-D projection constants
    double2* v2d = nullptr; double2* unk2d = nullptr; int n_orders2d = 0, so_order2d = -1;
    // 2-D DFT as two real DMMA GEMMs per transform: cos / sin matrices [2][N][ldw], row-major temporaries [B*G]
    double* dft_cs = nullptr; int dft_ldw = 0; double2 *T2a = nullptr, *T2b = nullptr;
    // folded DFT (polar.cuh): cos / sin half matrices [2][H][H], H = M+1 padded to 16; row buffers of max(N, 2H) complex per shell
    double* dft_fold = nullptr; int dft_H = 0, dft_fold_on = 1; long long t2_run = 0;
    std::map<int, std::pair<HankelTile*, int>> dft_fold_tiles;
    std::map<int, std::pair<HankelTile*, int>> dft_tiles;     // per number of shells S: tiles for the C pass and the S pass
    int hankel_skip = 0, hankel_n_sum = 0;
    double hk_fwd_scale = 0, hk_inv_scale = 0;
    long long G = 0, C = 0;
    // tables
    double2* tw = nullptr;
    double *FE = nullptr, *FO = nullptr, *IE = nullptr, *IO = nullptr, *hankel_w = nullptr, *int_wt = nullptr, *q_pts = nullptr;
    // workspaces
    double2 *A0 = nullptr, *C0 = nullptr, *C1 = nullptr, *W0 = nullptr, *W1 = nullptr, *W2 = nullptr;
    double2 *A0s = nullptr, *C0s = nullptr, *rt0 = nullptr;   // shell-0 side path of the fused ft_stab step
    int fused_ft_stab = 1;
    bool leg2 = false;                              // v3 Legendre kernels (K2 <= 64, NP <= 64)
    bool leg3_big = false;                          // ... the <KS 16, NCG 4> instantiation (K2 > 32 or NP > 32)
    int half_spectrum = 1;                          // real intensity fields: transform only the m >= 0 half (3-D, v2 Legendre)
    // Chunked transform (experiment, OFF by default): a transform is cut into chunks of sht_chunk runs; chunk i runs its phi-FFT
    // and its Legendre kernel back to back on stream i % sht_streams through the same slot of A0, so that the m-major
    // intermediate `a` could stay inside the 126 MB L2.  Measured at 128 runs (profiles/r02a_sht_chunk_sweep.md): 9.2 - 13.5 ms
    // per step for the six transforms against 8.24 ms unchunked -- the smaller launches lose more than the L2 hits gain.
    int sht_chunk = 0, sht_streams = 3;
    // TMA-fed operand tiles (hankel3_tma_kernel: per-row bulk copies + mbarrier).  OFF by default: measured at 128 runs the Hankel
    // group takes 3.57 ms per step against 2.42 ms with the cp.async ring (profiles/r02g_hankel_tma_ab.md) -- 80 row-sized bulk
    // copies per K chunk issued by one warp cost more than 256 threads issuing six 16-byte cp.async each.
    int hankel_tma = 0;
    int leg_min_groups = 0;                         // >0: at least that many shell groups per Legendre CTA (measured at 16 / 32 runs: 8 and 16 are not faster than the wave rule)
    cudaStream_t sht_side[4] = {}; cudaEvent_t sht_fork = nullptr, sht_join[4] = {};
    long long launches_side = 0;
    // host-buffer pipeline (xfb_mtip_step_host)
    cudaStream_t s_in = nullptr, s_out = nullptr; cudaEvent_t ev_start = nullptr; std::vector<cudaEvent_t> ev_in, ev_comp;
    double2* stage_out = nullptr; int host_chunk = 16;
    HankelTile* hk_tiles = nullptr; int hk_tiles_n = 0;
    std::map<int, std::pair<HankelTile*, int>> hk_cache;     // tile lists per batch size (the host pipeline uses several chunk sizes)

    // projection
    bool has_proj = false;
    std::vector<ProcOrder> orders;
    std::vector<int> ncols_all;                     // columns of V_l as given, per order (also for zeroed / pass-through orders)
    ProcOrder* orders_dev = nullptr;
    int *kind_dev = nullptr, *act_index_dev = nullptr;
    uint8_t* radial_mask_dev = nullptr;
    double* v0_dev = nullptr;
    double inv_sqrt_np = 1.0, sv_cutoff = 1e-15; int max_sweeps = 40;
    double *pd_dev = nullptr, *vt_dev = nullptr;
    // fxs_unknowns on request (xfb_get_unknowns): identity accumulator table, single-run scratch, I_00 of the last projection
    double *pp = nullptr, *gn_u = nullptr, *pp_u = nullptr, *sigma_u = nullptr; double2* i00 = nullptr;
    long long proj_calls = 0, unk_stamp = -1; int unk_run = -1, last_proj_nb = 0;
    long long xt_run = 0, g_run = 0, vw_run = 0;
    double *xt = nullptr, *tt = nullptr, *g = nullptr, *gn = nullptr, *vw = nullptr, *sigma = nullptr;
    int* sweeps_dev = nullptr;
    GemmProblem *gemmM_dev = nullptr, *gemmT_dev = nullptr, *gemmY_dev = nullptr; int *gemmM_tp = nullptr, *gemmT_tp = nullptr, *gemmY_tp = nullptr;
    int gemmY_tiles = 0, gemmY_tiles_run = 0;
    int gemmM_tiles = 0, gemmT_tiles = 0, gemm_nb = -1, gemmM_tiles_run = 0, gemmT_tiles_run = 0;
    size_t jacobi_smem = 0; int n_sm = 148; bool jacobi_big = false; int* jac_counter = nullptr;
    int sig_ld = 0;                                 // singular values per (run, order): max(N_r, largest n_cols)
    int jac_split = 1, jac_n_big = 0;               // orders [0, jac_n_big) have 2l+1 > 64: one problem per SM; the rest two per SM
    // degree-2 invariants B_l = I_l I_l^H of a batch and deg2_invariant_l2_diff (procrustes.cuh)
    double *d2_x = nullptr, *d2_b = nullptr, *d2_ref = nullptr, *d2_norm = nullptr, *d2_hist = nullptr;
    GemmProblem* d2_gemm = nullptr; int* d2_tp = nullptr; int d2_tiles_run = 0; int d2_hist_cap = 0; bool d2_metric = false;
    // real projection
    bool has_real = false; RealDesc rd{}; uint8_t* init_support_dev = nullptr; double2* avg_mean = nullptr;
    // loop state
    double2 *rho_pool = nullptr, *rh_pool = nullptr; uint8_t* mask_pool = nullptr;
    LoopState ls{}; int* ls_ints = nullptr; double* ls_dbl = nullptr;
    // Two halves of the batch in flight on two streams (xfb_mtip_iterate): while one half sits in the latency-bound Jacobi kernel
    // on a part of the SMs, the HBM-bound transform kernels of the other half run on the rest.  OFF by default: measured at 128
    // runs (profiles/r02e_dual_stream_sweep.md) the step takes 22.98 ms at the best SM split against 22.93 ms on one stream --
    // the HBM-bound kernels need (nearly) all SMs to reach their bandwidth, so SM time is conserved and nothing is gained.
    int dual = 0, dual_min = 32, dual_big_sms = 50, dual_small_sms = 20; bool dual_active = false;
    cudaStream_t s_half = nullptr; cudaEvent_t ev_hfork = nullptr, ev_hjoin = nullptr, ev_stagger = nullptr; cudaEvent_t stagger_pending = nullptr;
    cudaEvent_t jac_fork[2] = {}, jac_join[2] = {};
    // CUDA graph of one whole iteration per (method, ft_stab, non-FXS, batch): captured on the second call of a kind (the first runs
    // eagerly and warms every lazily built table), replayed afterwards with the step's scalars in device memory (IterParams)
    int use_graph = 1; long long graph_replays = 0; IterParams* iter_params = nullptr; const IterParams* ip_active = nullptr;
    std::map<int, cudaGraphExec_t> graphs; std::map<int, int> graph_warm, graph_launches; bool graph_broken = false;
    cudaStream_t s_graph = nullptr; cudaEvent_t ev_gfork = nullptr, ev_gjoin = nullptr;      // the caller's stream may be the legacy default stream, which cannot be captured
    int run_base = 0, ctx = 0;                      // scratch of the current enqueue starts at this run; stream / event set in use
    int n_batch = 0, it_done = 0, outer_it = 0; bool non_fxs = false; double *fix_int = nullptr, *fix_cand = nullptr; double *partial = nullptr, *err = nullptr, *mm = nullptr; int red_blocks = 0;
    bool loop_alloc = false;
    int64_t launches = 0, bytes = 0;
    // profiling
    bool prof = false; std::vector<ProfEvent> prof_events; double prof_ms[PG_COUNT] = {}; int64_t prof_n[PG_COUNT] = {};
};

// ---- launch bookkeeping ---------------------------------------------------------------
static inline void prof_begin(xfb_plan* p, int group, cudaStream_t st) {
    if (!p->prof) return;
    ProfEvent e; e.group = group;
    cudaEventCreate(&e.a); cudaEventCreate(&e.b);
    cudaEventRecord(e.a, st);
    p->prof_events.push_back(e);
}
static inline void prof_end(xfb_plan* p, cudaStream_t st) {
    if (!p->prof) return;
    cudaEventRecord(p->prof_events.back().b, st);
}
#define XFB_LAUNCH(p, group, st, ...)            \
    do {                                         \
        prof_begin((p), (group), (st));          \
        __VA_ARGS__;                             \
        prof_end((p), (st));                     \
        (p)->launches++;                         \
        XFB_CUDA(cudaGetLastError());            \
    } while (0)

template <typename T>
static int dev_alloc(xfb_plan* p, T** ptr, size_t n) {
    XFB_CUDA(cudaMalloc((void**)ptr, n * sizeof(T)));
    p->bytes += (int64_t)(n * sizeof(T));
    return 0;
}
template <typename T>
static int dev_upload(xfb_plan* p, T** ptr, const T* host, size_t n) {
    if (dev_alloc(p, ptr, n)) return 1;
    XFB_CUDA(cudaMemcpy(*ptr, host, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

static inline SlotView flat_view(const double2* ptr, long long run_stride) {
    SlotView v; v.base = const_cast<double2*>(ptr); v.slot = nullptr; v.slot_stride = 0; v.run_stride = run_stride; return v;
}
static inline int ew_blocks(long long n) { return (int)std::min<long long>(cdiv64(n, 256 * 4), 148 * 16); }

// Per-run scratch is indexed by the ABSOLUTE run while a part of the batch is enqueued: the pointers of the plan are moved to
// run b0 for the duration of the enqueue (kernel arguments are taken by value at launch) and restored afterwards.  Parts of
// the batch can then be in flight on different streams at the same time (two halves, host-pipeline chunks), and the retained
// projection data (M^T, I_00) of every run stay where xfb_get_unknowns looks for them.
struct ScratchShift {
    xfb_plan* p; int b0;
    template <typename T> static void mv(T*& q, long long n) { if (q) q += n; }
    void apply(long long s) {
        const long long b = (long long)b0 * s;
        mv(p->W0, b * p->G); mv(p->W1, b * p->G); mv(p->W2, b * p->G); mv(p->C0, b * p->C); mv(p->C1, b * p->C);
        mv(p->T2a, b * p->t2_run); mv(p->T2b, b * p->t2_run);
        if (p->dims == 3) { mv(p->A0, b * p->n_r * p->M2 * p->n_theta); mv(p->A0s, b * p->M2 * p->n_theta); mv(p->C0s, b * p->NLM); mv(p->rt0, b * p->n_theta * p->n_phi); }
        mv(p->xt, b * p->xt_run); mv(p->tt, b * p->xt_run); mv(p->g, b * p->g_run); mv(p->gn, b * p->g_run); mv(p->pp, b * p->g_run);
        mv(p->vw, b * p->vw_run); mv(p->sigma, b * (long long)p->orders.size() * p->sig_ld); mv(p->sweeps_dev, b * (long long)p->orders.size());
        mv(p->i00, b * p->n_r); mv(p->partial, b * p->red_blocks * 2); mv(p->avg_mean, b * p->n_r); mv(p->unk2d, b * p->n_orders2d);
        mv(p->d2_x, b * p->NLM * p->n_r); mv(p->d2_b, b * (p->L + 1) * p->n_r * p->n_r); mv(p->fix_int, b * p->G);
    }
    ScratchShift(xfb_plan* p_, int b0_) : p(p_), b0(b0_) { apply(+1); p->run_base += b0; }
    ~ScratchShift() { p->run_base -= b0; apply(-1); }
};


// captured iterations bake the plan's constants (projection, real-space options, switches) into their kernel arguments
static void graphs_invalidate(xfb_plan* p) {
    for (auto& kv : p->graphs) cudaGraphExecDestroy(kv.second);
    p->graphs.clear(); p->graph_warm.clear(); p->graph_launches.clear();
}

extern "C" {

const char* xfb_last_error(void) { return g_xfb_err.c_str(); }

int xfb_device_count(int* n) { XFB_CUDA(cudaGetDeviceCount(n)); return 0; }
int xfb_set_device(int dev) { XFB_CUDA(cudaSetDevice(dev)); return 0; }

int xfb_plan_destroy(xfb_plan* p);
namespace { struct PlanGuard { xfb_plan* p; ~PlanGuard() { if (p) xfb_plan_destroy(p); } }; }   // failure paths free the plan and its buffers

int xfb_plan_create(xfb_plan** out, const xfb_plan_desc* d) {
    if (!out || !d) XFB_FAIL("null argument");
    const int dims = (d->dimensions == 2) ? 2 : 3;
    if (dims == 3) {
        if (d->n_theta % 8 != 0) XFB_FAIL("n_theta=%d must be a multiple of 8", d->n_theta);
        if (d->n_theta <= d->l_max) XFB_FAIL("n_theta=%d must exceed l_max=%d (exact Gauss quadrature)", d->n_theta, d->l_max);
        if (d->n_phi <= 2 * d->l_max) XFB_FAIL("n_phi=%d must exceed 2*l_max", d->n_phi);
        if (d->n_phi < 16 || d->n_phi > 512 || (d->n_phi & (d->n_phi - 1))) XFB_FAIL("n_phi=%d must be a power of two in [16,512]", d->n_phi);
    } else {
        if (d->n_theta != 1) XFB_FAIL("2-D plan: n_theta must be 1 (got %d)", d->n_theta);
        if (d->n_phi < 3 || d->n_phi > 1023) XFB_FAIL("2-D plan: n_phi=%d outside [3,1023]", d->n_phi);
        if (d->n_phi != 2 * d->l_max + 1) XFB_FAIL("2-D plan: n_phi=%d must be 2*max_order+1 (harmonic_transforms.py:44-47)", d->n_phi);
    }
    if (d->max_batch < 1) XFB_FAIL("max_batch must be >= 1");
    int ndev = 0;
    XFB_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev == 0) XFB_FAIL("no CUDA device: xfb200 has no CPU fallback");
    xfb_plan* p = new xfb_plan();
    PlanGuard guard{p};
    p->dims = dims;
    p->L = d->l_max; p->n_r = d->n_r; p->n_theta = d->n_theta; p->n_phi = d->n_phi; p->max_batch = d->max_batch;
    p->hankel_skip = d->hankel_skip; p->hankel_n_sum = d->hankel_n_sum;
    p->hk_fwd_scale = d->hankel_fwd_scale; p->hk_inv_scale = d->hankel_inv_scale;
    p->G = (long long)p->n_r * p->n_theta * p->n_phi;
    if (dims == 2) {
        // polar plan: coefficient arrays are [n_phi][S]; no Legendre stage, no phi-FFT tables
        p->NLM = p->n_phi; p->M2 = p->n_phi; p->n_hankel = p->n_phi; p->wt_div = 1; p->fused_ft_stab = 0;
        p->C = (long long)p->NLM * p->n_r;
        if (dev_upload(p, &p->hankel_w, d->hankel_w, (size_t)p->n_hankel * p->hankel_n_sum * p->n_r)) return 1;
        if (dev_upload(p, &p->int_wt, d->int_weight, (size_t)p->n_r * p->n_phi)) return 1;      // per point: trapz in phi (PolarIntegrator)
        if (dev_upload(p, &p->q_pts, d->q_points, (size_t)p->n_r)) return 1;
        const size_t B2 = p->max_batch;
        if (dev_alloc(p, &p->C0, B2 * p->C)) return 1;
        if (dev_alloc(p, &p->C1, B2 * p->C)) return 1;
        if (dev_alloc(p, &p->W0, B2 * p->G)) return 1;
        if (dev_alloc(p, &p->W1, B2 * p->G)) return 1;
        if (dev_alloc(p, &p->W2, B2 * p->G)) return 1;
        {   // cos / sin matrices of the DFT, extended precision, rows padded to an even leading dimension
            const int N = p->n_phi;
            p->dft_ldw = (N + 1) & ~1;
            std::vector<double> cs((size_t)2 * N * p->dft_ldw, 0.0);
            for (int a = 0; a < N; ++a)
                for (int b = 0; b < N; ++b) {
                    const long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double)(((long long)a * b) % N) / (long double)N;
                    cs[(size_t)a * p->dft_ldw + b] = (double)cosl(ang);
                    cs[(size_t)(N + a) * p->dft_ldw + b] = (double)sinl(ang);
                }
            if (dev_upload(p, &p->dft_cs, cs.data(), cs.size())) return 1;
            {   // folded transform: cos / sin of 2 pi (j k mod N) / N for j, k = 0 .. M, zero padded to H
                const int M = N / 2, H = (M + 1 + 15) & ~15;
                p->dft_H = H;
                std::vector<double> fw((size_t)2 * H * H, 0.0);
                for (int a = 0; a <= M; ++a)
                    for (int b = 0; b <= M; ++b) {
                        const long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double)(((long long)a * b) % N) / (long double)N;
                        fw[(size_t)a * H + b] = (double)cosl(ang);
                        if (a > 0 && b > 0) fw[(size_t)H * H + (size_t)a * H + b] = (double)sinl(ang);
                    }
                if (dev_upload(p, &p->dft_fold, fw.data(), fw.size())) return 1;
                if (const char* e = getenv("XFB_DFT_FOLD")) p->dft_fold_on = atoi(e) != 0;
                p->t2_run = (long long)p->n_r * std::max(N, 2 * H);
            }
            if (dev_alloc(p, &p->T2a, B2 * (size_t)p->t2_run)) return 1;
            if (dev_alloc(p, &p->T2b, B2 * (size_t)p->t2_run)) return 1;
            XFB_CUDA(cudaFuncSetAttribute(hankel2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hankel2_smem()));
        }
        XFB_CUDA(cudaFuncSetAttribute(dft2d_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dft_smem(p->n_phi)));
        XFB_CUDA(cudaFuncSetAttribute(dft2d_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dft_smem(p->n_phi)));
        guard.p = nullptr;
        *out = p;
        return 0;
    }
    p->NLM = (p->L + 1) * (p->L + 1); p->M2 = 2 * p->L + 1; p->K2 = p->n_theta / 2;
    p->NP = ((p->L / 2 + 1) + 7) / 8 * 8;
    p->n_hankel = p->L + 1; p->wt_div = p->n_phi;
    p->C = (long long)p->NLM * p->n_r;
    const size_t tab = (size_t)(p->L + 1) * p->K2 * p->NP;
    if (d->legendre_len != (int64_t)(4 * tab)) { XFB_FAIL("legendre table length %lld != %lld", (long long)d->legendre_len, (long long)(4 * tab)); }
    // twiddles exp(-2 pi i k / n) in extended precision
    std::vector<double2> tw(p->n_phi);
    for (int k = 0; k < p->n_phi; ++k) {
        const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)p->n_phi;
        tw[k] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
    if (dev_upload(p, &p->tw, tw.data(), tw.size())) return 1;
    if (dev_upload(p, &p->FE, d->legendre, tab)) return 1;
    if (dev_upload(p, &p->FO, d->legendre + tab, tab)) return 1;
    if (dev_upload(p, &p->IE, d->legendre + 2 * tab, tab)) return 1;
    if (dev_upload(p, &p->IO, d->legendre + 3 * tab, tab)) return 1;
    if (dev_upload(p, &p->hankel_w, d->hankel_w, (size_t)(p->L + 1) * p->hankel_n_sum * p->n_r)) return 1;
    if (dev_upload(p, &p->int_wt, d->int_weight, (size_t)p->n_r * p->n_theta)) return 1;
    if (dev_upload(p, &p->q_pts, d->q_points, (size_t)p->n_r)) return 1;
    const size_t B = p->max_batch;
    if (dev_alloc(p, &p->A0, B * p->n_r * p->M2 * p->n_theta)) return 1;
    if (dev_alloc(p, &p->C0, B * p->C)) return 1;
    if (dev_alloc(p, &p->C1, B * p->C)) return 1;
    if (dev_alloc(p, &p->W0, B * p->G)) return 1;
    if (dev_alloc(p, &p->W1, B * p->G)) return 1;
    if (dev_alloc(p, &p->W2, B * p->G)) return 1;
    if (dev_alloc(p, &p->A0s, B * p->M2 * p->n_theta)) return 1;
    if (dev_alloc(p, &p->C0s, B * p->NLM)) return 1;
    if (dev_alloc(p, &p->rt0, B * p->n_theta * p->n_phi)) return 1;
    if (const char* e = getenv("XFB_SHT_CHUNK")) p->sht_chunk = std::max(0, atoi(e));          // experiments: sweep without rebuilding
    if (const char* e = getenv("XFB_SHT_STREAMS")) p->sht_streams = std::min(4, std::max(1, atoi(e)));
    if (const char* e = getenv("XFB_LEG_MINGROUPS")) p->leg_min_groups = std::max(0, atoi(e));
    if (const char* e = getenv("XFB_HANKEL_TMA")) p->hankel_tma = atoi(e) != 0;
    if (const char* e = getenv("XFB_GRAPH")) p->use_graph = atoi(e) != 0;
    if (const char* e = getenv("XFB_DUAL")) p->dual = atoi(e) != 0;
    if (const char* e = getenv("XFB_DUAL_BIG")) p->dual_big_sms = std::max(1, atoi(e));
    if (const char* e = getenv("XFB_DUAL_SMALL")) p->dual_small_sms = std::max(1, atoi(e));
    if (const char* e = getenv("XFB_DUAL_MIN")) p->dual_min = std::max(2, atoi(e));
    p->leg2 = (p->n_theta / 2 <= 64 && p->NP <= 64 && p->n_theta % 4 == 0);
    p->leg3_big = p->leg2 && (p->n_theta / 2 > 32 || p->NP > 32);
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&p->n_sm, cudaDevAttrMultiProcessorCount, dev); }
    if (p->leg2 && !p->leg3_big) XFB_CUDA(cudaFuncSetAttribute(legendre3_forward_kernel<LEG2_FR, LEG2_FST, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)legendre3_fwd_smem(p->n_theta)));
    if (p->leg3_big) XFB_CUDA(cudaFuncSetAttribute(legendre3_forward_kernel<LEG2_FR, LEG3_BIG_ST, 16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)legendre3_fwd_smem(p->n_theta, LEG3_BIG_ST)));
    if (p->leg2 && !p->leg3_big) XFB_CUDA(cudaFuncSetAttribute(legendre3_inverse_kernel<LEG2_IR, LEG2_IST, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)legendre3_inv_smem(p->NP)));
    if (p->leg3_big) XFB_CUDA(cudaFuncSetAttribute(legendre3_inverse_kernel<LEG2_IR, LEG3_BIG_ST, 16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)legendre3_inv_smem(p->NP, LEG3_BIG_ST)));
    XFB_CUDA(cudaFuncSetAttribute(legendre_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)legendre_fwd_smem(p->n_theta)));
    XFB_CUDA(cudaFuncSetAttribute(legendre_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)legendre_inv_smem(p->n_theta, p->NP)));
    guard.p = nullptr;
    *out = p;
    return 0;
}

int xfb_plan_destroy(xfb_plan* p) {
    if (!p) return 0;
    void* ptrs[] = {p->tw, p->FE, p->FO, p->IE, p->IO, p->hankel_w, p->int_wt, p->q_pts, p->A0, p->C0, p->C1, p->W0, p->W1, p->W2, p->A0s, p->C0s, p->rt0, p->stage_out,
                    p->v2d, p->unk2d, p->dft_cs, p->dft_fold, p->T2a, p->T2b, p->orders_dev, p->kind_dev, p->act_index_dev, p->radial_mask_dev, p->v0_dev, p->pd_dev, p->vt_dev,
                    p->xt, p->tt, p->g, p->gn, p->vw, p->sigma, p->sweeps_dev, p->jac_counter, p->pp, p->gn_u, p->pp_u, p->sigma_u, p->i00, p->gemmY_dev, p->gemmY_tp, p->gemmM_dev, p->gemmT_dev, p->gemmM_tp, p->gemmT_tp,
                    p->init_support_dev, p->avg_mean, p->fix_int, p->fix_cand, p->d2_x, p->d2_b, p->d2_ref, p->d2_norm, p->d2_hist, p->d2_gemm, p->d2_tp, p->rho_pool, p->rh_pool, p->mask_pool, p->ls_ints, p->ls_dbl, p->partial, p->err, p->mm};
    for (void* q : ptrs) if (q) cudaFree(q);
    for (auto& kv : p->hk_cache) cudaFree(kv.second.first);
    for (auto& kv : p->dft_tiles) cudaFree(kv.second.first);
    for (auto& kv : p->dft_fold_tiles) cudaFree(kv.second.first);
    for (auto& kv : p->graphs) cudaGraphExecDestroy(kv.second);
    if (p->iter_params) cudaFree(p->iter_params);
    if (p->s_graph) cudaStreamDestroy(p->s_graph);
    for (cudaEvent_t e : {p->ev_gfork, p->ev_gjoin}) if (e) cudaEventDestroy(e);
    for (auto& e : p->prof_events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (auto e : p->ev_in) cudaEventDestroy(e);
    for (auto e : p->ev_comp) cudaEventDestroy(e);
    if (p->ev_start) cudaEventDestroy(p->ev_start);
    if (p->s_in) cudaStreamDestroy(p->s_in);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    for (int i = 1; i < 4; ++i) { if (p->sht_side[i]) cudaStreamDestroy(p->sht_side[i]); if (p->sht_join[i]) cudaEventDestroy(p->sht_join[i]); }
    if (p->sht_fork) cudaEventDestroy(p->sht_fork);
    if (p->s_half) cudaStreamDestroy(p->s_half);
    for (cudaEvent_t e : {p->ev_hfork, p->ev_hjoin, p->ev_stagger, p->jac_fork[0], p->jac_fork[1], p->jac_join[0], p->jac_join[1]}) if (e) cudaEventDestroy(e);
    delete p;
    return 0;
}

int64_t xfb_plan_workspace_bytes(const xfb_plan* p) { return p ? p->bytes : 0; }
int64_t xfb_plan_launch_count(const xfb_plan* p) { return p ? p->launches : 0; }

int xfb_profile_enable(xfb_plan* p, int32_t on) {
    p->prof = on != 0;
    for (auto& e : p->prof_events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    p->prof_events.clear();
    for (int i = 0; i < PG_COUNT; ++i) { p->prof_ms[i] = 0; p->prof_n[i] = 0; }
    return 0;
}
int xfb_profile_read(xfb_plan* p, int32_t n_max, char* names, double* ms, int64_t* launches, int32_t* n_out) {
    XFB_CUDA(cudaDeviceSynchronize());
    for (auto& e : p->prof_events) {
        float t = 0;
        XFB_CUDA(cudaEventElapsedTime(&t, e.a, e.b));
        p->prof_ms[e.group] += t; p->prof_n[e.group]++;
        cudaEventDestroy(e.a); cudaEventDestroy(e.b);
    }
    p->prof_events.clear();
    int n = std::min<int>(n_max, PG_COUNT);
    for (int i = 0; i < n; ++i) {
        strncpy(names + i * 32, kProfNames[i], 31); names[i * 32 + 31] = 0;
        ms[i] = p->prof_ms[i]; launches[i] = p->prof_n[i];
    }
    *n_out = n;
    return 0;
}

}  // extern "C"

// ---- internal building blocks -----------------------------------------------------------
static int transpose_i(xfb_plan* p, const double2* in, double2* out, int rows, int cols, cudaStream_t st);

// 2-D circular-harmonic DFT of S rows as two passes of the DMMA GEMM kernel: rows [S][N] complex times the real cos and
// sin matrices, out = scale (X C -+ i X S)  (forward: exp(-i..), inverse: exp(+i..)).
static int dft2d_gemm_i(xfb_plan* p, const double2* rows_in, double2* rows_out, int S, int inverse, double scale, cudaStream_t st) {
    const int N = p->n_phi;
    auto hit = p->dft_tiles.find(S);
    if (hit == p->dft_tiles.end()) {
        std::vector<HankelTile> tiles;
        for (int pass = 0; pass < 2; ++pass)                      // matrix 0 = cos (phase 1), matrix 1 = sin (phase -+i)
            for (int r = 0; r < S; r += HK_BM) tiles.push_back(HankelTile{pass, r, S, pass});
        HankelTile* dev = nullptr;
        XFB_CUDA(cudaMalloc((void**)&dev, tiles.size() * sizeof(HankelTile)));
        XFB_CUDA(cudaMemcpyAsync(dev, tiles.data(), tiles.size() * sizeof(HankelTile), cudaMemcpyHostToDevice, st));
        XFB_CUDA(cudaStreamSynchronize(st));
        hit = p->dft_tiles.emplace(S, std::make_pair(dev, (int)tiles.size() / 2)).first;
    }
    const HankelTile* tl = hit->second.first;
    const int nt = hit->second.second;
    dim3 g(nt, cdiv(N, HK_BN));
    // phase of the sin pass: forward multiplies by -i (ph = 1, inverse = 0), inverse by +i (ph = 1, inverse = 1)
    XFB_LAUNCH(p, PG_FFT, st, hankel2_kernel<<<g, 256, hankel2_smem(), st>>>(rows_in, rows_out, p->dft_cs, tl, N, N, 0, scale, inverse, p->dft_ldw, 0));
    XFB_LAUNCH(p, PG_FFT, st, hankel2_kernel<<<g, 256, hankel2_smem(), st>>>(rows_in, rows_out, p->dft_cs, tl + nt, N, N, 0, scale, inverse, p->dft_ldw, 1));
    return 0;
}


// Folded circular-harmonic transform (polar.cuh): fold -> ONE launch of the DMMA GEMM kernel over the stacked [2 S][H] row array
// (rows [0, S): e rows x cos matrix, rows [S, 2 S): o rows x sin matrix) -> combine.  The row buffers are T2a (e | o) and T2b (A | B).
static int dft2d_fold_gemm_i(xfb_plan* p, int S, cudaStream_t st) {
    const int H = p->dft_H;
    auto hit = p->dft_fold_tiles.find(S);
    if (hit == p->dft_fold_tiles.end()) {
        std::vector<HankelTile> tiles;
        for (int blk = 0; blk < 2; ++blk)
            for (int r = 0; r < S; r += HK_BM) tiles.push_back(HankelTile{blk, blk * S + r, (blk + 1) * S, 0});
        HankelTile* dev = nullptr;
        XFB_CUDA(cudaMalloc((void**)&dev, tiles.size() * sizeof(HankelTile)));
        XFB_CUDA(cudaMemcpyAsync(dev, tiles.data(), tiles.size() * sizeof(HankelTile), cudaMemcpyHostToDevice, st));
        XFB_CUDA(cudaStreamSynchronize(st));
        hit = p->dft_fold_tiles.emplace(S, std::make_pair(dev, (int)tiles.size())).first;
    }
    dim3 g(hit->second.second, cdiv(H, HK_BN));
    XFB_LAUNCH(p, PG_FFT, st, hankel2_kernel<<<g, 256, hankel2_smem(), st>>>(p->T2a, p->T2b, p->dft_fold, hit->second.first, H, H, 0, 1.0, 0, H, 0));
    return 0;
}
static int dft2d_forward_folded_i(xfb_plan* p, SlotView in, int shells_per_run, double2* c_out, int S, cudaStream_t st) {
    const int N = p->n_phi, H = p->dft_H;
    XFB_LAUNCH(p, PG_FFT, st, dft_fold_rows_kernel<<<ew_blocks((long long)S * H), 256, 0, st>>>(in, shells_per_run, p->T2a, S, N, H));
    if (dft2d_fold_gemm_i(p, S, st)) return 1;
    XFB_LAUNCH(p, PG_FFT, st,
               dft_combine_to_orders_kernel<<<dim3(cdiv(N / 2 + 1, 32), cdiv(S, 32)), dim3(32, 8), 0, st>>>(p->T2b, c_out, S, N, H, 1.0 / N));
    return 0;
}
static int dft2d_inverse_folded_i(xfb_plan* p, const double2* c_in, double2* grid_out, int S, int herm, cudaStream_t st) {
    const int N = p->n_phi, H = p->dft_H;
    XFB_LAUNCH(p, PG_FFT, st, dft_fold_orders_kernel<<<dim3(cdiv(H, 32), cdiv(S, 32)), dim3(32, 8), 0, st>>>(c_in, p->T2a, S, N, H, herm));
    if (dft2d_fold_gemm_i(p, S, st)) return 1;
    XFB_LAUNCH(p, PG_FFT, st, dft_combine_to_rows_kernel<<<ew_blocks((long long)S * (N / 2 + 1)), 256, 0, st>>>(p->T2b, grid_out, S, N, H));
    return 0;
}

// ---- v3 Legendre launch + the chunked (L2-resident intermediate) transform -------------------------------------------
// CTAs per order m for a launch over S shells: every CTA should walk over several shell groups (the cp.async ring needs a
// few groups to hide the load latency and the table fragments are loaded once per CTA), and the whole grid should not
// exceed LEG3_WAVES waves of the resident CTAs.  share: number of streams that run such launches concurrently.
static int leg3_grid_x(const xfb_plan* p, int S, int pos_only, int share) {
    const int groups = cdiv(S, pos_only ? LEG2_FR : LEG2_FR / 2);
    const int resident = (p->leg3_big ? 1 : LEG3_MINB) * p->n_sm;
    const int cap = std::max(1, (resident * LEG3_WAVES) / ((p->L + 1) * std::max(1, share)));
    int gx = std::min(groups, cap);
    if (share > 1) gx = std::max(1, std::min(gx, cdiv(groups, 4)));       // chunked launches: at least 4 groups per CTA
    // small batches (the 16 / 32-run shards of 8 / 4 GPUs): at least leg_min_groups shell groups per CTA, so the table fragments
    // and the pipeline fill are amortised, unless that leaves less than one wave of CTAs
    if (p->leg_min_groups > 0) gx = std::min(gx, std::max(cdiv(resident, p->L + 1), cdiv(groups, p->leg_min_groups)));
    return std::max(1, gx);
}
// forward: a [S][M2][n_theta] -> c (+ chunk offset, row stride c_stride);  inverse: c -> a
static int launch_legendre3(xfb_plan* p, bool forward, double2* a, double2* c, int S, long long c_stride, int pos_only, int gx, cudaStream_t st) {
    dim3 g2(gx, p->L + 1);
    if (forward) {
        if (p->leg3_big)
            legendre3_forward_kernel<LEG2_FR, LEG3_BIG_ST, 16, 4><<<g2, 256, legendre3_fwd_smem(p->n_theta, LEG3_BIG_ST), st>>>(a, c, p->FE, p->FO, S, p->L, p->n_theta, p->NP, pos_only, c_stride);
        else
            legendre3_forward_kernel<LEG2_FR, LEG2_FST, 8, 2><<<g2, 128, legendre3_fwd_smem(p->n_theta), st>>>(a, c, p->FE, p->FO, S, p->L, p->n_theta, p->NP, pos_only, c_stride);
    } else {
        if (p->leg3_big)
            legendre3_inverse_kernel<LEG2_IR, LEG3_BIG_ST, 16, 4><<<g2, 256, legendre3_inv_smem(p->NP, LEG3_BIG_ST), st>>>(c, a, p->IE, p->IO, S, p->L, p->n_theta, p->NP, pos_only, c_stride);
        else
            legendre3_inverse_kernel<LEG2_IR, LEG2_IST, 8, 2><<<g2, 128, legendre3_inv_smem(p->NP), st>>>(c, a, p->IE, p->IO, S, p->L, p->n_theta, p->NP, pos_only, c_stride);
    }
    XFB_CUDA(cudaGetLastError());
    return 0;
}
static inline int sht_chunk_shells(const xfb_plan* p) { return p->sht_chunk * p->n_r; }
static bool sht_chunkable(const xfb_plan* p, int S, int shells_per_run) {
    if (p->dims != 3 || !p->leg2 || p->sht_chunk < 1 || p->sht_streams < 1 || shells_per_run < 1) return false;
    const int cs = sht_chunk_shells(p);
    return S > cs && cs % shells_per_run == 0 && S % shells_per_run == 0 && 2 * p->sht_chunk <= p->max_batch;
}
static int sht_streams_init(xfb_plan* p) {
    if (p->sht_fork) return 0;
    XFB_CUDA(cudaEventCreateWithFlags(&p->sht_fork, cudaEventDisableTiming));
    for (int i = 1; i < 4; ++i) {
        XFB_CUDA(cudaStreamCreateWithFlags(&p->sht_side[i], cudaStreamNonBlocking));
        XFB_CUDA(cudaEventCreateWithFlags(&p->sht_join[i], cudaEventDisableTiming));
    }
    return 0;
}
struct ShtFork {            // streams of one chunked transform: stream 0 is the caller's, the others are forked from / joined to it
    xfb_plan* p; cudaStream_t st; int ns; bool used[4];
    cudaStream_t stream(int k) const { return k ? p->sht_side[k] : st; }
};
static int sht_fork(xfb_plan* p, cudaStream_t st, int n_chunks, ShtFork* f) {
    if (sht_streams_init(p)) return 1;
    f->p = p; f->st = st;
    f->ns = std::max(1, std::min(std::min(p->sht_streams, 4), std::min(n_chunks, p->max_batch / p->sht_chunk)));
    XFB_CUDA(cudaEventRecord(p->sht_fork, st));
    for (int k = 1; k < f->ns; ++k) XFB_CUDA(cudaStreamWaitEvent(p->sht_side[k], p->sht_fork, 0));
    return 0;
}
static int sht_join(ShtFork* f) {
    for (int k = 1; k < f->ns; ++k) {
        XFB_CUDA(cudaEventRecord(f->p->sht_join[k], f->p->sht_side[k]));
        XFB_CUDA(cudaStreamWaitEvent(f->st, f->p->sht_join[k], 0));
    }
    return 0;
}
static int sht_forward_chunked(xfb_plan* p, SlotView in, int spr, double2* c_out, int S, cudaStream_t st, const double2* sub, int half, int square) {
    const int cs = sht_chunk_shells(p), n_chunks = cdiv(S, cs);
    const size_t a_slot = (size_t)cs * p->M2 * p->n_theta;
    const long long shell = (long long)p->n_theta * p->n_phi;
    ShtFork f;
    prof_begin(p, PG_SHT, st);
    if (sht_fork(p, st, n_chunks, &f)) return 1;
    for (int ci = 0, s0 = 0; s0 < S; ++ci, s0 += cs) {
        const int k = ci % f.ns, Sc = std::min(cs, S - s0), b0 = s0 / spr;
        cudaStream_t sk = f.stream(k);
        SlotView v = in;
        v.base += (long long)b0 * in.run_stride;
        if (v.slot) v.slot += b0;
        double2* a = p->A0 + (size_t)k * a_slot;
        if (launch_fft(true, p->n_phi, v, spr, sub ? sub + (long long)s0 * shell : nullptr, a, p->tw, Sc, p->n_theta, p->L, sk, half | (square ? 2 : 0))) return 1;
        if (launch_legendre3(p, true, a, c_out + s0, Sc, S, half, leg3_grid_x(p, Sc, half, f.ns), sk)) return 1;
        p->launches += 2;
    }
    if (sht_join(&f)) return 1;
    prof_end(p, st);
    return 0;
}
static int sht_inverse_chunked(xfb_plan* p, const double2* c_in, double2* grid_out, int S, cudaStream_t st, int herm, const double2* mod_rho_hat,
                               SlotView mod_out, int spr) {
    const int cs = sht_chunk_shells(p), n_chunks = cdiv(S, cs);
    const size_t a_slot = (size_t)cs * p->M2 * p->n_theta;
    const long long shell = (long long)p->n_theta * p->n_phi;
    ShtFork f;
    prof_begin(p, PG_SHT, st);
    if (sht_fork(p, st, n_chunks, &f)) return 1;
    for (int ci = 0, s0 = 0; s0 < S; ++ci, s0 += cs) {
        const int k = ci % f.ns, Sc = std::min(cs, S - s0), b0 = s0 / spr;
        cudaStream_t sk = f.stream(k);
        double2* a = p->A0 + (size_t)k * a_slot;
        if (launch_legendre3(p, false, a, const_cast<double2*>(c_in) + s0, Sc, S, herm, leg3_grid_x(p, Sc, herm, f.ns), sk)) return 1;
        SlotView mo = mod_out;
        if (mod_rho_hat) { mo.base += (long long)b0 * mod_out.run_stride; if (mo.slot) mo.slot += b0; }
        if (launch_fft(false, p->n_phi, flat_view(a, 0), spr, nullptr, grid_out + (long long)s0 * shell, p->tw, Sc, p->n_theta, p->L, sk, herm,
                       mod_rho_hat ? mod_rho_hat + (long long)s0 * shell : nullptr, mo)) return 1;
        p->launches += 2;
    }
    if (sht_join(&f)) return 1;
    prof_end(p, st);
    return 0;
}

// real_only: the input field is real.  2-D: rfft semantics.  3-D (v2 Legendre only): only the m >= 0 half of the spectrum
// is produced (c_{l,-m} = (-1)^m conj c_{l,m} is redundant) -- the caller must consume m >= 0 only.  Returns the mode used
// in *half_used.
static int sht_forward_i(xfb_plan* p, SlotView in, int shells_per_run, double2* c_out, int S, cudaStream_t st,
                         const double2* sub = nullptr, int real_only = 0, int* half_used = nullptr, int square = 0) {
    if (half_used) *half_used = 0;
    if (p->dims == 2) {   // circular harmonic transform: fft(x)/n_phi  (mathLibrary.py:469-475,484-490)
        if (!sub && S <= p->max_batch * p->n_r && p->dft_fold_on && S % shells_per_run == 0)
            return dft2d_forward_folded_i(p, in, shells_per_run, c_out, S, st);
        if (!sub && S <= p->max_batch * p->n_r) {
            // dense DMMA path: rows (gathered out of their slots if needed) x [cos | sin] -> row-major coefficients -> [N][S]
            // (real_only needs nothing here: the callers pass fields whose imaginary part is exactly zero)
            const double2* rows = in.base;
            const bool flat = !in.slot && (in.run_stride == (long long)shells_per_run * p->n_phi || S <= shells_per_run);
            if (!flat) {
                const int nb_ = cdiv(S, shells_per_run);
                XFB_LAUNCH(p, PG_MISC, st, gather_slot_kernel<<<dim3(ew_blocks(p->G), nb_), 256, 0, st>>>(in, p->T2a, p->G));
                rows = p->T2a;
            }
            if (dft2d_gemm_i(p, rows, p->T2b, S, 0, 1.0 / p->n_phi, st)) return 1;
            return transpose_i(p, p->T2b, c_out, S, p->n_phi, st);
        }
        XFB_LAUNCH(p, PG_FFT, st,
                   dft2d_forward_kernel<<<cdiv(S, DFT_ROWS), DFT_THREADS, dft_smem(p->n_phi), st>>>(in, shells_per_run, sub, c_out, S, p->n_phi,
                                                                                                   1.0 / p->n_phi, real_only));
        return 0;
    }
    const int half = (real_only && p->leg2 && p->half_spectrum) ? 1 : 0;
    if (half_used) *half_used = half;
    if (sht_chunkable(p, S, shells_per_run)) return sht_forward_chunked(p, in, shells_per_run, c_out, S, st, sub, half, square);
    XFB_LAUNCH(p, PG_FFT, st, if (launch_fft(true, p->n_phi, in, shells_per_run, sub, p->A0, p->tw, S, p->n_theta, p->L, st, half | (square ? 2 : 0))) return 1);
    if (p->leg2) {     // small configuration: table-resident, cp.async double-buffered kernel
        XFB_LAUNCH(p, PG_LEGENDRE, st, if (launch_legendre3(p, true, p->A0, c_out, S, S, half, leg3_grid_x(p, S, half, 1), st)) return 1);
        return 0;
    }
    dim3 g(cdiv(S, 16), p->L + 1);
    XFB_LAUNCH(p, PG_LEGENDRE, st,
               legendre_forward_kernel<<<g, LEG_THREADS, legendre_fwd_smem(p->n_theta), st>>>(p->A0, c_out, p->FE, p->FO, S, p->L, p->n_theta, p->NP));
    return 0;
}
static int sht_inverse_i(xfb_plan* p, const double2* c_in, double2* grid_out, int S, cudaStream_t st, int herm = 0,
                         const double2* mod_rho_hat = nullptr, SlotView mod_out = SlotView{}, int shells_per_run = 1) {
    if (p->dims == 2) {   // ifft(c * n_phi) / irfft(c * n_phi, n_phi)  (mathLibrary.py:478-482,492-496)
        if (S <= p->max_batch * p->n_r && p->dft_fold_on) return dft2d_inverse_folded_i(p, c_in, grid_out, S, herm, st);
        if (S <= p->max_batch * p->n_r) {
            if (transpose_i(p, c_in, p->T2a, p->n_phi, S, st)) return 1;                 // [N][S] -> rows [S][N]
            if (herm) {
                const long long n = (long long)S * (p->n_phi / 2 + 1);
                XFB_LAUNCH(p, PG_MISC, st, hermitian_complete_rows_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, st>>>(p->T2a, S, p->n_phi));
            }
            return dft2d_gemm_i(p, p->T2a, grid_out, S, 1, 1.0, st);
        }
        XFB_LAUNCH(p, PG_FFT, st,
                   dft2d_inverse_kernel<<<cdiv(S, DFT_ROWS), DFT_THREADS, dft_smem(p->n_phi), st>>>(c_in, grid_out, S, p->n_phi, herm));
        return 0;
    }
    // herm (3-D): the coefficients belong to a real field and only m >= 0 is valid in c_in
    if (sht_chunkable(p, S, shells_per_run)) return sht_inverse_chunked(p, c_in, grid_out, S, st, herm, mod_rho_hat, mod_out, shells_per_run);
    if (p->leg2) {
        XFB_LAUNCH(p, PG_LEGENDRE, st, if (launch_legendre3(p, false, p->A0, const_cast<double2*>(c_in), S, S, herm, leg3_grid_x(p, S, herm, 1), st)) return 1);
    } else {
    dim3 g(cdiv(S, herm ? 32 : 16), p->L + 1);
    XFB_LAUNCH(p, PG_LEGENDRE, st,
               legendre_inverse_kernel<<<g, LEG_THREADS, legendre_inv_smem(p->n_theta, p->NP), st>>>(c_in, p->A0, p->IE, p->IO, S, p->L, p->n_theta, p->NP,
                                                                                                     herm));
    }
    XFB_LAUNCH(p, PG_FFT, st,
               if (launch_fft(false, p->n_phi, flat_view(p->A0, 0), shells_per_run, nullptr, grid_out, p->tw, S, p->n_theta, p->L, st, herm, mod_rho_hat,
                              mod_out)) return 1);
    return 0;
}
// chunk schedule of the host pipeline (xfb_mtip_step_host): chunks of host_chunk runs for large batches, nb/8 (>= 4: below that the
// kernels are latency bound) for small ones, so that a shard of 16 or 32 runs (the per-GPU share of 128 runs on 8 / 4 GPUs) is
// pipelined as well; full chunks in the middle, tapered at both ends (1/4, 1/2) so that the exposed first copy-in and last
// copy-out of the three-stage pipeline (copy-in | iterate | copy-out) are short
static std::vector<int> host_chunk_sizes(const xfb_plan* p, int nb, int ft_stab) {
    int chunk = p->host_chunk;
    while (chunk > 4 && chunk * 8 > nb) chunk /= 2;
    const bool pipelined = (!ft_stab || p->fused_ft_stab) && nb > chunk;   // W2 is free as H2D staging then
    std::vector<int> sizes;
    if (!pipelined) { sizes.push_back(nb); return sizes; }
    const int cs = chunk;
    std::vector<int> head, tail;
    int left = nb;
    for (int f = 4; f >= 2 && left > 4 * cs; f /= 2) {
        const int h = std::max(1, cs / f);
        head.push_back(h); tail.insert(tail.begin(), h); left -= 2 * h;
    }
    sizes = head;
    while (left > 0) { const int n = std::min(cs, left); sizes.push_back(n); left -= n; }
    sizes.insert(sizes.end(), tail.begin(), tail.end());
    return sizes;
}

// tile list of the Hankel GEMM for a batch of nb runs (cached per batch size; the first use of a size allocates, uploads and
// synchronises -- xfb_mtip_init does that ahead for the batch and for the chunk sizes of the host pipeline)
static int hankel_tiles_i(xfb_plan* p, int nb, cudaStream_t st) {
    if (p->hk_cache.find(nb) != p->hk_cache.end()) return 0;
    std::vector<HankelTile> tiles;
    if (p->dims == 2) {
        // one weight matrix per DFT index j; order m = j (j <= M) or j - N: (-i)^m = (-i)^(m mod 4)
        for (int j = 0; j < p->n_phi; ++j) {
            const int m = (j <= p->L) ? j : j - p->n_phi;
            const int r0 = j * nb, r1 = (j + 1) * nb;
            for (int r = r0; r < r1; r += HK_BM) tiles.push_back(HankelTile{j, r, r1, ((m % 4) + 4) % 4});
        }
    } else {
        for (int l = p->L; l >= 0; --l) {   // largest orders first
            const int r0 = l * l * nb, r1 = (l + 1) * (l + 1) * nb;
            for (int r = r0; r < r1; r += HK_BM) tiles.push_back(HankelTile{l, r, r1, l & 3});
        }
    }
    HankelTile* dev = nullptr;
    XFB_CUDA(cudaMalloc((void**)&dev, tiles.size() * sizeof(HankelTile)));
    XFB_CUDA(cudaMemcpyAsync(dev, tiles.data(), tiles.size() * sizeof(HankelTile), cudaMemcpyHostToDevice, st));
    XFB_CUDA(cudaStreamSynchronize(st));
    p->hk_cache.emplace(nb, std::make_pair(dev, (int)tiles.size()));
    return 0;
}
static int hankel_i(xfb_plan* p, int dir, const double2* c_in, double2* c_out, int nb, cudaStream_t st) {
    if (hankel_tiles_i(p, nb, st)) return 1;
    auto hit = p->hk_cache.find(nb);
    p->hk_tiles = hit->second.first; p->hk_tiles_n = hit->second.second;
    dim3 g(p->hk_tiles_n, cdiv(p->n_r, HK_BN));
    if (p->hankel_tma && p->n_r % HK_BN == 0 && p->hankel_n_sum % HK_BK == 0) {
        // full K chunks and column tiles: operand tiles fed by the TMA engine (bulk copies + mbarrier per stage)
        static XfbPerDeviceOnce attr3_once;
        if (xfb_first_on_device(attr3_once)) XFB_CUDA(cudaFuncSetAttribute(hankel3_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hankel3_smem()));
        XFB_LAUNCH(p, PG_HANKEL, st,
                   hankel3_tma_kernel<<<g, 256, hankel3_smem(), st>>>(c_in, c_out, p->hankel_w, p->hk_tiles, p->n_r, p->hankel_n_sum, p->hankel_skip,
                                                                     dir == 0 ? p->hk_fwd_scale : p->hk_inv_scale, dir, p->n_r, 0));
        return 0;
    }
    if (p->n_r % 2 == 0) {
        static XfbPerDeviceOnce attr_once;
        if (xfb_first_on_device(attr_once)) XFB_CUDA(cudaFuncSetAttribute(hankel2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hankel2_smem()));
        XFB_LAUNCH(p, PG_HANKEL, st,
                   hankel2_kernel<<<g, 256, hankel2_smem(), st>>>(c_in, c_out, p->hankel_w, p->hk_tiles, p->n_r, p->hankel_n_sum, p->hankel_skip,
                                                                 dir == 0 ? p->hk_fwd_scale : p->hk_inv_scale, dir, p->n_r, 0));
        return 0;
    }
    XFB_LAUNCH(p, PG_HANKEL, st,
               hankel_kernel<<<g, 256, 0, st>>>(c_in, c_out, p->hankel_w, p->hk_tiles, p->n_r, p->hankel_n_sum, p->hankel_skip,
                                                dir == 0 ? p->hk_fwd_scale : p->hk_inv_scale, dir));
    return 0;
}
static int ft_i(xfb_plan* p, int dir, SlotView in, double2* out, int nb, cudaStream_t st, const double2* sub = nullptr) {
    const int S = nb * p->n_r;
    if (sht_forward_i(p, in, p->n_r, p->C0, S, st, sub)) return 1;
    if (hankel_i(p, dir, p->C0, p->C1, nb, st)) return 1;
    return sht_inverse_i(p, p->C1, out, S, st);
}
// shell 0 of IFT(f) from the reciprocal coefficients c = SHT(f) (internal layout): rt0[nb][n_theta][n_phi]
static int ift_shell0_i(xfb_plan* p, const double2* c, int nb, cudaStream_t st) {
    const int rows = p->NLM * nb;
    XFB_LAUNCH(p, PG_MISC, st,
               hankel_row0_kernel<<<cdiv(rows, 8), 256, 0, st>>>(c, p->C0s, p->hankel_w, rows, nb, p->n_r, p->hankel_n_sum, p->hankel_skip,
                                                                  p->hk_inv_scale, 1));
    dim3 g(cdiv(nb, 16), p->L + 1);
    XFB_LAUNCH(p, PG_MISC, st,
               legendre_inverse_kernel<<<g, LEG_THREADS, legendre_inv_smem(p->n_theta, p->NP), st>>>(p->C0s, p->A0s, p->IE, p->IO, nb, p->L, p->n_theta, p->NP, 0));
    XFB_LAUNCH(p, PG_MISC, st,
               if (launch_fft(false, p->n_phi, flat_view(p->A0s, 0), 1, nullptr, p->rt0, p->tw, nb, p->n_theta, p->L, st)) return 1);
    return 0;
}

// Problem descriptors of the two grouped GEMMs, built ONCE for the plan capacity: the list is run-major, so the
// descriptors (and tile ids) of a smaller batch are a prefix of it.
static int build_gemm_groups(xfb_plan* p, int nb_req, cudaStream_t st) {
    if (p->gemm_nb > 0) {
        p->gemmM_tiles = nb_req * p->gemmM_tiles_run; p->gemmT_tiles = nb_req * p->gemmT_tiles_run; p->gemmY_tiles = nb_req * p->gemmY_tiles_run;
        return 0;
    }
    const int nb = p->max_batch;
    std::vector<GemmProblem> pm, pt, py; std::vector<int> tpm, tpt, tpy;
    for (int b = 0; b < nb; ++b) {
        for (const ProcOrder& o : p->orders) {
            GemmProblem a{};
            a.A = p->pd_dev + o.pd_off; a.a_rs = p->n_r; a.a_cs = 1;
            a.B = p->xt + (size_t)b * p->xt_run + o.xt_off; a.b_rs = 1; a.b_cs = p->n_r;
            a.C = p->g + (size_t)b * p->g_run + o.g_off; a.c_rs = jacobi_stride(o.n_c); a.c_cs = 1;
            a.M = o.n_cols; a.N = o.n_c; a.K = p->n_r; a.alpha = 1.0;
            a.tile0 = (int)tpm.size(); a.tiles_n = cdiv(a.N, GG_BN);
            for (int t = 0; t < cdiv(a.M, GG_BM) * a.tiles_n; ++t) tpm.push_back((int)pm.size());
            pm.push_back(a);
            GemmProblem c{};
            c.A = p->gn + (size_t)b * p->g_run + o.g_off; c.a_rs = 1; c.a_cs = jacobi_stride(o.n_c);
            c.B = p->vw + (size_t)b * p->vw_run + o.vw_off; c.b_rs = jacobi_wstride(p->n_r); c.b_cs = 1;
            c.C = p->tt + (size_t)b * p->xt_run + o.xt_off; c.c_rs = p->n_r; c.c_cs = 1;
            c.M = o.n_c; c.N = p->n_r; c.K = o.n_cols; c.alpha = 1.0;
            c.tile0 = (int)tpt.size(); c.tiles_n = cdiv(c.N, GG_BN);
            for (int t = 0; t < cdiv(c.M, GG_BM) * c.tiles_n; ++t) tpt.push_back((int)pt.size());
            pt.push_back(c);
            GemmProblem y{};      // vw = pp V_l^T : [n_cols x n_cols] . [n_cols x N_r]
            y.A = p->pp + (size_t)b * p->g_run + o.g_off; y.a_rs = jacobi_stride(o.n_c); y.a_cs = 1;
            y.B = p->vt_dev + o.pd_off; y.b_rs = p->n_r; y.b_cs = 1;
            y.C = p->vw + (size_t)b * p->vw_run + o.vw_off; y.c_rs = jacobi_wstride(p->n_r); y.c_cs = 1;
            y.M = o.n_cols; y.N = p->n_r; y.K = o.n_cols; y.alpha = 1.0;
            y.tile0 = (int)tpy.size(); y.tiles_n = cdiv(y.N, GG_BN);
            for (int t = 0; t < cdiv(y.M, GG_BM) * y.tiles_n; ++t) tpy.push_back((int)py.size());
            py.push_back(y);
        }
    }
    XFB_CUDA(cudaMemcpyAsync(p->gemmM_dev, pm.data(), pm.size() * sizeof(GemmProblem), cudaMemcpyHostToDevice, st));
    XFB_CUDA(cudaMemcpyAsync(p->gemmT_dev, pt.data(), pt.size() * sizeof(GemmProblem), cudaMemcpyHostToDevice, st));
    XFB_CUDA(cudaMemcpyAsync(p->gemmM_tp, tpm.data(), tpm.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    XFB_CUDA(cudaMemcpyAsync(p->gemmT_tp, tpt.data(), tpt.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    XFB_CUDA(cudaMemcpyAsync(p->gemmY_dev, py.data(), py.size() * sizeof(GemmProblem), cudaMemcpyHostToDevice, st));
    XFB_CUDA(cudaMemcpyAsync(p->gemmY_tp, tpy.data(), tpy.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    XFB_CUDA(cudaStreamSynchronize(st));
    p->gemmM_tiles_run = (int)tpm.size() / nb; p->gemmT_tiles_run = (int)tpt.size() / nb; p->gemmY_tiles_run = (int)tpy.size() / nb; p->gemm_nb = nb;
    p->gemmM_tiles = nb_req * p->gemmM_tiles_run; p->gemmT_tiles = nb_req * p->gemmT_tiles_run; p->gemmY_tiles = nb_req * p->gemmY_tiles_run;
    return 0;
}

// one-sided Jacobi over all (order, run) problems of a batch; kernel variant by problem size (procrustes.cuh).
// L <= 63 class: the orders with 2l+1 > 64 (one problem per SM, 512 threads) are launched on the caller's stream, the
// orders with 2l+1 <= 64 (two problems per SM, 256 threads each) on a forked stream: their CTAs fill the SMs as the
// large problems drain, and both launches pull their problems from their own work queue.
#define JAC_SMALL_SMEM (111 * 1024)
static int launch_jacobi(xfb_plan* p, const double* g, double* gn, double* pp, double* sigma, int nb, int* sweeps, cudaStream_t st) {
    const int na = (int)p->orders.size();
    const int smem_doubles = (int)(p->jacobi_smem / 8);
    if (!p->jac_counter) { if (dev_alloc(p, &p->jac_counter, 4)) return 1; }
    int* counter = p->jac_counter + 2 * p->ctx;
    XFB_CUDA(cudaMemsetAsync(counter, 0, 2 * sizeof(int), st));
    const long long sig_run = (long long)na * p->sig_ld;
    if (p->stagger_pending) {                 // dual mode: the other half starts its first iteration when this one enters the Jacobi
        XFB_CUDA(cudaEventRecord(p->stagger_pending, st));
        p->stagger_pending = nullptr;
    }
    if (p->jacobi_big) {
        procrustes_jacobi_kernel<16, 256, false, 1, 32><<<std::min(na * nb, p->n_sm), 256, p->jacobi_smem, st>>>(
            g, gn, pp, sigma, p->orders_dev, na, nb, p->sig_ld, p->g_run, sig_run, p->sv_cutoff, 1e-15, p->max_sweeps, sweeps, smem_doubles, counter, 0, na);
        XFB_CUDA(cudaGetLastError());
        return 0;
    }
    const int n_big = p->jac_split ? p->jac_n_big : na, n_small = na - n_big;
    // SMs given to the two launches: all of them, or the dual-mode shares (the rest is left to the other half's transforms)
    const int sms_big = p->dual_active ? std::max(1, std::min(p->n_sm, p->dual_big_sms)) : p->n_sm;
    const int sms_small = p->dual_active ? std::max(1, std::min(p->n_sm, p->dual_small_sms)) : p->n_sm;
    cudaStream_t s_small = st;
    if (n_big > 0 && n_small > 0) {          // fork: the small-order launch goes to a side stream
        if (sht_streams_init(p)) return 1;
        if (!p->jac_fork[p->ctx]) {
            XFB_CUDA(cudaEventCreateWithFlags(&p->jac_fork[p->ctx], cudaEventDisableTiming));
            XFB_CUDA(cudaEventCreateWithFlags(&p->jac_join[p->ctx], cudaEventDisableTiming));
        }
        s_small = p->sht_side[1 + p->ctx];
        XFB_CUDA(cudaEventRecord(p->jac_fork[p->ctx], st));
        XFB_CUDA(cudaStreamWaitEvent(s_small, p->jac_fork[p->ctx], 0));
    }
    if (n_big > 0)
        procrustes_jacobi_kernel<8, 512, true, 1, 16><<<std::min(n_big * nb, sms_big), 512, p->jacobi_smem, st>>>(
            g, gn, pp, sigma, p->orders_dev, n_big, nb, p->sig_ld, p->g_run, sig_run, p->sv_cutoff, 1e-15, p->max_sweeps, sweeps, smem_doubles, counter, 0, na);
    if (n_small > 0)
        procrustes_jacobi_kernel<4, 256, true, 2, 8><<<std::min(n_small * nb, 2 * sms_small), 256, JAC_SMALL_SMEM, s_small>>>(
            g, gn, pp, sigma, p->orders_dev, n_small, nb, p->sig_ld, p->g_run, sig_run, p->sv_cutoff, 1e-15, p->max_sweeps, sweeps, JAC_SMALL_SMEM / 8,
            counter + 1, n_big, na);
    XFB_CUDA(cudaGetLastError());
    if (n_big > 0 && n_small > 0) {          // join
        XFB_CUDA(cudaEventRecord(p->jac_join[p->ctx], s_small));
        XFB_CUDA(cudaStreamWaitEvent(st, p->jac_join[p->ctx], 0));
    }
    return 0;
}

// I coefficients (internal layout) -> projected coefficients
static int project_i(xfb_plan* p, const double2* c_in, double2* c_out, int nb, cudaStream_t st, int half = 0) {
    if (!p->has_proj) XFB_FAIL("projection constants not set (xfb_plan_set_projection)");
    const int S = nb * p->n_r;
    if (p->dims == 2) {
        p->proj_calls++; p->last_proj_nb = nb;
        XFB_LAUNCH(p, PG_PROC_PACK, st,
                   project2d_kernel<<<nb, 256, 0, st>>>(c_in, c_out, p->v2d, p->radial_mask_dev, p->q_pts, p->unk2d, p->n_orders2d, p->L, p->n_phi,
                                                        p->n_r, S, p->inv_sqrt_np, p->so_order2d));
        return 0;
    }
    const int na = (int)p->orders.size();
    // row (l=0,m=0) of the coefficient array = I_00(q) of every run: kept for xfb_get_unknowns
    XFB_CUDA(cudaMemcpyAsync(p->i00, c_in, (size_t)S * sizeof(double2), cudaMemcpyDeviceToDevice, st));
    p->proj_calls++; p->last_proj_nb = std::max(p->run_base + nb, p->run_base ? p->last_proj_nb : 0);
    if (na > 0) {
        if (build_gemm_groups(p, nb, st)) return 1;
        XFB_LAUNCH(p, PG_PROC_PACK, st,
                   procrustes_pack_kernel<<<dim3(na, nb), 256, 0, st>>>(c_in, p->xt, p->orders_dev, p->n_r, S, p->xt_run));
        XFB_LAUNCH(p, PG_PROC_GEMM, st, grouped_gemm_kernel<<<p->gemmM_tiles, 128, 0, st>>>(p->gemmM_dev, p->gemmM_tp, p->run_base * p->gemmM_tiles_run));
        XFB_LAUNCH(p, PG_PROC_JACOBI, st,
                   if (launch_jacobi(p, p->g, p->gn, p->pp, p->sigma, nb, p->sweeps_dev, st)) return 1);
        XFB_LAUNCH(p, PG_PROC_GEMM, st, grouped_gemm_kernel<<<p->gemmY_tiles, 128, 0, st>>>(p->gemmY_dev, p->gemmY_tp, p->run_base * p->gemmY_tiles_run));
        XFB_LAUNCH(p, PG_PROC_GEMM, st, grouped_gemm_kernel<<<p->gemmT_tiles, 128, 0, st>>>(p->gemmT_dev, p->gemmT_tp, p->run_base * p->gemmT_tiles_run));
    }
    XFB_LAUNCH(p, PG_PROC_PACK, st,
               procrustes_unpack_kernel<<<dim3(p->L + 1, nb), 256, 0, st>>>(c_in, c_out, p->tt, p->orders_dev, p->kind_dev, p->act_index_dev,
                                                                            p->radial_mask_dev, p->v0_dev, p->inv_sqrt_np, p->L, p->n_r, S,
                                                                            p->xt_run, half));
    return 0;
}

// ---- degree-2 invariants of a batch of coefficient arrays (internal layout, real field) ---------------------------------
static int ensure_deg2_alloc(xfb_plan* p, cudaStream_t st) {
    if (p->d2_x) return 0;
    if (p->dims != 3) XFB_FAIL("degree-2 invariants on the device are built for the 3-D plan");
    const size_t B = p->max_batch, n_r = p->n_r, L1 = p->L + 1;
    if (dev_alloc(p, &p->d2_x, B * (size_t)p->NLM * n_r)) return 1;
    if (dev_alloc(p, &p->d2_b, B * L1 * n_r * n_r)) return 1;
    std::vector<GemmProblem> pr; std::vector<int> tp;
    for (size_t b = 0; b < B; ++b)
        for (int l = 0; l <= p->L; ++l) {
            GemmProblem g{};       // B_l [N_r x N_r] = X_l^T-view [N_r x (2l+1)] . X_l [(2l+1) x N_r]
            const double* X = p->d2_x + b * (size_t)p->NLM * n_r + (size_t)l * l * n_r;
            g.A = X; g.a_rs = 1; g.a_cs = (long long)n_r;
            g.B = X; g.b_rs = (long long)n_r; g.b_cs = 1;
            g.C = p->d2_b + (b * L1 + l) * n_r * n_r; g.c_rs = (long long)n_r; g.c_cs = 1;
            g.M = (int)n_r; g.N = (int)n_r; g.K = 2 * l + 1; g.alpha = 1.0;
            g.tile0 = (int)tp.size(); g.tiles_n = cdiv(g.N, GG_BN);
            for (int t = 0; t < cdiv(g.M, GG_BM) * g.tiles_n; ++t) tp.push_back((int)pr.size());
            pr.push_back(g);
        }
    if (dev_alloc(p, &p->d2_gemm, pr.size())) return 1;
    if (dev_alloc(p, &p->d2_tp, tp.size())) return 1;
    XFB_CUDA(cudaMemcpyAsync(p->d2_gemm, pr.data(), pr.size() * sizeof(GemmProblem), cudaMemcpyHostToDevice, st));
    XFB_CUDA(cudaMemcpyAsync(p->d2_tp, tp.data(), tp.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    XFB_CUDA(cudaStreamSynchronize(st));
    p->d2_tiles_run = (int)(tp.size() / B);
    return 0;
}
// c: coefficients [(L+1)^2][S] of a real field (m >= 0 valid) -> p->d2_b [nb][L+1][N_r][N_r] real
static int deg2_invariants_i(xfb_plan* p, const double2* c, int nb, cudaStream_t st) {
    if (ensure_deg2_alloc(p, st)) return 1;
    const int S = nb * p->n_r;
    XFB_LAUNCH(p, PG_PROC_PACK, st, pack_real_all_kernel<<<dim3(p->L + 1, nb), 256, 0, st>>>(c, p->d2_x, p->n_r, S, (long long)p->NLM * p->n_r));
    XFB_LAUNCH(p, PG_PROC_GEMM, st, grouped_gemm_kernel<<<nb * p->d2_tiles_run, 128, 0, st>>>(p->d2_gemm, p->d2_tp, p->run_base * p->d2_tiles_run));
    return 0;
}
static int deg2_diff_i(xfb_plan* p, const double2* c, int nb, double* err_out, long long err_run_stride, cudaStream_t st) {
    if (!p->d2_ref) XFB_FAIL("deg2_invariant_l2_diff: reference invariants not set (xfb_plan_set_deg2_reference)");
    if (deg2_invariants_i(p, c, nb, st)) return 1;
    XFB_LAUNCH(p, PG_MISC, st,
               deg2_diff_kernel<<<dim3(p->L + 1, nb), 256, 0, st>>>(p->d2_b, p->d2_ref, p->d2_norm, p->radial_mask_dev, p->n_r, p->L + 1,
                                                                   (long long)(p->L + 1) * p->n_r * p->n_r, err_out, err_run_stride));
    return 0;
}

static int transpose_i(xfb_plan* p, const double2* in, double2* out, int rows, int cols, cudaStream_t st) {
    dim3 g(cdiv(cols, 32), cdiv(rows, 32)), b(32, 8);
    XFB_LAUNCH(p, PG_MISC, st, transpose_c128_kernel<<<g, b, 0, st>>>(in, out, rows, cols));
    return 0;
}

static int ensure_loop_alloc(xfb_plan* p) {
    if (p->loop_alloc) return 0;
    const size_t B = p->max_batch;
    if (dev_alloc(p, &p->rho_pool, 3 * B * p->G)) return 1;
    if (dev_alloc(p, &p->rh_pool, 3 * B * p->G)) return 1;
    if (dev_alloc(p, &p->mask_pool, 3 * B * p->G)) return 1;
    if (dev_alloc(p, &p->ls_ints, 14 * B)) return 1;
    const int hist_cap = 1 << 14;
    if (dev_alloc(p, &p->ls_dbl, 2 * B + B * hist_cap)) return 1;
    int* q = p->ls_ints;
    p->ls.rho_cur = q; p->ls.rho_best = q + B; p->ls.rho_next = q + 2 * B;
    p->ls.rh_cur = q + 3 * B; p->ls.rh_best = q + 4 * B; p->ls.rh_next = q + 5 * B;
    p->ls.mask_cur = q + 6 * B; p->ls.mask_best = q + 7 * B; p->ls.mask_next = q + 8 * B;
    p->ls.enforce_cur = q + 9 * B; p->ls.enforce_best = q + 10 * B; p->ls.enforce_rep = q + 11 * B; p->ls.best_iter = q + 12 * B; p->ls.nonfinite = q + 13 * B;
    p->ls.best_err = p->ls_dbl; p->ls.last_err = p->ls_dbl + B; p->ls.hist = p->ls_dbl + 2 * B; p->ls.hist_cap = hist_cap;
    p->loop_alloc = true;
    return 0;
}
static int ensure_reduce_alloc(xfb_plan* p) {
    if (p->partial) return 0;
    p->red_blocks = 148 * 2;
    if (dev_alloc(p, &p->partial, (size_t)p->max_batch * p->red_blocks * 2)) return 1;
    if (dev_alloc(p, &p->err, (size_t)p->max_batch * 2)) return 1;
    if (dev_alloc(p, &p->mm, (size_t)p->max_batch * 2)) return 1;
    return 0;
}

static int real_update_i(xfb_plan* p, int method, double beta, const double2* rho_ift, const double2* rho_rt, SlotView prev, SlotView next,
                         const uint8_t* support, const int* support_slot, long long support_slot_stride, const int* enforce, double* err_out,
                         int nb, cudaStream_t st, const double2* rt0 = nullptr) {
    if (!p->has_real) XFB_FAIL("real projection options not set (xfb_plan_set_real)");
    if (ensure_reduce_alloc(p)) return 1;
    // blocks per run: RU_WAVES x 148 CTAs per launch (8 .. 128 measured at 128 runs: 1.27, 1.23, 1.18, 1.20, 1.24 ms)
    int bpr = std::max(1, std::min(p->red_blocks, (148 * RU_WAVES + nb - 1) / nb));
    bpr = (int)std::min<long long>(bpr, cdiv64(p->G, RU_THREADS));
    const bool has_avg = p->rd.avg_shells > 0;
    if (has_avg)      // average_center: angular means of the first shells at that position of the chain (fxs_Projections.py:96-110)
        XFB_LAUNCH(p, PG_REAL_UPDATE, st,
                   average_center_kernel<<<dim3(p->rd.avg_shells, nb), 256, 0, st>>>(rho_ift, rho_rt, prev, support, support_slot, support_slot_stride, enforce,
                                                                                   p->init_support_dev, p->rd, p->n_theta, p->n_phi, p->G, rt0, p->avg_mean));
    XFB_LAUNCH(p, PG_REAL_UPDATE, st,
               real_update_kernel<<<dim3(bpr, nb), RU_THREADS, 0, st>>>(rho_ift, rho_rt, prev, next, support, support_slot, support_slot_stride,
                                                                        enforce, p->init_support_dev, p->int_wt, p->rd, method, beta,
                                                                        p->n_theta, p->n_phi, p->wt_div, p->G, p->partial, rt0, has_avg ? p->avg_mean : nullptr, p->ip_active));
    XFB_LAUNCH(p, PG_MISC, st, reduce_pairs_kernel<<<nb, 32, 0, st>>>(p->partial, bpr, err_out));
    return 0;
}

static int shrinkwrap_i(xfb_plan* p, SlotView rho, double sigma, double threshold, uint8_t* out, const int* out_slot,
                        long long out_slot_stride, int nb, cudaStream_t st) {
    if (ensure_reduce_alloc(p)) return 1;
    const int eb = ew_blocks(p->G);
    XFB_LAUNCH(p, PG_POINTWISE, st, abs_kernel<<<dim3(eb, nb), 256, 0, st>>>(rho, p->W0, p->G));
    if (ft_i(p, 0, flat_view(p->W0, p->G), p->W1, nb, st)) return 1;
    const long long shell = (long long)p->n_theta * p->n_phi;
    XFB_LAUNCH(p, PG_POINTWISE, st,
               mul_gauss_kernel<<<dim3(nb * p->n_r, (int)std::min<long long>(cdiv64(shell, 256), 8)), 256, 0, st>>>(p->W1, p->q_pts, sigma, p->n_r, shell));
    if (ft_i(p, 1, flat_view(p->W1, p->G), p->W0, nb, st)) return 1;
    int bpr = std::max(1, std::min(p->red_blocks, (148 * RU_WAVES + nb - 1) / nb));
    XFB_LAUNCH(p, PG_POINTWISE, st, sw_minmax_kernel<<<dim3(bpr, nb), 256, 0, st>>>(p->W0, p->G, p->partial));
    XFB_LAUNCH(p, PG_MISC, st, sw_minmax_final_kernel<<<nb, 32, 0, st>>>(p->partial, bpr, p->mm));
    XFB_LAUNCH(p, PG_POINTWISE, st, sw_mask_kernel<<<dim3(eb, nb), 256, 0, st>>>(p->W0, p->mm, threshold, out, out_slot, out_slot_stride, p->G));
    return 0;
}

extern "C" {

// a plan can be re-targeted to new invariants: the projection constants and their per-run workspaces are released and rebuilt
static void projection_release(xfb_plan* p) {
    void** ptrs[] = {(void**)&p->orders_dev, (void**)&p->kind_dev, (void**)&p->act_index_dev, (void**)&p->radial_mask_dev, (void**)&p->v0_dev,
                     (void**)&p->pd_dev, (void**)&p->vt_dev, (void**)&p->xt, (void**)&p->tt, (void**)&p->g, (void**)&p->gn, (void**)&p->vw,
                     (void**)&p->sigma, (void**)&p->sweeps_dev, (void**)&p->pp, (void**)&p->gn_u, (void**)&p->pp_u, (void**)&p->sigma_u,
                     (void**)&p->i00, (void**)&p->gemmM_dev, (void**)&p->gemmT_dev, (void**)&p->gemmY_dev, (void**)&p->gemmM_tp,
                     (void**)&p->gemmT_tp, (void**)&p->gemmY_tp, (void**)&p->v2d, (void**)&p->unk2d, (void**)&p->d2_ref, (void**)&p->d2_norm};
    cudaDeviceSynchronize();
    for (void** q : ptrs) if (*q) { cudaFree(*q); *q = nullptr; }
    p->orders.clear(); p->ncols_all.clear();
    p->gemm_nb = -1; p->proj_calls = 0; p->unk_stamp = -1; p->unk_run = -1; p->last_proj_nb = 0; p->d2_metric = false;
    p->has_proj = false;
}

int xfb_plan_set_projection(xfb_plan* p, const xfb_projection_desc* d) {
    if (!p || !d) XFB_FAIL("null argument");
    graphs_invalidate(p);
    if (p->has_proj) projection_release(p);
    if (p->dims != 3) XFB_FAIL("xfb_plan_set_projection is the 3-D setter; use xfb_plan_set_projection_2d");
    const int L = p->L, n_r = p->n_r;
    std::vector<int> kind(L + 1, ORD_PASS), act(L + 1, -1);
    std::vector<double> pd, vt, v0(n_r, 0.0);
    std::vector<double> q(n_r);
    XFB_CUDA(cudaMemcpy(q.data(), p->q_pts, n_r * sizeof(double), cudaMemcpyDeviceToHost));
    struct Tmp { int l, n_cols; std::vector<double> pd, vt; };
    std::vector<Tmp> tmp;
    p->ncols_all.assign(L + 1, 0);
    for (int l = 0; l < d->n_orders && l <= L; ++l) {
        const int nc = d->n_cols[l];
        p->ncols_all[l] = nc;
        const double* V = d->v[l];
        // n_cols > N_r happens when the data come on a finer q grid (n_cols = min(N_q_data, 2l+1)) and are regridded to N_r points
        if (nc < 1 || nc > 2 * l + 1) XFB_FAIL("order %d: n_cols=%d outside 1..2l+1", l, nc);
        if (l == 0) {
            kind[0] = ORD_ZEROTH;
            for (int k = 0; k < n_r; ++k) v0[k] = V[(size_t)k * nc];
            continue;
        }
        bool zero = true;
        for (size_t i = 0; i < (size_t)n_r * nc; ++i) if (V[i] != 0.0) { zero = false; break; }
        if (zero) { kind[l] = ORD_ZERO; continue; }
        kind[l] = ORD_ACTIVE;
        Tmp t; t.l = l; t.n_cols = nc; t.pd.resize((size_t)nc * n_r); t.vt.resize((size_t)nc * n_r);
        for (int i = 0; i < nc; ++i)
            for (int k = 0; k < n_r; ++k) {
                const double v = V[(size_t)k * nc + i];
                t.vt[(size_t)i * n_r + k] = v;
                t.pd[(size_t)i * n_r + k] = v * (q[k] * q[k]);     // V^T diag(q)^2  (fxs_Projections.py:753-754)
            }
        tmp.push_back(std::move(t));
    }
    std::sort(tmp.begin(), tmp.end(), [](const Tmp& a, const Tmp& b) { return a.l > b.l; });   // largest first
    long long pd_off = 0, xt_off = 0, g_off = 0, vw_off = 0;
    size_t smem_max = 0; int max_nc = 0;
    for (size_t i = 0; i < tmp.size(); ++i) {
        ProcOrder o{};
        o.l = tmp[i].l; o.n_cols = tmp[i].n_cols; o.n_c = 2 * o.l + 1;
        o.pd_off = pd_off; o.xt_off = xt_off; o.g_off = g_off; o.vw_off = vw_off;
        pd_off += (long long)o.n_cols * n_r; xt_off += (long long)o.n_c * n_r;
        g_off += (long long)o.n_cols * jacobi_stride(o.n_c); vw_off += (long long)o.n_cols * jacobi_wstride(n_r);
        act[o.l] = (int)i;
        p->orders.push_back(o);
        pd.insert(pd.end(), tmp[i].pd.begin(), tmp[i].pd.end());
        vt.insert(vt.end(), tmp[i].vt.begin(), tmp[i].vt.end());
        max_nc = std::max(max_nc, o.n_c);
    }
    p->xt_run = xt_off; p->g_run = g_off; p->vw_run = vw_off;
    smem_max = (size_t)226 * 1024;                    // the Jacobi kernel takes the whole SM: G (and W) live in shared memory when they fit
    p->jacobi_smem = smem_max;
    if (n_r > 256 || max_nc > 256) XFB_FAIL("Procrustes kernel supports N_r <= 256 and 2l+1 <= 256 (got N_r=%d, 2l+1=%d)", n_r, max_nc);
    p->jacobi_big = (n_r > 128 || max_nc > 128);
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&p->n_sm, cudaDevAttrMultiProcessorCount, dev); }
    if (dev_upload(p, &p->kind_dev, kind.data(), kind.size())) return 1;
    if (dev_upload(p, &p->act_index_dev, act.data(), act.size())) return 1;
    if (dev_upload(p, &p->radial_mask_dev, d->radial_mask, (size_t)(L + 1) * n_r)) return 1;
    if (dev_upload(p, &p->v0_dev, v0.data(), v0.size())) return 1;
    const size_t B = p->max_batch, na = p->orders.size();
    if (na > 0) {
        if (dev_upload(p, &p->orders_dev, p->orders.data(), na)) return 1;
        if (dev_upload(p, &p->pd_dev, pd.data(), pd.size())) return 1;
        if (dev_upload(p, &p->vt_dev, vt.data(), vt.size())) return 1;
        if (dev_alloc(p, &p->xt, B * p->xt_run)) return 1;
        if (dev_alloc(p, &p->tt, B * p->xt_run)) return 1;
        if (dev_alloc(p, &p->g, B * p->g_run)) return 1;
        XFB_CUDA(cudaMemset(p->g, 0, B * p->g_run * sizeof(double)));       // the column pads stay zero: the GEMM writes only [n_cols][n_c]
        if (dev_alloc(p, &p->gn, B * p->g_run)) return 1;
        if (dev_alloc(p, &p->vw, B * p->vw_run)) return 1;
        XFB_CUDA(cudaMemset(p->vw, 0, B * p->vw_run * sizeof(double)));     // rows of dropped columns are never written: keep them finite
        p->sig_ld = n_r;
        for (const ProcOrder& o : p->orders) p->sig_ld = std::max(p->sig_ld, o.n_cols);
        if (dev_alloc(p, &p->sigma, B * na * p->sig_ld)) return 1;
        if (dev_alloc(p, &p->sweeps_dev, B * na)) return 1;
        if (dev_alloc(p, &p->pp, B * p->g_run)) return 1;
        if (dev_alloc(p, &p->gn_u, (size_t)p->g_run)) return 1;
        if (dev_alloc(p, &p->pp_u, (size_t)p->g_run)) return 1;
        if (dev_alloc(p, &p->sigma_u, na * p->sig_ld)) return 1;
        if (dev_alloc(p, &p->gemmM_dev, B * na)) return 1;
        if (dev_alloc(p, &p->gemmT_dev, B * na)) return 1;
        if (dev_alloc(p, &p->gemmY_dev, B * na)) return 1;
        size_t tiles_m = 0, tiles_t = 0, tiles_y = 0;
        for (const ProcOrder& o : p->orders) {
            tiles_m += (size_t)cdiv(o.n_cols, GG_BM) * cdiv(o.n_c, GG_BN);
            tiles_t += (size_t)cdiv(o.n_c, GG_BM) * cdiv(n_r, GG_BN);
            tiles_y += (size_t)cdiv(o.n_cols, GG_BM) * cdiv(n_r, GG_BN);
        }
        if (dev_alloc(p, &p->gemmM_tp, B * tiles_m)) return 1;
        if (dev_alloc(p, &p->gemmT_tp, B * tiles_t)) return 1;
        if (dev_alloc(p, &p->gemmY_tp, B * tiles_y)) return 1;
        if (p->jacobi_big) XFB_CUDA(cudaFuncSetAttribute(procrustes_jacobi_kernel<16, 256, false, 1, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        else {
            XFB_CUDA(cudaFuncSetAttribute(procrustes_jacobi_kernel<8, 512, true, 1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
            XFB_CUDA(cudaFuncSetAttribute(procrustes_jacobi_kernel<4, 256, true, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, JAC_SMALL_SMEM));
        }
        p->jac_n_big = 0;
        for (const ProcOrder& o : p->orders) if (o.n_c > 64) p->jac_n_big++;        // sorted largest first: a prefix
        if (const char* e = getenv("XFB_JACOBI_SPLIT")) p->jac_split = atoi(e) != 0;   // experiments
    } else {
        // dummy so the unpack kernel has valid pointers
        ProcOrder o{};
        if (dev_upload(p, &p->orders_dev, &o, 1)) return 1;
        if (dev_alloc(p, &p->tt, 1)) return 1;
    }
    if (dev_alloc(p, &p->i00, B * n_r)) return 1;
    p->inv_sqrt_np = 1.0 / d->sqrt_n_particles;
    p->sv_cutoff = d->sv_cutoff > 0 ? d->sv_cutoff : 1e-15;
    p->max_sweeps = d->max_sweeps > 0 ? d->max_sweeps : 40;
    p->has_proj = true;
    return 0;
}

int xfb_plan_set_projection_2d(xfb_plan* p, int32_t n_orders, const double* v, const uint8_t* radial_mask, double sqrt_n_particles,
                               int32_t so_order_id) {
    if (!p || !v || !radial_mask) XFB_FAIL("null argument");
    if (p->dims != 2) XFB_FAIL("xfb_plan_set_projection_2d needs a 2-D plan");
    if (p->has_proj) projection_release(p);
    if (n_orders < 1 || n_orders > p->L + 1) XFB_FAIL("n_orders=%d outside 1..%d", n_orders, p->L + 1);
    if (dev_upload(p, &p->v2d, (const double2*)v, (size_t)n_orders * p->n_r)) return 1;
    if (dev_upload(p, &p->radial_mask_dev, radial_mask, (size_t)(p->L + 1) * p->n_r)) return 1;
    if (dev_alloc(p, &p->unk2d, (size_t)p->max_batch * n_orders)) return 1;
    p->n_orders2d = n_orders; p->so_order2d = so_order_id;
    p->inv_sqrt_np = 1.0 / sqrt_n_particles;
    p->has_proj = true;
    return 0;
}

// 2-D fxs_unknowns of the last projection: out_dev [n_orders] complex (fxs_Projections.py:727-748)
int xfb_get_unknowns_2d(xfb_plan* p, int32_t run, double* out_dev, void* stream) {
    if (p->dims != 2 || !p->has_proj) XFB_FAIL("xfb_get_unknowns_2d needs a 2-D plan with projection constants");
    if (p->proj_calls == 0) XFB_FAIL("xfb_get_unknowns_2d: no invariant projection has run yet");
    if (run < 0 || run >= p->last_proj_nb) XFB_FAIL("run=%d outside the last projected batch (%d)", run, p->last_proj_nb);
    XFB_CUDA(cudaMemcpyAsync(out_dev, p->unk2d + (size_t)run * p->n_orders2d, (size_t)p->n_orders2d * sizeof(double2), cudaMemcpyDeviceToDevice,
                             (cudaStream_t)stream));
    return 0;
}

int xfb_plan_set_real(xfb_plan* p, const xfb_real_desc* d, const uint8_t* init_support_host) {
    if (!p || !d || !init_support_host) XFB_FAIL("null argument");
    if (d->n_ops < 0 || d->n_ops > 4) XFB_FAIL("n_ops out of range");
    graphs_invalidate(p);
    p->rd.n_ops = d->n_ops;
    for (int i = 0; i < 4; ++i) { p->rd.ops[i] = d->ops[i]; p->rd.considered[i] = d->hio_considered[i]; }
    p->rd.use_lo = d->use_lo; p->rd.use_hi = d->use_hi; p->rd.lo = d->lo; p->rd.hi = d->hi; p->rd.imag_limit = d->imag_limit;
    p->rd.err_inside = d->error_inside_initial_support;
    p->rd.avg_shells = 0;
    int n_avg = 0;
    for (int i = 0; i < d->n_ops; ++i) {
        if (d->ops[i] < XFB_OP_SUPPORT || d->ops[i] > XFB_OP_AVERAGE_CENTER) XFB_FAIL("real projection op %d unknown", d->ops[i]);
        n_avg += d->ops[i] == XFB_OP_AVERAGE_CENTER;
    }
    if (n_avg > 1) XFB_FAIL("average_center may appear once in the real projection chain");
    if (n_avg) {
        // density[:thresh] with thresh = int(max_radial_id): numpy clips the slice to the N_r shells (fxs_Projections.py:97,101)
        p->rd.avg_shells = std::max(0, std::min((int)d->average_center_shells, p->n_r));
        if (p->rd.avg_shells > 0 && !p->avg_mean) { if (dev_alloc(p, &p->avg_mean, (size_t)p->max_batch * p->n_r)) return 1; }
    }
    if (!p->init_support_dev) { if (dev_alloc(p, &p->init_support_dev, (size_t)p->G)) return 1; }
    XFB_CUDA(cudaMemcpy(p->init_support_dev, init_support_host, (size_t)p->G, cudaMemcpyHostToDevice));
    p->has_real = true;
    return 0;
}

// ---- operator level -------------------------------------------------------------------------
int xfb_sht_forward(xfb_plan* p, const double* grid, double* direct, int32_t n_shells, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_shells < 1 || n_shells > p->max_batch * p->n_r) XFB_FAIL("n_shells=%d outside plan capacity", n_shells);
    if (sht_forward_i(p, flat_view((const double2*)grid, 0), n_shells, p->C0, n_shells, st)) return 1;
    return transpose_i(p, p->C0, (double2*)direct, p->NLM, n_shells, st);
}
int xfb_sht_inverse(xfb_plan* p, const double* direct, double* grid, int32_t n_shells, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_shells < 1 || n_shells > p->max_batch * p->n_r) XFB_FAIL("n_shells=%d outside plan capacity", n_shells);
    if (transpose_i(p, (const double2*)direct, p->C0, n_shells, p->NLM, st)) return 1;
    return sht_inverse_i(p, p->C0, (double2*)grid, n_shells, st);
}
int xfb_hankel_apply(xfb_plan* p, int32_t dir, const double* in, double* out, int32_t nb, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (nb < 1 || nb > p->max_batch) XFB_FAIL("n_batch=%d outside plan capacity", nb);
    const int S = nb * p->n_r;
    if (transpose_i(p, (const double2*)in, p->C0, S, p->NLM, st)) return 1;
    if (hankel_i(p, dir, p->C0, p->C1, nb, st)) return 1;
    return transpose_i(p, p->C1, (double2*)out, p->NLM, S, st);
}
int xfb_ft(xfb_plan* p, int32_t dir, const double* in, double* out, int32_t nb, void* stream) {
    if (nb < 1 || nb > p->max_batch) XFB_FAIL("n_batch=%d outside plan capacity", nb);
    return ft_i(p, dir, flat_view((const double2*)in, p->G), (double2*)out, nb, (cudaStream_t)stream);
}
int xfb_project_invariants(xfb_plan* p, const double* in, double* out, int32_t nb, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (nb < 1 || nb > p->max_batch) XFB_FAIL("n_batch=%d outside plan capacity", nb);
    const int S = nb * p->n_r;
    if (transpose_i(p, (const double2*)in, p->C0, S, p->NLM, st)) return 1;
    if (project_i(p, p->C0, p->C1, nb, st)) return 1;
    return transpose_i(p, p->C1, (double2*)out, p->NLM, S, st);
}
// fxs_unknowns of the LAST invariant projection (approximate_unknowns, fxs_Projections.py:752-767):
// unk_l = U V^H of svd(PD_l I_l), complex [n_l][2l+1].  Recomputed on request from the retained M^T (p->g): the Jacobi
// kernel is run once more for that run with the accumulator started from the identity, so that it yields J itself.
int xfb_get_unknowns(xfb_plan* p, int32_t run, int32_t order, double* out_dev, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (p->dims != 3) XFB_FAIL("xfb_get_unknowns is the 3-D entry point; use xfb_get_unknowns_2d");
    if (!p->has_proj) XFB_FAIL("projection constants not set");
    if (p->proj_calls == 0) XFB_FAIL("xfb_get_unknowns: no invariant projection has run yet");
    if (run < 0 || run >= p->last_proj_nb) XFB_FAIL("run=%d outside the last projected batch (%d)", run, p->last_proj_nb);
    if (order < 0 || order > p->L) XFB_FAIL("order=%d outside 0..%d", order, p->L);
    std::vector<int> kind(p->L + 1), act(p->L + 1);
    XFB_CUDA(cudaMemcpy(kind.data(), p->kind_dev, kind.size() * sizeof(int), cudaMemcpyDeviceToHost));
    XFB_CUDA(cudaMemcpy(act.data(), p->act_index_dev, act.size() * sizeof(int), cudaMemcpyDeviceToHost));
    const int kd = kind[order], n_c = 2 * order + 1;
    if (kd == ORD_PASS) XFB_FAIL("order %d is not a used order", order);
    if (kd == ORD_ZEROTH) {
        unknown_zeroth_kernel<<<1, 128, 0, st>>>(p->i00 + (size_t)run * p->n_r, p->v0_dev, p->q_pts, p->n_r, (double2*)out_dev);
        XFB_CUDA(cudaGetLastError());
        return 0;
    }
    if (kd == ORD_ZERO) {   // PD_l = 0 [n_cols x N_r]: numpy's svd of the zero [n_cols x (2l+1)] matrix returns identity factors
        const int n_l = p->ncols_all[order];
        unknown_identity_kernel<<<cdiv(n_l * n_c, 256), 256, 0, st>>>((double2*)out_dev, n_l, n_c);
        XFB_CUDA(cudaGetLastError());
        return 0;
    }
    const int na = (int)p->orders.size();
    if (p->unk_run != run || p->unk_stamp != p->proj_calls) {
        if (launch_jacobi(p, p->g + (size_t)run * p->g_run, p->gn_u, p->pp_u, p->sigma_u, 1, nullptr, st)) return 1;
        p->unk_run = run; p->unk_stamp = p->proj_calls;
    }
    const ProcOrder o = p->orders[act[order]];
    unknown_assemble_kernel<<<o.n_cols, 128, 0, st>>>(p->gn_u + o.g_off, p->pp_u + o.g_off, jacobi_stride(o.n_c), o.n_cols, o.n_c, o.l,
                                                      (double2*)out_dev);
    XFB_CUDA(cudaGetLastError());
    return 0;
}

// ---- degree-2 invariants (fxs_invariant_tools.py:915-923) and deg2_invariant_l2_diff (fxs_IO_methods.py:412-447) ----------
int xfb_plan_set_deg2_reference(xfb_plan* p, const double* bref_host, const double* norm_host) {
    if (!p || !bref_host || !norm_host) XFB_FAIL("null argument");
    if (p->dims != 3) XFB_FAIL("xfb_plan_set_deg2_reference: 3-D plans only");
    graphs_invalidate(p);
    if (!p->has_proj) XFB_FAIL("set the projection constants first (the radial mask is shared)");
    const size_t n = (size_t)(p->L + 1) * p->n_r * p->n_r;
    if (!p->d2_ref) { if (dev_alloc(p, &p->d2_ref, n)) return 1; if (dev_alloc(p, &p->d2_norm, (size_t)p->L + 1)) return 1; }
    XFB_CUDA(cudaMemcpy(p->d2_ref, bref_host, n * sizeof(double), cudaMemcpyHostToDevice));
    XFB_CUDA(cudaMemcpy(p->d2_norm, norm_host, (size_t)(p->L + 1) * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}
int xfb_deg2_invariants(xfb_plan* p, const double* direct_in, double* bl_out, int32_t nb, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (nb < 1 || nb > p->max_batch) XFB_FAIL("n_batch=%d outside plan capacity", nb);
    const int S = nb * p->n_r;
    if (transpose_i(p, (const double2*)direct_in, p->C0, S, p->NLM, st)) return 1;
    if (deg2_invariants_i(p, p->C0, nb, st)) return 1;
    XFB_CUDA(cudaMemcpyAsync(bl_out, p->d2_b, (size_t)nb * (p->L + 1) * p->n_r * p->n_r * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return 0;
}
int xfb_deg2_invariant_diff(xfb_plan* p, const double* direct_in, double* err_out, int32_t nb, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (nb < 1 || nb > p->max_batch) XFB_FAIL("n_batch=%d outside plan capacity", nb);
    const int S = nb * p->n_r;
    if (transpose_i(p, (const double2*)direct_in, p->C0, S, p->NLM, st)) return 1;
    return deg2_diff_i(p, p->C0, nb, err_out, p->L + 1, st);
}
int xfb_mtip_enable_deg2_metric(xfb_plan* p, int32_t on, int32_t history_capacity) {
    if (!on) { p->d2_metric = false; return 0; }
    if (!p->d2_ref) XFB_FAIL("deg2 metric: reference invariants not set (xfb_plan_set_deg2_reference)");
    if (history_capacity < 1) XFB_FAIL("deg2 metric: history capacity must be >= 1");
    if (p->d2_hist && history_capacity > p->d2_hist_cap) { cudaFree(p->d2_hist); p->d2_hist = nullptr; }
    if (!p->d2_hist) {
        if (dev_alloc(p, &p->d2_hist, (size_t)p->max_batch * history_capacity * (p->L + 1))) return 1;
        p->d2_hist_cap = history_capacity;
    }
    p->d2_metric = true;
    return 0;
}
// out_dev [n_batch][capacity][L+1] (the first min(iterations done, capacity) rows of every run are valid)
int xfb_mtip_get_deg2_errors(xfb_plan* p, double* out_dev, int32_t capacity, void* stream) {
    if (!p->d2_hist) XFB_FAIL("deg2 metric was not enabled");
    if (capacity != p->d2_hist_cap) XFB_FAIL("capacity %d != the enabled history capacity %d", capacity, p->d2_hist_cap);
    XFB_CUDA(cudaMemcpyAsync(out_dev, p->d2_hist, (size_t)p->n_batch * capacity * (p->L + 1) * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

int xfb_modify_intensity(xfb_plan* p, const double* rho_hat, const double* i_proj, double* out, int32_t nb, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (nb < 1 || nb > p->max_batch) XFB_FAIL("n_batch=%d outside plan capacity", nb);
    XFB_LAUNCH(p, PG_POINTWISE, st,
               modify_intensity_kernel<<<dim3(ew_blocks(p->G), nb), 256, 0, st>>>((const double2*)rho_hat, (const double2*)i_proj,
                                                                                  flat_view((const double2*)out, p->G), p->G));
    return 0;
}
int xfb_real_update(xfb_plan* p, int32_t method, double beta, const double* rho_ift, const double* rho_rt, const double* rho_prev,
                    const uint8_t* support, const int32_t* enforce, double* rho_next, double* err, int32_t nb, void* stream) {
    if (nb < 1 || nb > p->max_batch) XFB_FAIL("n_batch=%d outside plan capacity", nb);
    return real_update_i(p, method, beta, (const double2*)rho_ift, (const double2*)rho_rt, flat_view((const double2*)rho_prev, p->G),
                         flat_view((const double2*)rho_next, p->G), support, nullptr, 0, enforce, err, nb, (cudaStream_t)stream);
}
int xfb_shrinkwrap(xfb_plan* p, const double* rho, double sigma, double threshold, uint8_t* support_out, int32_t nb, void* stream) {
    if (nb < 1 || nb > p->max_batch) XFB_FAIL("n_batch=%d outside plan capacity", nb);
    return shrinkwrap_i(p, flat_view((const double2*)rho, p->G), sigma, threshold, support_out, nullptr, 0, nb, (cudaStream_t)stream);
}

// ---- loop level -----------------------------------------------------------------------------
static SlotView pool_view(double2* pool, const int* slot, const xfb_plan* p) {
    SlotView v; v.base = pool; v.slot = slot; v.slot_stride = (long long)p->max_batch * p->G; v.run_stride = p->G; return v;
}

int xfb_mtip_init(xfb_plan* p, const double* rho0, int32_t nb, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (nb < 1 || nb > p->max_batch) XFB_FAIL("n_batch=%d outside plan capacity", nb);
    if (!p->has_proj || !p->has_real) XFB_FAIL("plan needs projection and real options before xfb_mtip_init");
    if (ensure_loop_alloc(p) || ensure_reduce_alloc(p)) return 1;
    // everything that allocates or synchronises on the first use of a batch size happens here, not inside an iteration: tile lists
    // of the Hankel GEMM for the batch and for the chunk sizes of the host pipeline, descriptors of the grouped GEMMs
    if (hankel_tiles_i(p, nb, st)) return 1;
    for (int n : host_chunk_sizes(p, nb, 1)) if (hankel_tiles_i(p, n, st)) return 1;
    if (p->dims == 3 && !p->orders.empty() && build_gemm_groups(p, nb, st)) return 1;
    p->n_batch = nb; p->it_done = 0;
    const int B = p->max_batch, tb = 128, gb = cdiv(B, tb);
    auto fill_i = [&](int* q, int v) { fill_i32_kernel<<<gb, tb, 0, st>>>(q, v, B); p->launches++; };
    fill_i(p->ls.rho_cur, 0); fill_i(p->ls.rho_best, 0); fill_i(p->ls.rho_next, 1);
    fill_i(p->ls.rh_cur, 0); fill_i(p->ls.rh_best, 0); fill_i(p->ls.rh_next, 1);
    fill_i(p->ls.mask_cur, 0); fill_i(p->ls.mask_best, 0); fill_i(p->ls.mask_next, 1);
    fill_i(p->ls.enforce_cur, 1); fill_i(p->ls.enforce_best, 1); fill_i(p->ls.enforce_rep, 1); fill_i(p->ls.best_iter, 0); fill_i(p->ls.nonfinite, 0);
    p->outer_it = 0;
    fill_f64_kernel<<<gb, tb, 0, st>>>(p->ls.best_err, INFINITY, B);
    fill_f64_kernel<<<gb, tb, 0, st>>>(p->ls.last_err, INFINITY, B);
    fill_u8_kernel<<<ew_blocks((long long)B * p->G), 256, 0, st>>>(p->mask_pool, 1, (long long)B * p->G);
    p->launches += 3;
    XFB_CUDA(cudaGetLastError());
    // reconstruct.py:959-962: rho_hat0 = FT(rho0); rho = IFT(rho_hat0); both go to slot 0
    if (ft_i(p, 0, flat_view((const double2*)rho0, p->G), p->rh_pool, nb, st)) return 1;
    if (ft_i(p, 1, flat_view(p->rh_pool, p->G), p->rho_pool, nb, st)) return 1;
    return 0;
}

// one iteration of runs [b0, b0+nb) of the loop state; `it_index` is the error-history column it writes
static int iterate_range(xfb_plan* p, int b0, int nb, int method, int ft_stab, double beta, int it_index, cudaStream_t st) {
    if (ensure_reduce_alloc(p)) return 1;            // before the shift: lazily allocated scratch must exist at its base address
    if (p->d2_metric && ensure_deg2_alloc(p, st)) return 1;
    ScratchShift shift(p, b0);
    const int S = nb * p->n_r;
    const int eb = ew_blocks(p->G);
    const long long pool_stride = (long long)p->max_batch * p->G;
    auto pv = [&](double2* pool, const int* slot) {
        SlotView v; v.base = pool + (long long)b0 * p->G; v.slot = slot + b0; v.slot_stride = pool_stride; v.run_stride = p->G; return v;
    };
    LoopState ls = p->ls;
    ls.rho_cur += b0; ls.rho_best += b0; ls.rho_next += b0; ls.rh_cur += b0; ls.rh_best += b0; ls.rh_next += b0;
    ls.mask_cur += b0; ls.mask_best += b0; ls.mask_next += b0; ls.enforce_cur += b0; ls.enforce_best += b0; ls.enforce_rep += b0; ls.best_iter += b0; ls.nonfinite += b0;
    ls.best_err += b0; ls.last_err += b0; ls.hist += (long long)b0 * ls.hist_cap;
    double* err = p->err + 2 * b0;
    const bool fused = ft_stab && p->fused_ft_stab;
    NvtxRange nvtx_it(method == 0 ? "xfb:HIO iteration" : "xfb:ER iteration");
    // 1. rho_hat = FT(rho)                                   (reconstruct.py:585)
    nvtxRangePushA("xfb:FT(rho)");
    if (!fused) {
        if (ft_i(p, 0, pv(p->rho_pool, p->ls.rho_cur), p->W0, nb, st)) return 1;
    } else {
        if (sht_forward_i(p, pv(p->rho_pool, p->ls.rho_cur), p->n_r, p->C0, S, st)) return 1;
        if (hankel_i(p, 0, p->C0, p->C1, nb, st)) return 1;
        // C1 = SHT(rho_hat) (exact Gauss quadrature of a band-limited field): shell 0 of IFT(rho_hat) from it
        if (ift_shell0_i(p, p->C1, nb, st)) return 1;
        if (sht_inverse_i(p, p->C1, p->W0, S, st)) return 1;
    }
    nvtxRangePop();
    // 2. |rho_hat|^2 -> I_lm                                 (:519-520)
    nvtxRangePushA("xfb:MTIP_start (SHT, invariant projection, modified intensity)");
    // with the register phi-FFT both pointwise kernels are fused into the transforms: |.|^2 when the rows are loaded,
    // the modified-intensity formula when the synthesised rows are stored
    const bool fuse_pw = p->dims == 3 && fft2_covers(p->n_phi, p->n_theta);
    if (p->non_fxs) {      // MTIP_start_non_FXS (reconstruct.py:529-534): no invariant projection, fixed intensity instead
        if (!p->fix_int) XFB_FAIL("non-FXS iteration without a fixed intensity (xfb_mtip_fix_intensity)");
        XFB_LAUNCH(p, PG_POINTWISE, st,
                   fixed_intensity_kernel<<<dim3(eb, nb), 256, 0, st>>>(p->W0, p->fix_int, pv(p->rh_pool, p->ls.rh_next), p->G));
    } else {
    int half = 0;                                             // |rho_hat|^2 is real: half spectrum (2-D: the 'real' transform, reconstruct.py:347-348)
    if (fuse_pw) {
        if (sht_forward_i(p, flat_view(p->W0, p->G), p->n_r, p->C0, S, st, nullptr, 1, &half, 1)) return 1;
    } else {
        XFB_LAUNCH(p, PG_POINTWISE, st, square_kernel<<<ew_blocks((long long)nb * p->G), 256, 0, st>>>(p->W0, p->W1, (long long)nb * p->G));
        if (sht_forward_i(p, flat_view(p->W1, p->G), p->n_r, p->C0, S, st, nullptr, 1, &half)) return 1;
    }
    // optional reciprocal metric deg2_invariant_l2_diff of the current iterate's I_lm (MTIP_start sketch, reconstruct.py:526)
    if (p->d2_metric && it_index < p->d2_hist_cap) {
        const long long rs = (long long)p->d2_hist_cap * (p->L + 1);
        if (deg2_diff_i(p, p->C0, nb, p->d2_hist + (long long)b0 * rs + (long long)it_index * (p->L + 1), rs, st)) return 1;
    }
    // 3. projection onto the invariants                      (:521-523)
    if (project_i(p, p->C0, p->C1, nb, st, half)) return 1;
    // 4. I_proj on the grid, modified intensity              (:524-525)
    if (fuse_pw) {
        if (sht_inverse_i(p, p->C1, p->W1, S, st, half, p->W0, pv(p->rh_pool, p->ls.rh_next), p->n_r)) return 1;
    } else {
        if (sht_inverse_i(p, p->C1, p->W1, S, st, p->dims == 2 ? 1 : half)) return 1;
        XFB_LAUNCH(p, PG_POINTWISE, st,
                   modify_intensity_kernel<<<dim3(eb, nb), 256, 0, st>>>(p->W0, p->W1, pv(p->rh_pool, p->ls.rh_next), p->G));
    }
    }
    nvtxRangePop();
    // 5. back to real space (+ ft_stab correction terms)     (:586-588 / :579)
    NvtxRange nvtx_real("xfb:IFT + real projection + HIO/ER + error");
    // 6. real projection + HIO/ER + error                    (:589-590)
    const uint8_t* mask = p->mask_pool + (long long)b0 * p->G;
    if (fused) {
        // IFT is linear: IFT(rho_hat') + (rho - IFT(rho_hat)) = IFT(rho_hat' - rho_hat) + rho   (r >= 1)
        if (ft_i(p, 1, pv(p->rh_pool, p->ls.rh_next), p->W1, nb, st, p->W0)) return 1;
        if (real_update_i(p, method, beta, p->W1, nullptr, pv(p->rho_pool, p->ls.rho_cur), pv(p->rho_pool, p->ls.rho_next), mask,
                          ls.mask_cur, pool_stride, ls.enforce_cur, err, nb, st, p->rt0)) return 1;
    } else {
        if (ft_i(p, 1, pv(p->rh_pool, p->ls.rh_next), p->W1, nb, st)) return 1;
        if (ft_stab) { if (ft_i(p, 1, flat_view(p->W0, p->G), p->W2, nb, st)) return 1; }
        if (real_update_i(p, method, beta, p->W1, ft_stab ? p->W2 : nullptr, pv(p->rho_pool, p->ls.rho_cur),
                          pv(p->rho_pool, p->ls.rho_next), mask, ls.mask_cur, pool_stride, ls.enforce_cur, err, nb, st)) return 1;
    }
    // 7. bookkeeping                                         (:924-939)
    XFB_LAUNCH(p, PG_MISC, st, loop_update_kernel<<<cdiv(nb, 128), 128, 0, st>>>(ls, err, it_index, nb, p->outer_it, p->ip_active));
    return 0;
}

int xfb_mtip_iterate(xfb_plan* p, int32_t method, int32_t ft_stab, int32_t n_iter, const double* betas, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = p->n_batch;
    if (nb < 1) XFB_FAIL("xfb_mtip_init has not been called");
    // Dual mode: the two halves of the batch run their iterations on two streams, the second half one projection behind the
    // first, so that the latency-bound Jacobi kernel of one half (on its share of the SMs) overlaps the HBM-bound transform
    // kernels of the other half.  The runs are independent and every run's arithmetic is unchanged (bit-identical results).
    const bool dual = p->dual && p->dims == 3 && !p->jacobi_big && nb >= p->dual_min && n_iter > 0 && (!ft_stab || p->fused_ft_stab) && !p->sht_chunk;
    if (!dual) {
        // graph replay needs: no per-launch profiling events, no host-side history offsets (deg2 metric), a 3-D plan
        const bool graphable = p->use_graph && !p->graph_broken && !p->prof && !p->d2_metric && p->dims == 3 && !p->sht_chunk;
        const int key = (method & 1) | ((ft_stab & 1) << 1) | ((p->non_fxs ? 1 : 0) << 2) | (nb << 3);
        // the iterations of a graphable call run on an internal stream forked from the caller's (the legacy default stream cannot be captured)
        cudaStream_t ws = st;
        if (graphable) {
            if (!p->s_graph) {
                XFB_CUDA(cudaStreamCreateWithFlags(&p->s_graph, cudaStreamNonBlocking));
                XFB_CUDA(cudaEventCreateWithFlags(&p->ev_gfork, cudaEventDisableTiming));
                XFB_CUDA(cudaEventCreateWithFlags(&p->ev_gjoin, cudaEventDisableTiming));
            }
            XFB_CUDA(cudaEventRecord(p->ev_gfork, st));
            XFB_CUDA(cudaStreamWaitEvent(p->s_graph, p->ev_gfork, 0));
            ws = p->s_graph;
        }
        for (int it = 0; it < n_iter; ++it) {
            const double beta = betas ? betas[it] : 0.0;
            auto g = graphable ? p->graphs.find(key) : p->graphs.end();
            if (graphable && !p->graph_broken && g == p->graphs.end() && p->graph_warm[key] > 0) {
                // second iteration of this kind: capture it (nothing executes during the capture), instantiate, then replay
                if (!p->iter_params) { if (dev_alloc(p, &p->iter_params, 1)) return 1; }
                const int64_t l0 = p->launches;
                p->ip_active = p->iter_params;
                cudaGraph_t graph = nullptr;
                cudaError_t e = cudaStreamBeginCapture(ws, cudaStreamCaptureModeThreadLocal);
                int rc = 1;
                if (e == cudaSuccess) {
                    rc = iterate_range(p, 0, nb, method, ft_stab, beta, p->it_done, ws);
                    e = cudaStreamEndCapture(ws, &graph);
                }
                p->ip_active = nullptr;
                cudaGraphExec_t exec = nullptr;
                if (rc == 0 && e == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
                    p->graph_launches[key] = (int)(p->launches - l0);
                    p->launches = l0;
                    g = p->graphs.emplace(key, exec).first;
                } else {
                    cudaGetLastError();                   // capture not possible here: stay on the eager path for good
                    p->graph_broken = true;
                    p->launches = l0;
                }
                if (graph) cudaGraphDestroy(graph);
            }
            if (graphable && !p->graph_broken && g != p->graphs.end()) {
                iter_params_kernel<<<1, 1, 0, ws>>>(p->iter_params, beta, p->it_done, p->outer_it);
                XFB_CUDA(cudaGraphLaunch(g->second, ws));
                p->launches += p->graph_launches[key] + 1;
                p->graph_replays++;
                p->proj_calls++;                          // host bookkeeping of the replayed projection (cache stamp of xfb_get_unknowns)
            } else {
                if (iterate_range(p, 0, nb, method, ft_stab, beta, p->it_done, ws)) return 1;
                p->graph_warm[key]++;
            }
            p->it_done++;
        }
        if (graphable) {
            XFB_CUDA(cudaEventRecord(p->ev_gjoin, ws));
            XFB_CUDA(cudaStreamWaitEvent(st, p->ev_gjoin, 0));
        }
        return 0;
    }
    if (!p->s_half) {
        XFB_CUDA(cudaStreamCreateWithFlags(&p->s_half, cudaStreamNonBlocking));
        XFB_CUDA(cudaEventCreateWithFlags(&p->ev_hfork, cudaEventDisableTiming));
        XFB_CUDA(cudaEventCreateWithFlags(&p->ev_hjoin, cudaEventDisableTiming));
        XFB_CUDA(cudaEventCreateWithFlags(&p->ev_stagger, cudaEventDisableTiming));
    }
    const int n0 = (nb + 1) / 2, n1 = nb - n0;
    XFB_CUDA(cudaEventRecord(p->ev_hfork, st));
    XFB_CUDA(cudaStreamWaitEvent(p->s_half, p->ev_hfork, 0));
    p->dual_active = true;
    int rc = 0;
    for (int it = 0; it < n_iter && !rc; ++it) {
        const double beta = betas ? betas[it] : 0.0;
        p->ctx = 0;
        if (it == 0) p->stagger_pending = p->ev_stagger;            // recorded right before the first half's first Jacobi launch
        rc = iterate_range(p, 0, n0, method, ft_stab, beta, p->it_done + it, st);
        if (!rc && it == 0) {
            if (p->stagger_pending) { XFB_CUDA(cudaEventRecord(p->ev_stagger, st)); p->stagger_pending = nullptr; }   // (no Jacobi in this iteration)
            XFB_CUDA(cudaStreamWaitEvent(p->s_half, p->ev_stagger, 0));
        }
        p->ctx = 1;
        if (!rc) rc = iterate_range(p, n0, n1, method, ft_stab, beta, p->it_done + it, p->s_half);
    }
    p->ctx = 0; p->dual_active = false; p->stagger_pending = nullptr;
    XFB_CUDA(cudaEventRecord(p->ev_hjoin, p->s_half));
    XFB_CUDA(cudaStreamWaitEvent(st, p->ev_hjoin, 0));
    if (rc) return 1;
    p->it_done += n_iter;
    return 0;
}

// dual mode switches: enable, minimum batch, SMs given to the one-problem-per-SM and the two-problems-per-SM Jacobi launches
int xfb_plan_set_dual_stream(xfb_plan* p, int32_t on, int32_t min_batch, int32_t big_sms, int32_t small_sms) {
    if (!p) XFB_FAIL("null plan");
    p->dual = on != 0;
    if (min_batch > 1) p->dual_min = min_batch;
    if (big_sms > 0) p->dual_big_sms = big_sms;
    if (small_sms > 0) p->dual_small_sms = small_sms;
    return 0;
}

// diagnostics: Jacobi sweeps used by the last projection, host array [n_batch][n_active_orders] (orders largest first)
int xfb_debug_jacobi_sweeps(xfb_plan* p, int32_t* out_host, int32_t capacity, int32_t* n_orders_out, int32_t* orders_out) {
    const int na = (int)p->orders.size();
    if (n_orders_out) *n_orders_out = na;
    for (int i = 0; i < na && orders_out; ++i) orders_out[i] = p->orders[i].l;
    const int n = std::min<int>(capacity, na * std::max(1, p->gemm_nb));
    if (n > 0 && out_host) XFB_CUDA(cudaMemcpy(out_host, p->sweeps_dev, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}

// Second precompiled kernel of the GPU-access layer: the `apply_matrix` demo of the reference's framework test
// (tests/test_framework_integration.py:230-309): out[i][j] = sum_q matrix[i][q] vect[q][j], q ascending with one FMA per term (the
// OpenCL kernel's loop under the default FP contraction).  No plan needed; all pointers are device pointers.
__global__ void apply_matrix_kernel(double* __restrict__ out, const double* __restrict__ matrix, const double* __restrict__ vect, long long nq, long long nvec) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= nq * nvec) return;
    const long long i = idx / nvec, j = idx - i * nvec;
    double value = 0.0;
    for (long long q = 0; q < nq; ++q) value = fma(matrix[i * nq + q], vect[q * nvec + j], value);
    out[idx] = value;
}
int xfb_apply_matrix(const double* matrix_dev, const double* vect_dev, double* out_dev, int64_t nq, int64_t nvec, void* stream) {
    if (!matrix_dev || !vect_dev || !out_dev || nq < 1 || nvec < 1) XFB_FAIL("xfb_apply_matrix: bad arguments");
    const long long n = nq * nvec;
    apply_matrix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out_dev, matrix_dev, vect_dev, nq, nvec);
    XFB_CUDA(cudaGetLastError());
    return 0;
}

// diagnostics: iterations replayed from a captured CUDA graph so far
int64_t xfb_plan_graph_replays(const xfb_plan* p) { return p ? p->graph_replays : 0; }

// diagnostics: phase cycle counters of the Jacobi kernel (zeros unless the library was built with -DJAC_TIMING); reset after reading
int xfb_debug_jacobi_phase_cycles(double* out8_host) {
    unsigned long long h[8] = {};
    XFB_CUDA(cudaDeviceSynchronize());
    XFB_CUDA(cudaMemcpyFromSymbol(h, g_jac_phase, sizeof(h)));
    for (int i = 0; i < 8; ++i) out8_host[i] = (double)h[i];
    unsigned long long z[8] = {};
    XFB_CUDA(cudaMemcpyToSymbol(g_jac_phase, z, sizeof(z)));
    return 0;
}

int xfb_plan_set_fused_ft_stab(xfb_plan* p, int32_t on) { graphs_invalidate(p); p->fused_ft_stab = (on && p->dims == 3) ? 1 : 0; return 0; }

// L2-resident phi-Fourier intermediate: runs per chunk (0 = one launch over the whole batch, the round-1
// behaviour) and number of streams the chunks are spread over (1 .. 4).
int xfb_plan_set_sht_chunk(xfb_plan* p, int32_t runs_per_chunk, int32_t streams) {
    if (!p) XFB_FAIL("null plan");
    if (runs_per_chunk < 0 || streams < 1 || streams > 4) XFB_FAIL("sht chunk: runs_per_chunk >= 0 and 1 <= streams <= 4");
    graphs_invalidate(p);
    p->sht_chunk = runs_per_chunk; p->sht_streams = streams;
    return 0;
}

int xfb_mtip_shrinkwrap(xfb_plan* p, double sigma, double threshold, double error_limit, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = p->n_batch;
    if (nb < 1) XFB_FAIL("xfb_mtip_init has not been called");
    if (shrinkwrap_i(p, pool_view(p->rho_pool, p->ls.rho_cur, p), sigma, threshold, p->mask_pool, p->ls.mask_next,
                     (long long)p->max_batch * p->G, nb, st)) return 1;
    XFB_LAUNCH(p, PG_MISC, st, sw_update_kernel<<<cdiv(nb, 128), 128, 0, st>>>(p->ls, error_limit, p->it_done > 0 ? 1 : 0, nb));
    return 0;
}

// ---- sketch / option tail of the loop driver ---------------------------------------------------------------------------
// sub-loop iteration index recorded with a new best error (state['best_iteration'], reconstruct.py:938)
int xfb_mtip_set_outer_iteration(xfb_plan* p, int32_t outer_iteration) { p->outer_it = outer_iteration; return 0; }

// end of a sub-loop with a finite best_density_not_in_first_n_iterations (reconstruct.py:945-949)
int xfb_mtip_select_best(xfb_plan* p, int32_t n_first, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = p->n_batch;
    if (nb < 1) XFB_FAIL("xfb_mtip_init has not been called");
    XFB_LAUNCH(p, PG_MISC, st, select_best_kernel<<<cdiv(nb, 128), 128, 0, st>>>(p->ls, n_first, nb));
    return 0;
}

// non-FXS methods (HIO_non_FXS / ER_non_FXS, reconstruct.py:899-904): the intensity is fixed to |reciprocal density| of the pair
// the reference's `hist` variable names when the block starts.  snapshot: candidate <- |current reciprocal density| (the host
// takes it at the start of a sub-loop and before the last iteration of every HIO / ER block); fix: fixed <- candidate.
int xfb_mtip_snapshot_intensity(xfb_plan* p, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = p->n_batch;
    if (nb < 1) XFB_FAIL("xfb_mtip_init has not been called");
    if (!p->fix_cand) { if (dev_alloc(p, &p->fix_cand, (size_t)p->max_batch * p->G)) return 1; }
    XFB_LAUNCH(p, PG_POINTWISE, st, abs_real_kernel<<<dim3(ew_blocks(p->G), nb), 256, 0, st>>>(pool_view(p->rh_pool, p->ls.rh_cur, p), p->fix_cand, p->G));
    return 0;
}
int xfb_mtip_fix_intensity(xfb_plan* p, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!p->fix_cand) XFB_FAIL("xfb_mtip_fix_intensity: no intensity snapshot has been taken");
    if (!p->fix_int) { if (dev_alloc(p, &p->fix_int, (size_t)p->max_batch * p->G)) return 1; }
    XFB_CUDA(cudaMemcpyAsync(p->fix_int, p->fix_cand, (size_t)p->n_batch * p->G * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return 0;
}
int xfb_mtip_set_non_fxs(xfb_plan* p, int32_t on) { p->non_fxs = on != 0; return 0; }

// SW_center (reconstruct.py:606-613,886-897): shrink-wrap of the current density x, then the history pair becomes
// (reciprocal = x, real = FT(x)) -- the reference's sketch returns (support, x, FT(x)) and the loop binds it to
// (support, ft_density, density); see oracle/mtip.py.  Called once per repeat.
int xfb_mtip_shrinkwrap_center(xfb_plan* p, double sigma, double threshold, double error_limit, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = p->n_batch;
    if (nb < 1) XFB_FAIL("xfb_mtip_init has not been called");
    if (p->it_done < 1) XFB_FAIL("SW_center before the first iteration: the reference reads error_dict['main'][-1] (reconstruct.py:887)");
    if (xfb_mtip_shrinkwrap(p, sigma, threshold, error_limit, stream)) return 1;
    const int eb = ew_blocks(p->G);
    XFB_LAUNCH(p, PG_MISC, st, copy_slot_kernel<<<dim3(eb, nb), 256, 0, st>>>(pool_view(p->rho_pool, p->ls.rho_cur, p), pool_view(p->rh_pool, p->ls.rh_next, p), p->G));
    if (ft_i(p, 0, pool_view(p->rho_pool, p->ls.rho_cur, p), p->W2, nb, st)) return 1;
    XFB_LAUNCH(p, PG_MISC, st, scatter_slot_kernel<<<dim3(eb, nb), 256, 0, st>>>(p->W2, pool_view(p->rho_pool, p->ls.rho_next, p), p->G));
    XFB_LAUNCH(p, PG_MISC, st, rotate_pair_kernel<<<cdiv(nb, 128), 128, 0, st>>>(p->ls, nb));
    return 0;
}

int xfb_mtip_get_grid(xfb_plan* p, int32_t which, void* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = p->n_batch;
    if (nb < 1) XFB_FAIL("xfb_mtip_init has not been called");
    const int eb = ew_blocks(p->G);
    switch (which) {
        case 0: XFB_LAUNCH(p, PG_MISC, st, gather_slot_kernel<<<dim3(eb, nb), 256, 0, st>>>(pool_view(p->rho_pool, p->ls.rho_cur, p), (double2*)out, p->G)); break;
        case 1: XFB_LAUNCH(p, PG_MISC, st, gather_slot_kernel<<<dim3(eb, nb), 256, 0, st>>>(pool_view(p->rh_pool, p->ls.rh_cur, p), (double2*)out, p->G)); break;
        case 2: XFB_LAUNCH(p, PG_MISC, st, gather_slot_kernel<<<dim3(eb, nb), 256, 0, st>>>(pool_view(p->rho_pool, p->ls.rho_best, p), (double2*)out, p->G)); break;
        case 3: XFB_LAUNCH(p, PG_MISC, st, gather_slot_kernel<<<dim3(eb, nb), 256, 0, st>>>(pool_view(p->rh_pool, p->ls.rh_best, p), (double2*)out, p->G)); break;
        case 4: XFB_LAUNCH(p, PG_MISC, st, effective_support_kernel<<<dim3(eb, nb), 256, 0, st>>>(p->mask_pool, p->ls.mask_cur, p->ls.enforce_rep, p->init_support_dev, (long long)p->max_batch * p->G, p->G, (uint8_t*)out)); break;
        case 5: XFB_LAUNCH(p, PG_MISC, st, effective_support_kernel<<<dim3(eb, nb), 256, 0, st>>>(p->mask_pool, p->ls.mask_best, p->ls.enforce_best, p->init_support_dev, (long long)p->max_batch * p->G, p->G, (uint8_t*)out)); break;
        default: XFB_FAIL("which=%d unknown", which);
    }
    return 0;
}

// End-to-end step with HOST buffers (pinned recommended): H2D of the batch's densities, one iteration, D2H of the
// updated densities and of the per-run (numerator, denominator) of the real-space error.  The batch is cut into
// chunks that flow through three streams (copy-in | compute | copy-out) so PCIe transfers overlap the kernels.
int xfb_mtip_step_host(xfb_plan* p, int32_t method, int32_t ft_stab, double beta, const double* rho_in_host, double* rho_out_host,
                       double* err_out_host, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = p->n_batch;
    if (nb < 1) XFB_FAIL("xfb_mtip_init has not been called");
    const int eb = ew_blocks(p->G);
    if (!p->s_in) {
        XFB_CUDA(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
        XFB_CUDA(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
        XFB_CUDA(cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming));
        if (dev_alloc(p, &p->stage_out, (size_t)p->max_batch * p->G)) return 1;
    }
    const std::vector<int> sizes = host_chunk_sizes(p, nb, ft_stab);
    const int nchunk = (int)sizes.size();
    std::vector<int> first(nchunk, 0);
    for (int c = 1; c < nchunk; ++c) first[c] = first[c - 1] + sizes[c - 1];
    while ((int)p->ev_in.size() < nchunk) {
        cudaEvent_t a, b;
        XFB_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        XFB_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        p->ev_in.push_back(a); p->ev_comp.push_back(b);
    }
    // the side streams start after everything already queued on the caller's stream
    XFB_CUDA(cudaEventRecord(p->ev_start, st));
    XFB_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_start, 0));
    XFB_CUDA(cudaStreamWaitEvent(p->s_out, p->ev_start, 0));
    const double2* hin = (const double2*)rho_in_host;
    double2* hout = (double2*)rho_out_host;
    for (int c = 0; c < nchunk; ++c) {
        const int b0 = first[c], n = sizes[c];
        const size_t off = (size_t)b0 * p->G, bytes = (size_t)n * p->G * sizeof(double2);
        XFB_CUDA(cudaMemcpyAsync(p->W2 + off, hin + off, bytes, cudaMemcpyHostToDevice, p->s_in));
        XFB_CUDA(cudaEventRecord(p->ev_in[c], p->s_in));
    }
    const long long pool_stride = (long long)p->max_batch * p->G;
    for (int c = 0; c < nchunk; ++c) {
        const int b0 = first[c], n = sizes[c];
        const size_t off = (size_t)b0 * p->G, bytes = (size_t)n * p->G * sizeof(double2);
        XFB_CUDA(cudaStreamWaitEvent(st, p->ev_in[c], 0));
        SlotView cur; cur.base = p->rho_pool + off; cur.slot = p->ls.rho_cur + b0; cur.slot_stride = pool_stride; cur.run_stride = p->G;
        XFB_LAUNCH(p, PG_MISC, st, scatter_slot_kernel<<<dim3(eb, n), 256, 0, st>>>(p->W2 + off, cur, p->G));
        if (iterate_range(p, b0, n, method, ft_stab, beta, p->it_done, st)) return 1;
        XFB_LAUNCH(p, PG_MISC, st, gather_slot_kernel<<<dim3(eb, n), 256, 0, st>>>(cur, p->stage_out + off, p->G));
        XFB_CUDA(cudaEventRecord(p->ev_comp[c], st));
        XFB_CUDA(cudaStreamWaitEvent(p->s_out, p->ev_comp[c], 0));
        XFB_CUDA(cudaMemcpyAsync(hout + off, p->stage_out + off, bytes, cudaMemcpyDeviceToHost, p->s_out));
    }
    p->it_done++;
    XFB_CUDA(cudaMemcpyAsync(err_out_host, p->err, (size_t)nb * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    XFB_CUDA(cudaStreamSynchronize(st));
    XFB_CUDA(cudaStreamSynchronize(p->s_out));
    return 0;
}

// diagnostics: number of iterations whose error metric was NaN / inf, per run (host array [n_batch]); synchronises the stream
int xfb_mtip_get_nonfinite(xfb_plan* p, int32_t* out_host, void* stream) {
    if (p->n_batch < 1) XFB_FAIL("xfb_mtip_init has not been called");
    XFB_CUDA(cudaMemcpyAsync(out_host, p->ls.nonfinite, (size_t)p->n_batch * sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    XFB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

int xfb_plan_set_host_chunk(xfb_plan* p, int32_t runs) { if (runs < 1) XFB_FAIL("chunk must be >= 1"); p->host_chunk = runs; return 0; }

int xfb_mtip_get_errors(xfb_plan* p, double* hist, int32_t cap, double* best, int32_t* n_done, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = p->n_batch;
    if (nb < 1) XFB_FAIL("xfb_mtip_init has not been called");
    const int n = std::min(std::min(p->it_done, (int)cap), p->ls.hist_cap);
    if (hist && n > 0)
        XFB_CUDA(cudaMemcpy2DAsync(hist, (size_t)cap * sizeof(double), p->ls.hist, (size_t)p->ls.hist_cap * sizeof(double), (size_t)n * sizeof(double), nb,
                                   cudaMemcpyDeviceToDevice, st));
    if (best) XFB_CUDA(cudaMemcpyAsync(best, p->ls.best_err, nb * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (n_done) *n_done = p->it_done;
    return 0;
}

}  // extern "C"
