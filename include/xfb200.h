/* xfb200 -- C-ABI of the B200-native fxs MTIP phasing path.
 *
 * This is the drop-in boundary for the ONE hot path this repository replaces:
 * the MTIP iterative phasing loop of `xframe fxs reconstruct`
 * (reference: xframe/projects/fxs/reconstruct.py:768-1036 and the operators it
 * wires together).  The reference has no FFI of its own
 * for this path (it is pure Python + shtns + an OpenCL RPC layer); each entry
 * point below names the reference interface it stands in for.  The Python
 * binding a maintainer adds on the xFrame side is a ctypes stub, shown in
 * INTEGRATION.md; xframe_b200/_lib.py is that stub, kept in this tree.
 *
 * Conventions
 *  - plain C: pointers + sizes, no torch / C++ types.
 *  - every function returns 0 on success, non-zero on failure;
 *    xfb_last_error() returns a thread-local message.
 *  - "dev" pointers are CUDA device pointers owned by the caller (e.g. a torch
 *    tensor's data_ptr()); the plan owns its tables and workspaces.
 *  - every compute call takes the cudaStream_t (as void*) to launch on and
 *    never synchronises, unless the name ends in _host.
 *  - complex numbers are interleaved (re,im) IEEE doubles = numpy complex128.
 *  - a plan is not thread-safe: one plan per stream.
 *
 * Layouts (reference shapes)
 *  grid   : [n_batch][N_r][n_theta][n_phi] complex128, phi contiguous
 *           (shtns_plugin.py:142,157; gridLibrary.py:940-949)
 *  direct : [n_batch][N_r][(L+1)^2] complex128, index l*(l+1)+m
 *           (shtns_plugin.py:110-112,250-261)
 *  mask   : [n_batch][N_r][n_theta][n_phi] uint8 (numpy bool)
 */
#ifndef XFB200_H
#define XFB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xfb_plan xfb_plan;

/* Host-side description of one transform plan. All table pointers are HOST
 * pointers and are copied; they are produced by xframe_b200/tables.py with the
 * formulas of the reference (cited there). */
typedef struct {
    int32_t l_max;          /* L: harmonic_transforms.py:62 `max_order`                       */
    int32_t n_r;            /* N_r: settings grid.n_radial_points                              */
    int32_t n_theta;        /* Gauss nodes, multiple of 8                                      */
    int32_t n_phi;          /* power of two in [16,512], > 2L                                  */
    int32_t max_batch;      /* capacity B: runs processed together on this GPU                 */
    int32_t hankel_skip;    /* 0 midpoint / gauss, 1 trapz / Zernike: input row 0 skipped
                               (hankel_transforms.py:648-652)                                  */
    const double* cos_theta;      /* [n_theta] north->south (shtns_plugin.py:133)              */
    const double* gauss_w;        /* [n_theta]                                                 */
    const double* legendre;       /* packed P_l^m(x_j), see tables.py:pack_legendre            */
    int64_t legendre_len;
    const double* hankel_w;       /* [L+1][n_k][N_r]: w[l,p,k], p summed (hankel_transforms.py:399-410) */
    int32_t hankel_n_sum;         /* number of summed rows p (N_r or N_r-1)                    */
    double hankel_fwd_scale;      /* (r_max/N)^3 sqrt(2/pi)  (hankel_transforms.py:444)        */
    double hankel_inv_scale;      /* (q_max/N)^3 sqrt(2/pi)  (hankel_transforms.py:445)        */
    const double* int_weight;     /* [N_r][n_theta] quadrature weight of SphericalIntegrator
                                     (mathLibrary.py:1223-1235), phi sum weight 1               */
    const double* r_points;       /* [N_r] real radial grid                                    */
    const double* q_points;       /* [N_r] reciprocal radial grid                              */
    int32_t dimensions;           /* 3 (or 0): spherical plan as described above.
                                     2: polar plan (settings `dimensions: 2`): n_theta = 1, n_phi = 2*l_max+1 angular points
                                     (harmonic_transforms.py:44-47,60), no Legendre tables; hankel_w is
                                     [n_phi][n_k][N_r], one matrix per DFT index j with the sign of the negative orders
                                     folded in (w_{-m} = (-1)^m w_m, hankel_transforms.py:441); scales are
                                     (r_max/N)^2, (q_max/N)^2 (:438-439); int_weight is [N_r][n_phi], the weights of
                                     PolarIntegrator (mathLibrary.py:1242-1265)                 */
} xfb_plan_desc;

/* Reciprocal-projection constants (fxs_Projections.py:471-537,679-714,753-754). */
typedef struct {
    int32_t n_orders;             /* number of used orders (used_order_ids = arange(n))        */
    const int32_t* n_cols;        /* [n_orders] n_l = columns of V_l                           */
    const double* const* v;       /* [n_orders] V_l real part, row-major [N_r][n_l] (already *2, odd->0, V_0 set) */
    const uint8_t* radial_mask;   /* [L+1][N_r]                                                */
    double sqrt_n_particles;      /* fxs_Projections.py:870                                    */
    double sv_cutoff;             /* relative singular-value cut-off of the polar factor        */
    int32_t max_sweeps;
} xfb_projection_desc;

/* Real-space projection options (fxs_Projections.py:72-130, pythonLibrary.py:1289-1318). */
enum { XFB_OP_SUPPORT = 1, XFB_OP_VALUE_THRESHOLD = 2, XFB_OP_LIMIT_IMAG = 3, XFB_OP_AVERAGE_CENTER = 4 };
typedef struct {
    int32_t n_ops;
    int32_t ops[4];               /* application order, XFB_OP_*                               */
    int32_t hio_considered[4];    /* flags per op position: part of the HIO mask ('all' = every op) */
    int32_t use_lo, use_hi;
    double lo, hi;                /* value_threshold                                           */
    double imag_limit;            /* limit_imag threshold                                      */
    int32_t error_inside_initial_support;   /* fxs_IO_methods.py:287-300                       */
    int32_t average_center_shells;          /* average_center.max_radial_id (fxs_Projections.py:96-110) */
} xfb_real_desc;

const char* xfb_last_error(void);
int xfb_device_count(int* n);                       /* Multiprocessing.get_number_of_gpus (Multiprocessing.py:892-898) */
int xfb_set_device(int dev);

int xfb_plan_create(xfb_plan** out, const xfb_plan_desc* desc);
int xfb_plan_destroy(xfb_plan* p);
int xfb_plan_set_projection(xfb_plan* p, const xfb_projection_desc* d);
/* 2-D reciprocal projection constants: v = V_m(q) complex [n_orders][N_r] after regrid / odd->0 / V_0 = <I>
 * (fxs_Projections.py:679-706), radial_mask [l_max+1][N_r], so_order_id = pinned order of SO_freedom or -1 (:743-748). */
int xfb_plan_set_projection_2d(xfb_plan* p, int32_t n_orders, const double* v, const uint8_t* radial_mask, double sqrt_n_particles,
                               int32_t so_order_id);
int xfb_get_unknowns_2d(xfb_plan* p, int32_t run, double* out_dev /*[n_orders] complex*/, void* stream);
int xfb_plan_set_real(xfb_plan* p, const xfb_real_desc* d, const uint8_t* initial_support_host /*[N_r][n_theta][n_phi]*/);
int64_t xfb_plan_workspace_bytes(const xfb_plan* p);
/* ft_stab sketch (reconstruct.py:584-593): 1 (default) evaluates IFT(rho_hat') + (rho - IFT(rho_hat)) as
 * IFT(rho_hat' - rho_hat) + rho (linearity; one inverse transform instead of two), 0 follows the sketch literally. */
int xfb_plan_set_fused_ft_stab(xfb_plan* p, int32_t on);
/* Two halves of the batch in flight on two streams inside xfb_mtip_iterate (the Jacobi kernel of one half overlaps the
 * HBM-bound transforms of the other): enable, minimum batch, SMs given to the two Jacobi launches (<= 0 keeps the value). */
int xfb_plan_set_dual_stream(xfb_plan* p, int32_t on, int32_t min_batch, int32_t big_sms, int32_t small_sms);
/* L2-resident phi-Fourier intermediate: runs per transform chunk (0 = unchunked) and streams (1..4) the chunks are spread over */
int xfb_plan_set_sht_chunk(xfb_plan* p, int32_t runs_per_chunk, int32_t streams);
/* diagnostics: Jacobi sweeps of the last projection, host array [n_batch][n_active_orders]; orders_out lists the orders */
/* diagnostics: cycles per phase of the Jacobi kernel summed over all problems since the last call (library built with -DJAC_TIMING) */
int xfb_debug_jacobi_phase_cycles(double* out8_host);
int xfb_debug_jacobi_sweeps(xfb_plan* p, int32_t* out_host, int32_t capacity, int32_t* n_orders_out, int32_t* orders_out);

/* ---- operator level: the harmonic-transform / Hankel / FT interfaces ---- */
/* sh.forward_d / sh.inverse_d  (shtns_plugin.py:250-261): n_shells = n_batch*N_r or any count <= max_batch*N_r */
int xfb_sht_forward(xfb_plan* p, const double* grid_dev, double* direct_dev, int32_t n_shells, void* stream);
int xfb_sht_inverse(xfb_plan* p, const double* direct_dev, double* grid_dev, int32_t n_shells, void* stream);
/* zht / izht of generate_spherical_ht_gpu (hankel_transforms.py:660-766); dir 0 forward, 1 inverse */
int xfb_hankel_apply(xfb_plan* p, int32_t dir, const double* direct_in_dev, double* direct_out_dev, int32_t n_batch, void* stream);
/* ft / ift of generate_ft (fourier_transforms.py:57-85) */
int xfb_ft(xfb_plan* p, int32_t dir, const double* grid_in_dev, double* grid_out_dev, int32_t n_batch, void* stream);
/* approximate_unknowns + mtip_projection (fxs_Projections.py:752-872) on 'direct' coefficients */
int xfb_project_invariants(xfb_plan* p, const double* direct_in_dev, double* direct_out_dev, int32_t n_batch, void* stream);
/* unknowns of the last xfb_project_invariants / iteration: out [n_l][2l+1] complex128 for (run,order) */
int xfb_get_unknowns(xfb_plan* p, int32_t run, int32_t order, double* out_dev, void* stream);
/* project_to_modified_intensity (fxs_Projections.py:899-909) */
/* B_l = I_l I_l^H of the harmonic coefficients of a REAL field (harmonic_coeff_to_deg2_invariants_3d, fxs_invariant_tools.py:915-923):
 * direct_in [nb][N_r][(L+1)^2] complex -> bl_out [nb][L+1][N_r][N_r] real (the invariants of a real field are real symmetric). */
int xfb_deg2_invariants(xfb_plan* p, const double* direct_in_dev, double* bl_out_dev, int32_t n_batch, void* stream);
/* deg2_invariant_l2_diff (fxs_IO_methods.py:412-447): reference = masked V_l V_l^H with order 0 already divided by the number of
 * particles [L+1][N_r][N_r] real, norms = sum |masked reference|^2 per order BEFORE that division; err_out [nb][L+1], -1 where norm = 0. */
int xfb_plan_set_deg2_reference(xfb_plan* p, const double* bref_host, const double* norm_host);
int xfb_deg2_invariant_diff(xfb_plan* p, const double* direct_in_dev, double* err_out_dev, int32_t n_batch, void* stream);
/* the same metric inside the loop (settings main_loop.error.methods.reciprocal.calculate: [deg2_invariant_l2_diff]) */
int xfb_mtip_enable_deg2_metric(xfb_plan* p, int32_t on, int32_t history_capacity);
int xfb_mtip_get_deg2_errors(xfb_plan* p, double* out_dev /*[n_batch][capacity][L+1]*/, int32_t capacity, void* stream);
int xfb_modify_intensity(xfb_plan* p, const double* rho_hat_dev, const double* i_proj_dev, double* out_dev, int32_t n_batch, void* stream);
/* real_projection + HIO/ER + l2_projection_diff (fxs_Projections.py:110-130, fxs_IO_methods.py:56-68,97-128).
 * method 0 = HIO, 1 = ER. rho_rt_dev may be NULL (no ft_stab). support_dev: 1 inside SW support.
 * err_dev: [n_batch][2] doubles (numerator, denominator). */
int xfb_real_update(xfb_plan* p, int32_t method, double beta, const double* rho_ift_dev, const double* rho_rt_dev,
                    const double* rho_prev_dev, const uint8_t* support_dev, const int32_t* enforce_initial_dev,
                    double* rho_next_dev, double* err_dev, int32_t n_batch, void* stream);
/* SW sketch (reconstruct.py:598-605, fxs_Projections.py:245-258,294-298) */
int xfb_shrinkwrap(xfb_plan* p, const double* rho_dev, double sigma, double threshold, uint8_t* support_out_dev,
                   int32_t n_batch, void* stream);

/* ---- loop level: device-resident batch of independent phasing runs (reconstruct.py:854-951) ---- */
int xfb_mtip_init(xfb_plan* p, const double* rho0_dev, int32_t n_batch, void* stream);   /* create_initial_state :957-979 */
/* n_iter iterations of HIO (0) / ER (1); betas: host array [n_iter] (ExponentialRamp values, :911) */
int xfb_mtip_iterate(xfb_plan* p, int32_t method, int32_t ft_stab, int32_t n_iter, const double* betas_host, void* stream);
/* SW step incl. enforce_initial_support decision (:877-885); error_limit = if_error_bigger_than or +inf */
int xfb_mtip_shrinkwrap(xfb_plan* p, double sigma, double threshold, double error_limit, void* stream);
/* One iteration driven from HOST buffers (the end-to-end shape of one reference `process.run`, reconstruct.py:924,
 * whose inputs and outputs are numpy arrays): H2D of rho [n_batch grids], iterate, D2H of rho_next and of
 * err [n_batch][2]. Synchronises `stream` before returning. */
int xfb_mtip_step_host(xfb_plan* p, int32_t method, int32_t ft_stab, double beta, const double* rho_in_host,
                       double* rho_out_host, double* err_out_host, void* stream);
/* runs per chunk of the copy-in | compute | copy-out pipeline inside xfb_mtip_step_host (default 16) */
int xfb_plan_set_host_chunk(xfb_plan* p, int32_t runs);
/* diagnostics: iterations with a NaN / inf error metric per run (host array [n_batch]); synchronises the stream */
int xfb_mtip_get_nonfinite(xfb_plan* p, int32_t* out_host, void* stream);
/* outputs; which: 0 last real, 1 last reciprocal, 2 best real, 3 best reciprocal (grids);
 *          4 last support, 5 best support (uint8 grids); */
/* sketch / option tail of the loop driver (reconstruct.py:529-534,606-613,886-904,945-949) */
int xfb_mtip_set_outer_iteration(xfb_plan* p, int32_t outer_iteration);        /* state['best_iteration'] bookkeeping */
int xfb_mtip_select_best(xfb_plan* p, int32_t n_first, void* stream);          /* finite best_density_not_in_first_n_iterations */
int xfb_mtip_snapshot_intensity(xfb_plan* p, void* stream);                    /* candidate <- |current reciprocal density| */
int xfb_mtip_fix_intensity(xfb_plan* p, void* stream);                         /* fixed intensity <- candidate (non-FXS methods) */
int xfb_mtip_set_non_fxs(xfb_plan* p, int32_t on);                             /* following iterations run MTIP_start_non_FXS */
int xfb_mtip_shrinkwrap_center(xfb_plan* p, double sigma, double threshold, double error_limit, void* stream);   /* SW_center */
int xfb_mtip_get_grid(xfb_plan* p, int32_t which, void* out_dev, void* stream);
/* error history [n_batch][n_done] (real l2_projection_diff), best_error [n_batch], n_done */
int xfb_mtip_get_errors(xfb_plan* p, double* hist_dev, int32_t hist_capacity, double* best_dev, int32_t* n_done_host, void* stream);
/* counts kernels launched through this plan since creation (bench.py gpu_launches) */
/* GPU-access layer, second named kernel: `apply_matrix` of the reference's framework demo (tests/test_framework_integration.py:230-309):
 * out[nq][nvec] = matrix[nq][nq] . vect[nq][nvec], float64, device pointers */
int xfb_apply_matrix(const double* matrix_dev, const double* vect_dev, double* out_dev, int64_t nq, int64_t nvec, void* stream);
int64_t xfb_plan_launch_count(const xfb_plan* p);
int64_t xfb_plan_graph_replays(const xfb_plan* p);   /* iterations replayed from a captured CUDA graph (diagnostics) */
/* elapsed ms of the dominant kernel group between reset and now, measured with CUDA events on `stream` */
int xfb_profile_enable(xfb_plan* p, int32_t on);
int xfb_profile_read(xfb_plan* p, int32_t n_max, char* names /*n_max*32*/, double* ms, int64_t* launches, int32_t* n_out);

#ifdef __cplusplus
}
#endif
#endif
