"""Parity of the CUDA path (through the C-ABI) with the oracle.  Tolerances are relative L2:
  transforms / Hankel / FT            <= 1e-12  (FP64 path; north_star asks <= 1e-5 fp32-equivalent, tighter for FP64)
  invariant projection                <= 1e-6   (the reference's LAPACK SVD leaves singular directions below
                                                 ~1e-15 sigma_max undetermined; orders where that matters agree to
                                                 ~3e-8, well-conditioned orders to 1e-12 -- see DESIGN.md)
  elementwise                         <= 1e-14
  full loop error history / densities <= 1e-6
"""
import numpy as np
import pytest
import torch

from helpers import load_golden, golden_settings, golden_data, rel_l2
from oracle import mtip as O

pytestmark = pytest.mark.gpu
CASES = ['ref_small_ftstab', 'ref_small_plain', 'ref_medium_ops']


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to('cuda')


def N(t):
    return t.cpu().numpy()


@pytest.fixture(scope='module', params=CASES)
def case(request):
    from xframe_b200.plan import Plan
    g = load_golden(request.param)
    sd = golden_settings(g)
    m = O.MTIP(sd, golden_data(g))
    plan = Plan(m.l_max, len(m.rs), float(g['max_q']), n_theta=int(g['n_theta']), n_phi=int(g['n_phi']), max_batch=3)
    plan.set_projection(m.rp.projection_matrices, m.rp.radial_mask, m.rp.number_of_particles[0])
    popt = sd['projections']['real']['projections']
    plan.set_real(popt['apply'], m.real_pr.initial_support, popt['value_threshold']['threshold'], popt['limit_imag']['threshold'])
    yield g, sd, m, plan
    plan.close()


def test_plan_grids_match_reference(case):
    g, sd, m, plan = case
    assert np.array_equal(plan.rs, g['rs']) and np.array_equal(plan.qs, g['qs'])
    assert np.allclose(plan.thetas, g['thetas'], atol=1e-15) and np.allclose(plan.phis, g['phis'], atol=1e-15)


def test_sht_golden(case):
    g, sd, m, plan = case
    assert rel_l2(N(plan.sht_forward(T(g['x_grid']))), g['sht_forward_direct']) < 1e-12
    assert rel_l2(N(plan.sht_inverse(T(g['sht_forward_direct']))), g['sht_inverse_of_forward']) < 1e-12


def test_hankel_golden(case):
    g, sd, m, plan = case
    x = T(g['hankel_in_direct'])[None]
    assert rel_l2(N(plan.hankel(x))[0], g['hankel_fwd_direct']) < 1e-12
    assert rel_l2(N(plan.hankel(x, inverse=True))[0], g['hankel_inv_direct']) < 1e-12


def test_ft_golden_batched(case):
    g, sd, m, plan = case
    b3 = np.stack([g['sht_inverse_of_forward']] * 3)
    b3[1] *= 2.0
    f = N(plan.ft(T(b3)))
    assert rel_l2(f[0], g['ft_x']) < 1e-12 and rel_l2(f[1], 2 * g['ft_x']) < 1e-12 and rel_l2(f[2], g['ft_x']) < 1e-12
    assert rel_l2(N(plan.ift(T(b3)))[2], g['ift_x']) < 1e-12


def test_projection_golden(case):
    g, sd, m, plan = case
    ip = N(plan.project_invariants(T(np.stack([g['I_direct'], g['I_direct']]))))
    assert np.array_equal(ip[0], ip[1])                      # deterministic across the batch
    assert rel_l2(ip[1], g['Iproj_direct']) < 1e-6
    splits = np.arange(1, m.l_max + 1) ** 2
    got, ref = np.split(ip[0], splits, axis=1), np.split(g['Iproj_direct'], splits, axis=1)
    assert rel_l2(got[0], ref[0]) < 1e-15 and rel_l2(got[2], ref[2]) < 1e-12
    for l in range(1, m.l_max + 1, 2):
        assert not got[l].any()                              # odd orders are zeroed on the masked region
    # B_l = I'_l I'_l^H reproduces the measured invariants on the resolved subspace (fxs_invariant_tools.py:915-923)
    for l in (0, 2):
        assert rel_l2(got[l] @ got[l].conj().T, m.rp.deg2_invariants[l]) < 1e-8


def test_plan_can_be_retargeted_to_new_invariants(case):
    """xfb_plan_set_projection twice on one plan: the second set of constants replaces the first."""
    g, sd, m, plan = case
    from xframe_b200.plan import Plan
    p2 = Plan(m.l_max, len(m.rs), float(g['max_q']), n_theta=int(g['n_theta']), n_phi=int(g['n_phi']), max_batch=2)
    try:
        rng = np.random.default_rng(8)
        other = [rng.standard_normal(np.asarray(v).shape) for v in m.rp.projection_matrices]
        p2.set_projection(other, m.rp.radial_mask, 2.0)
        x = T(np.stack([g['I_direct'], g['I_direct']]))
        first = N(p2.project_invariants(x))
        p2.set_projection(m.rp.projection_matrices, m.rp.radial_mask, m.rp.number_of_particles[0])
        second = N(p2.project_invariants(x))
        assert rel_l2(first[0], g['Iproj_direct']) > 1e-3
        assert np.array_equal(second, N(plan.project_invariants(x)))
    finally:
        p2.close()


def test_elementwise_golden(case):
    g, sd, m, plan = case
    from xframe_b200.plan import HIO, ER
    assert rel_l2(N(plan.modify_intensity(T(g['rho_hat0'])[None], T(g['I_proj_grid'])[None]))[0], g['rho_hat_mod']) < 1e-14
    sup = torch.ones((1,) + plan.grid_shape, dtype=torch.uint8, device='cuda')
    nxt, err = plan.real_update(HIO, float(g['hio_beta']), T(g['rho_new'])[None], T(g['rho0'])[None], sup)
    assert rel_l2(N(nxt)[0], g['hio_out']) < 1e-14
    e = N(err)[0]
    assert abs(e[0] / e[1] - float(g['real_err'])) < 1e-11 * float(g['real_err'])
    nxt, _ = plan.real_update(ER, 0.0, T(g['rho_new'])[None], T(g['rho0'])[None], sup)
    assert np.array_equal(N(nxt)[0], g['rho_proj'])
    # ft_stab combine: rho_new = ift + (prev - rt) for radial index >= 1 (misk.py:325-329)
    rt = g['rho0'] * 0.25
    nxt2, _ = plan.real_update(ER, 0.0, T(g['rho_new'])[None], T(g['rho0'])[None], sup, rho_rt=T(rt)[None])
    comb = O.add_above_zero_index(g['rho_new'], g['rho0'] - rt)
    ref = m.real_pr.projection(comb.copy())[0]
    assert rel_l2(N(nxt2)[0], ref) < 1e-14
    sw = N(plan.shrinkwrap(T(g['rho0'])[None], 12.5, 0.09))[0]
    assert (sw != g['sw_mask']).mean() < 1e-3


def test_real_projection_chain_like_the_reference(case):
    """assemble_projection (fxs_Projections.py:110-130): names without a generate_<name>_projection are logged and ignored
    (the reference's default list carries 'assert_real'), average_center (:96-110) replaces the first shells by their angular
    mean at its position in the chain and never marks a point as changed."""
    g, sd, m, plan = case
    import copy
    from xframe_b200.plan import Plan, HIO, ER
    popt = copy.deepcopy(sd['projections']['real']['projections'])
    popt['apply'] = ['support', 'assert_real', 'average_center', 'value_threshold', 'limit_imag']
    popt['average_center'] = {'max_radial_id': 2}
    rp = O.RealProjection(popt, m.real_grid)
    p2 = Plan(m.l_max, len(m.rs), float(g['max_q']), n_theta=int(g['n_theta']), n_phi=int(g['n_phi']), max_batch=2)
    try:
        p2.set_real(popt['apply'], rp.initial_support, popt['value_threshold']['threshold'], popt['limit_imag']['threshold'],
                    average_center_shells=2)
        assert p2.real_projections == ('support', 'average_center', 'value_threshold', 'limit_imag')
        rng = np.random.default_rng(9)
        x = np.stack([g['rho_new'], g['rho_new'] * (1 + 0.3 * rng.standard_normal(g['rho_new'].shape))])
        prev = np.stack([g['rho0'], g['rho0']])
        sup = torch.ones((2,) + p2.grid_shape, dtype=torch.uint8, device='cuda')
        nxt, err = p2.real_update(ER, 0.0, T(x), T(prev), sup)
        for b in range(2):
            want, masks = rp.projection(x[b].copy())
            assert rel_l2(N(nxt)[b], want) < 1e-14
            assert rel_l2(N(nxt)[b][:2], want[:2]) < 1e-13            # the averaged shells
            hio = O.hybrid_input_output(0.4, x[b].copy(), [want, masks], prev[b])
            got, _ = p2.real_update(HIO, 0.4, T(x), T(prev), sup)
            assert rel_l2(N(got)[b], hio) < 1e-14
    finally:
        p2.close()


def test_unknowns_of_a_zeroed_order_keep_the_column_count():
    """xfb_get_unknowns for an order whose V_l was zeroed (odd_orders_to_0): the reference's svd of the zero [n_cols x 2l+1]
    matrix gives identity factors with n_cols rows -- also when n_cols < min(N_r, 2l+1)."""
    from xframe_b200.plan import Plan, HIO
    l_max, n_r = 5, 16
    plan = Plan(l_max, n_r, 0.2, n_theta=8, n_phi=16, max_batch=1)
    try:
        rng = np.random.default_rng(2)
        ncols = [1, 3, 5, 4, 9, 6]                       # orders 3 and 5: fewer columns than 2l+1 < N_r
        pm = [rng.standard_normal((n_r, c)) for c in ncols]
        for l in (1, 3, 5):
            pm[l][:] = 0
        plan.set_projection(pm, True, 1.0)
        plan.set_real(['support'], np.ones(plan.grid_shape, bool))
        c = rng.standard_normal((1, n_r, (l_max + 1) ** 2)) + 1j * rng.standard_normal((1, n_r, (l_max + 1) ** 2))
        plan.project_invariants(T(c))
        unk = plan.unknowns(0)
        for l in (3, 5):
            assert unk[l].shape == (ncols[l], 2 * l + 1)
            assert np.array_equal(unk[l], np.eye(ncols[l], 2 * l + 1).astype(complex))
    finally:
        plan.close()


def test_deg2_invariants_and_l2_diff_on_the_device(case):
    """B_l = I_l I_l^H (fxs_invariant_tools.py:915-923) and deg2_invariant_l2_diff (fxs_IO_methods.py:412-447) for the coefficients
    of real fields, against the oracle restatement (pinned to the reference in tests/test_under_reference.py).  Tolerance 1e-12."""
    g, sd, m, plan = case
    rng = np.random.default_rng(21)
    x = rng.standard_normal((2,) + plan.grid_shape) ** 2 + 0j                    # real, non-negative like |rho_hat|^2
    I = plan.sht_forward(T(x))
    got = N(plan.deg2_invariants(I))
    In = N(I)
    for b in range(2):
        Il = [In[b][:, l * l:(l + 1) * (l + 1)] for l in range(m.l_max + 1)]
        want = O.harmonic_coeff_to_deg2_invariants_3d(Il)
        assert np.abs(want.imag).max() < 1e-12 * np.abs(want).max()
        assert rel_l2(got[b], want.real) < 1e-12
    ref = O.harmonic_coeff_to_deg2_invariants_3d(m.rp.projection_matrices)
    rmask = np.broadcast_to(np.asarray(m.rp.radial_mask, bool), (m.l_max + 1, len(m.qs))).copy()
    rmask[:, :1] = False                                                          # exercise the invariant mask
    plan.set_deg2_reference(ref, rmask, 3.0)
    err = N(plan.deg2_invariant_diff(I))
    for b in range(2):
        Il = [In[b][:, l * l:(l + 1) * (l + 1)] for l in range(m.l_max + 1)]
        want = O.deg2_invariant_l2_diff(ref, rmask, 3.0, Il)
        assert np.array_equal(want == -1, err[b] == -1)                           # odd orders: zero reference
        ok = want != -1
        assert np.allclose(err[b][ok], want[ok], rtol=1e-11, atol=0)


def test_worker_reports_the_deg2_metric_per_iteration():
    """settings main_loop.error.methods.reciprocal.calculate: [deg2_invariant_l2_diff] -> error_dict['reciprocal'][...] [n_it, n_orders]
    (reconstruct.py:526, arrayfy_error_dict); the first row is the metric of the initial density's intensity coefficients."""
    import copy
    from xframe_b200.worker import ProjectWorker
    g = load_golden('ref_small_ftstab')
    sd = copy.deepcopy(golden_settings(g))
    sd['GPU'] = {'use': True, 'n_gpu_workers': 1}
    sd['main_loop']['error']['methods']['reciprocal'] = {'calculate': ['deg2_invariant_l2_diff']}
    w = ProjectWorker(sd, golden_data(g), n_reconstructions=2, initial_densities=[g['rho0'], g['rho0'] * 1.1])
    res, _ = w.run()
    plan = w.plan
    for k, r in enumerate(res):
        h = r['error_dict']['reciprocal']['deg2_invariant_l2_diff']
        assert h.shape == (len(r['error_dict']['main']), int(g['l_max']) + 1) and np.isfinite(h).all()
        rho0 = T(r['initial_density'])[None]
        fd = plan.ft(rho0)
        I = plan.sht_forward((fd * fd.conj()).real.to(torch.complex128).contiguous())
        assert np.allclose(N(plan.deg2_invariant_diff(I))[0][:h.shape[1]], h[0], rtol=1e-9, atol=1e-300)
    assert rel_l2(res[0]['error_dict']['main'], g['loop_main_error']) < 1e-6
    w.plan.close()


def test_modified_intensity_corner_cases(case):
    """project_to_modified_intensity (fxs_Projections.py:899-909) on crafted points: negative / zero / tiny / huge projected
    intensities and vanishing rho_hat -- zeros, infs and NaNs land exactly where numpy puts them (the device multiplier is
    two rsqrt sequences instead of a division and a square root; only SUBNORMAL operands are treated as zero)."""
    g, sd, m, plan = case
    rng = np.random.default_rng(5)
    rh = (rng.normal(size=plan.grid_shape) + 1j * rng.normal(size=plan.grid_shape))
    ip = rng.random(plan.grid_shape) * 3.0 + 0j
    flat_r, flat_i = rh.reshape(-1), ip.reshape(-1)
    flat_r[0:4] = [0.0, 0.0, 1e-160 + 0j, 3.0 - 4.0j]          # |rho_hat|^2 = 0, 0, underflow to 0, 25
    flat_i[0:4] = [2.0, 0.0, 1.0, -1.0]                        # -> NaN (0 * inf), NaN, inf / NaN components, 0 (masked)
    flat_r[4:8] = [1e-120 + 1e-120j, 1e120 - 1e120j, 2.0 + 0j, -1.0 + 1.0j]
    flat_i[4:8] = [1e-250, 1e250, 0.0, 1e-300]                 # tiny / huge ratios, zero intensity, tiny normal intensity
    with np.errstate(all='ignore'):
        ref = m.rp.project_to_modified_intensity(rh.copy(), (rh * rh.conj()).real.copy(), ip.copy())
    got = N(plan.modify_intensity(T(rh)[None], T(ip)[None]))[0]
    r, q = ref.reshape(-1), got.reshape(-1)
    for part in (np.real, np.imag):
        a, b = part(r), part(q)
        assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.isinf(a), np.isinf(b))
        assert np.array_equal(a == 0.0, b == 0.0)
        fin = np.isfinite(a)
        assert np.allclose(a[fin], b[fin], rtol=4e-15, atol=0.0)


@pytest.mark.parametrize('tag', ['ref_small_ftstab', 'ref_small_plain'])
def test_full_loop_against_reference_golden(tag):
    from xframe_b200.plan import Plan
    from xframe_b200.reconstruct import run_schedule
    g = load_golden(tag)
    sd = golden_settings(g)
    m = O.MTIP(sd, golden_data(g))
    plan = Plan(m.l_max, len(m.rs), float(g['max_q']), n_theta=int(g['n_theta']), n_phi=int(g['n_phi']), max_batch=2)
    plan.set_projection(m.rp.projection_matrices, m.rp.radial_mask, m.rp.number_of_particles[0])
    popt = sd['projections']['real']['projections']
    plan.set_real(popt['apply'], m.real_pr.initial_support, popt['value_threshold']['threshold'], popt['limit_imag']['threshold'])
    rho0 = np.stack([g['rho0'], g['rho0'] * 1.0])
    res = run_schedule(plan, sd, T(rho0))
    assert res['loop_iterations'] == int(g['loop_iterations'])
    for b in range(2):
        assert rel_l2(res['errors'][b], g['loop_main_error']) < 1e-6
        assert abs(res['best_error'][b] - float(g['loop_final_error'])) < 1e-6 * float(g['loop_final_error'])
        assert rel_l2(res['initial_density'][b], g['loop_initial_density']) < 1e-12
        assert rel_l2(res['last_real'][b], g['loop_last_real_density']) < 1e-6
        assert rel_l2(res['best_real'][b], g['loop_real_density']) < 1e-6
        assert rel_l2(res['last_reciprocal'][b], g['loop_last_reciprocal_density']) < 1e-6
        assert rel_l2(res['best_reciprocal'][b], g['loop_reciprocal_density']) < 1e-6
        assert (res['last_support'][b] != g['loop_last_support_mask']).mean() < 1e-3
        assert (res['best_support'][b] != g['loop_support_mask']).mean() < 1e-3
    plan.close()


def test_fused_ft_stab_equals_literal_sketch():
    """Default: IFT(rho_hat') + (rho - IFT(rho_hat)) evaluated as IFT(rho_hat' - rho_hat) + rho (linearity).  The literal
    two-transform form of the sketch (reconstruct.py:584-593) must give the same iterates up to rounding."""
    from xframe_b200.plan import Plan
    from xframe_b200.reconstruct import run_schedule
    g = load_golden('ref_small_ftstab')
    sd = golden_settings(g)
    m = O.MTIP(sd, golden_data(g))
    out = []
    for fused in (True, False):
        plan = Plan(m.l_max, len(m.rs), float(g['max_q']), n_theta=int(g['n_theta']), n_phi=int(g['n_phi']), max_batch=1)
        plan.set_projection(m.rp.projection_matrices, m.rp.radial_mask, m.rp.number_of_particles[0])
        popt = sd['projections']['real']['projections']
        plan.set_real(popt['apply'], m.real_pr.initial_support, popt['value_threshold']['threshold'], popt['limit_imag']['threshold'])
        plan.set_fused_ft_stab(fused)
        out.append(run_schedule(plan, sd, T(g['rho0'])[None]))
        plan.close()
    assert rel_l2(out[0]['errors'], out[1]['errors']) < 1e-9
    assert rel_l2(out[0]['last_real'], out[1]['last_real']) < 1e-9
    for res in out:
        assert rel_l2(res['errors'][0], g['loop_main_error']) < 1e-6
        assert rel_l2(res['last_real'][0], g['loop_last_real_density']) < 1e-6


def test_host_buffer_step_matches_device_loop():
    from xframe_b200.plan import Plan, HIO
    g = load_golden('ref_small_ftstab')
    sd = golden_settings(g)
    m = O.MTIP(sd, golden_data(g))
    plan = Plan(m.l_max, len(m.rs), float(g['max_q']), n_theta=int(g['n_theta']), n_phi=int(g['n_phi']), max_batch=2)
    plan.set_projection(m.rp.projection_matrices, m.rp.radial_mask, m.rp.number_of_particles[0])
    popt = sd['projections']['real']['projections']
    plan.set_real(popt['apply'], m.real_pr.initial_support, popt['value_threshold']['threshold'], popt['limit_imag']['threshold'])
    plan.mtip_init(T(np.stack([g['rho0'], g['rho0']])))
    start = plan.mtip_grid('last_real').cpu()
    plan.mtip_iterate(HIO, True, [0.5])
    ref = N(plan.mtip_grid('last_real'))
    h_in, h_out = start.clone().pin_memory(), torch.empty_like(start).pin_memory()
    h_err = torch.empty((2, 2), dtype=torch.float64).pin_memory()
    plan.mtip_step_host(HIO, True, 0.5, h_in, h_out, h_err)
    assert np.array_equal(h_out.numpy(), ref)
    assert torch.isfinite(h_err).all()
    # chunked pipeline (copy-in | compute | copy-out streams): one run per chunk must give the same bits
    plan.set_host_chunk(1)
    h_out2, h_err2 = torch.empty_like(start).pin_memory(), torch.empty((2, 2), dtype=torch.float64).pin_memory()
    plan.mtip_step_host(HIO, True, 0.5, h_in, h_out2, h_err2)
    assert np.array_equal(h_out2.numpy(), ref)
    assert np.array_equal(h_err2.numpy(), h_err.numpy())
    plan.close()


# ----------------------------------------------------------------------------------------------
# BASELINE.json full size (L=63, N_r=128, 64x128): oracle where it finishes in seconds, properties otherwise
# ----------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def full():
    from xframe_b200.plan import Plan
    plan = Plan(63, 128, 0.322416, n_theta=64, n_phi=128, max_batch=4)
    yield plan
    plan.close()


def test_full_size_transforms_against_oracle(full):
    plan = full
    from oracle.sht import sh
    s = sh(63, n_theta=64, n_phi=128)
    rng = np.random.default_rng(1234)
    c = rng.normal(size=(128, 64 ** 2)) + 1j * rng.normal(size=(128, 64 ** 2))
    x = s.inverse_d(c)                                           # band-limited synthesis
    assert rel_l2(N(plan.sht_inverse(T(c))), x) < 1e-12
    assert rel_l2(N(plan.sht_forward(T(x))), c) < 1e-11
    w = O.hankel_weights(63, 128, 2.0, 'midpoint')
    ft, ift = O.generate_ft(s, w, plan.rs.max(), 2.0, 63, 'midpoint', 'direct')
    assert rel_l2(N(plan.ft(T(x)[None]))[0], ft(x)) < 1e-11
    assert rel_l2(N(plan.ift(T(x)[None]))[0], ift(x)) < 1e-11
    zht, izht = O.generate_spherical_ht_direct(O.assemble_weights(w, plan.rs.max(), 2.0), 63)
    assert rel_l2(N(plan.hankel(T(c)[None]))[0], zht(c)) < 1e-12
    assert rel_l2(N(plan.hankel(T(c)[None], inverse=True))[0], izht(c)) < 1e-12


@pytest.mark.parametrize('L,n_theta,n_phi', [(31, 32, 64), (100, 112, 256), (20, 24, 64), (63, 64, 256)])
def test_sht_other_grid_sizes_against_oracle(L, n_theta, n_phi):
    """Covers the register FFT variants (64 = 8x8, 256 = 16x16) and the generic fallback (n_theta not a multiple of the
    CTA's theta block), incl. an anti-aliased grid (n_phi > 2L+2)."""
    from xframe_b200.plan import Plan
    from oracle.sht import sh
    n_r = 8
    plan = Plan(L, n_r, 0.05, n_theta=n_theta, n_phi=n_phi, max_batch=2)
    s = sh(L, n_theta=n_theta, n_phi=n_phi)
    rng = np.random.default_rng(L)
    c = rng.normal(size=(2 * n_r, (L + 1) ** 2)) + 1j * rng.normal(size=(2 * n_r, (L + 1) ** 2))
    x = s.inverse_d(c)
    assert rel_l2(N(plan.sht_inverse(T(c))), x) < 1e-12
    assert rel_l2(N(plan.sht_forward(T(x))), c) < 1e-11
    xg = rng.normal(size=x.shape) + 1j * rng.normal(size=x.shape)          # not band limited: exercises all m > L columns
    assert rel_l2(N(plan.sht_forward(T(xg))), s.forward_d(xg)) < 1e-11
    plan.close()


def test_full_size_properties(full):
    plan = full
    rng = np.random.default_rng(5)
    c = T(rng.normal(size=(2, 128, 64 ** 2)) + 1j * rng.normal(size=(2, 128, 64 ** 2)))
    x = plan.sht_inverse(c)
    assert rel_l2(N(plan.sht_forward(x)), N(c)) < 1e-11                          # analysis o synthesis = id (band-limited)
    a, b = x[0:1].contiguous(), x[1:2].contiguous()
    lin = plan.ft((2.0 * a + (0.5 - 1.5j) * b).contiguous())
    assert rel_l2(N(lin), N(2.0 * plan.ft(a) + (0.5 - 1.5j) * plan.ft(b))) < 1e-12   # linearity
    xr = plan.sht_inverse(c).real.to(torch.complex128).contiguous()
    cr = N(plan.sht_forward(xr))[0]
    for l, mm in [(1, 1), (5, 3), (40, 17), (63, 63)]:                           # Hermitian symmetry of real fields
        assert np.abs(cr[:, l * (l + 1) - mm] - (-1) ** mm * np.conj(cr[:, l * (l + 1) + mm])).max() < 1e-11 * np.abs(cr).max()


def test_full_size_projection_and_iteration_against_oracle(full):
    """L=63 / N_r=128 six-sphere inputs: projection step and one whole HIO_ft_stab iteration vs the oracle."""
    plan = full
    from xframe_b200 import setup_host as S
    from xframe_b200.settings import tutorial_settings
    from xframe_b200.plan import HIO
    sd = tutorial_settings(grid={'max_q': 0.322416, 'max_order': 63, 'n_phi': 128, 'n_theta': 64, 'n_radial_points': 128})
    data = S.invariants_from_density(plan, S.six_sphere_density(plan))
    m = O.MTIP(sd, data)
    # GPU-made invariants equal the oracle's own (same model, same transforms)
    ref_data = O.invariants_from_density(O.six_sphere_density(m.real_grid), m.ft, m.sh, m.qs)
    assert rel_l2(data['average_intensity'], ref_data['average_intensity']) < 1e-10
    ps = S.ProjectionSetup(plan.qs, data, 63, sd['projections']['reciprocal'])
    ps.apply_to(plan)
    popt = sd['projections']['real']['projections']
    plan.set_real(popt['apply'], S.initial_support(plan, popt['support']['initial_support']), popt['value_threshold']['threshold'],
                  popt['limit_imag']['threshold'])
    rho0 = m.density_guess(np.random.default_rng(1000))
    rho_hat = m.ft(rho0)
    I = m.sh.forward_l(O.square_grid(rho_hat))
    Ip = m.rp.mtip_projection(I, m.rp.approximate_unknowns(I))
    got_t = plan.project_invariants(T(np.concatenate(I, axis=1))[None])
    got = N(got_t)[0]
    assert rel_l2(got, np.concatenate(Ip, axis=1)) < 1e-6
    # Separating "degenerate subspace" from "error" at full size: I'_l I'_l^H = V_l (P P^H) V_l^H does not depend on which orthonormal
    # completion the SVD picks inside numerically degenerate subspaces, and the dropped directions (sigma < 1e-15 sigma_max) enter it
    # quadratically -- so it must equal the measured invariants V_l V_l^H to rounding on every projected order (tolerance 1e-10, the
    # FP64 bar of BASELINE.json configs[3]), for the device result AND for the reference formula.
    Bp = N(plan.deg2_invariants(got_t))[0]
    for l in range(0, 64, 2):
        V = np.asarray(m.rp.projection_matrices[l]).real
        want = V @ V.T if l else (V @ V.T) / m.rp.number_of_particles[0]
        assert rel_l2(Bp[l], want) < 1e-10, (l, rel_l2(Bp[l], want))
        ref_b = (Ip[l] @ Ip[l].conj().T).real
        assert rel_l2(ref_b, want) < 1e-10, (l, 'reference formula')
    # one full iteration from the same state
    m.results['errors'] = {'real': {'l2_projection_diff': []}, 'reciprocal': {}, 'main': []}
    m.beta = 0.5
    rho = m.ift(rho_hat)
    plan.mtip_init(T(rho0)[None])
    assert rel_l2(N(plan.mtip_grid('last_real'))[0], rho) < 1e-11
    rh_new, rho_next = m.io_step('HIO', rho, True)
    plan.mtip_iterate(HIO, True, [0.5])
    assert rel_l2(N(plan.mtip_grid('last_real'))[0], rho_next) < 1e-6
    assert rel_l2(N(plan.mtip_grid('last_reciprocal'))[0], rh_new) < 1e-6
    hist, best = plan.mtip_errors()
    e_ref = m.results['errors']['real']['l2_projection_diff'][-1]
    assert abs(float(hist[0, 0]) - e_ref) < 1e-6 * e_ref


def test_worker_result_schema():
    """Result-dict keys / dtypes / shapes as the reference's integration test asserts them
    (tests/test_fxs_integration.py:388-421), on the reference test's own tiny grid sizes."""
    from xframe_b200.worker import ProjectWorker
    g = load_golden('ref_small_ftstab')
    sd = golden_settings(g)
    sd['GPU'] = {'use': True, 'batch': 2, 'seed': 11}
    w = ProjectWorker(sd, golden_data(g), n_reconstructions=3)
    res, _ = w.run()
    assert len(res) == 3
    shape = (int(g['n_r']), int(g['n_theta']), int(g['n_phi']))
    n_it = len(g['loop_main_error'])
    for r in res:
        for k in ['real_density', 'last_real_density', 'reciprocal_density', 'last_reciprocal_density', 'initial_density']:
            assert r[k].dtype == np.complex128 and r[k].shape == shape and np.isfinite(r[k]).all()
        for k in ['support_mask', 'last_support_mask', 'initial_support']:
            assert r[k].dtype == bool and r[k].shape == shape
        assert r['error_dict']['main'].shape == (n_it,) and r['error_dict']['real']['l2_projection_diff'].shape == (n_it,)
        assert r['loop_iterations'] == int(g['loop_iterations'])
        assert r['last_deg2_invariant'].shape == (int(g['l_max']) + 1, shape[0], shape[0])
        assert r['final_error'] == r['error_dict']['main'].min()
        assert len(r['projection_matrices']) == int(g['l_max']) + 1
    assert res[0]['final_error'] != res[1]['final_error']       # different seeds -> different runs


def test_full_tutorial_reconstruction_statistics():
    """Whole tutorial schedule (600 iterations + 6 SW, L=63, N_r=128) for the seeds the oracle was run with
    (tests/golden/full_run_oracle.json, made by tools/oracle_full_run.py).  HIO is chaotic (differences grow ~2.5x per
    iteration), so trajectories are compared where they are comparable: the first iteration tightly, the final
    real-space error and the support size statistically."""
    import json, os
    from helpers import GOLDEN
    from xframe_b200 import setup_host as S
    from xframe_b200.plan import Plan
    from xframe_b200.settings import tutorial_settings
    from xframe_b200.worker import ProjectWorker
    ref = json.load(open(os.path.join(GOLDEN, 'full_run_oracle.json')))['runs']
    seeds = [r['seed'] for r in ref]
    sd = tutorial_settings(grid={'max_q': 0.322416, 'max_order': 63, 'n_phi': 128, 'n_theta': 64, 'n_radial_points': 128})
    sd['GPU'] = {'use': True, 'batch': len(seeds), 'seed': None}
    boot = Plan(63, 128, 0.322416, n_theta=64, n_phi=128, max_batch=1)
    data = S.invariants_from_density(boot, S.six_sphere_density(boot))
    boot.close()
    w = ProjectWorker(sd, data, n_reconstructions=len(seeds), seeds=seeds)
    res, _ = w.run()
    ref_final = np.median([r['final_error'] for r in ref])
    ref_last = np.median([r['last_error'] for r in ref])
    for r, o in zip(res, ref):
        e = r['error_dict']['main']
        assert len(e) == o['n_errors'] == 600
        assert abs(e[0] - o['errors_every_20'][0]) < 1e-6 * o['errors_every_20'][0]      # same guess, same first iteration
        assert np.isfinite(r['real_density']).all()
        assert ref_final / 4 < r['final_error'] < ref_final * 4
        assert ref_last / 3 < e[-1] < ref_last * 3
        assert 0.03 < r['last_support_mask'].mean() < 0.25
    w.plan.close()


@pytest.mark.parametrize('l_max,n_r', [(15, 24), (70, 40), (20, 160), (66, 136)])
def test_projection_random_matrices_all_kernel_variants(l_max, n_r):
    """Invariant projection on random, well-conditioned V_l against the reference formula I'_l = V_l U V^H,
    U S V^H = svd(V_l^H D^2 I_l) (fxs_Projections.py:761-767,835-841).  The sizes select the Jacobi kernel variants:
    column length <= 128 / N_r <= 128 (512 threads), column length > 128, N_r > 128, both (256 threads, G in global)."""
    from xframe_b200.plan import Plan
    rng = np.random.default_rng(l_max * 1000 + n_r)
    plan = Plan(l_max, n_r, 0.3, max_batch=2)
    V = [rng.standard_normal((n_r, min(n_r, 2 * l + 1))) * (1.0 if l % 2 == 0 else 0.0) for l in range(l_max + 1)]
    plan.set_projection(V, True, 1.0)
    f = torch.from_numpy(rng.standard_normal((2, n_r, plan.n_theta, plan.n_phi)) + 0j).cuda()
    I = N(plan.sht_forward(f))                                               # Hermitian-symmetric coefficients of real fields
    got = N(plan.project_invariants(T(I)))
    D2 = plan.qs ** 2
    for b in range(2):
        for l in range(0, l_max + 1):
            Il = I[b][:, l * l:(l + 1) ** 2]
            if l == 0:
                want = V[0].astype(complex)
            elif l % 2 == 1:
                want = np.zeros_like(Il)
            else:
                u, _, vh = np.linalg.svd((V[l].T * D2[None, :]) @ Il, full_matrices=False)
                want = V[l] @ (u @ vh)
            assert rel_l2(got[b][:, l * l:(l + 1) ** 2], want) < 1e-10, (b, l)
    orders, sweeps = plan.jacobi_sweeps()
    assert sweeps.max() < 40
    plan.close()


def test_config4_high_resolution_iteration_against_oracle():
    """BASELINE.json configs[3]: L=127 / N_r=256 / 128x256, max_q doubled.  Transforms <= 1e-10 per operator
    (SURVEY.md 8d), projection and one whole HIO_ft_stab iteration <= 1e-6 vs the oracle (256-thread Jacobi variant,
    column length and N_r up to 256, G in global memory for the large orders)."""
    from xframe_b200.plan import Plan, HIO
    from xframe_b200 import setup_host as S
    from xframe_b200.settings import tutorial_settings
    L, NR, NT, NP, MQ = 127, 256, 128, 256, 0.644832
    plan = Plan(L, NR, MQ, n_theta=NT, n_phi=NP, max_batch=2)
    sd = tutorial_settings(grid={'max_q': MQ, 'max_order': L, 'n_phi': NP, 'n_theta': NT, 'n_radial_points': NR})
    data = S.invariants_from_density(plan, S.six_sphere_density(plan))
    m = O.MTIP(sd, data)
    ps = S.ProjectionSetup(plan.qs, data, L, sd['projections']['reciprocal'])
    ps.apply_to(plan)
    popt = sd['projections']['real']['projections']
    plan.set_real(popt['apply'], S.initial_support(plan, popt['support']['initial_support']), popt['value_threshold']['threshold'],
                  popt['limit_imag']['threshold'])
    rho0 = m.density_guess(np.random.default_rng(1000))
    rho_hat = m.ft(rho0)
    assert rel_l2(N(plan.ft(T(rho0)[None]))[0], rho_hat) < 1e-10
    sq = O.square_grid(rho_hat)
    I = m.sh.forward_l(sq)
    assert rel_l2(N(plan.sht_forward(T(sq))), np.concatenate(I, axis=1)) < 1e-10
    Ip = m.rp.mtip_projection(I, m.rp.approximate_unknowns(I))
    got = N(plan.project_invariants(T(np.concatenate(I, axis=1))[None]))[0]
    assert rel_l2(got, np.concatenate(Ip, axis=1)) < 1e-6
    m.results['errors'] = {'real': {'l2_projection_diff': []}, 'reciprocal': {}, 'main': []}
    m.beta = 0.5
    rho = m.ift(rho_hat)
    plan.mtip_init(T(np.stack([rho0, rho0])))
    assert rel_l2(N(plan.mtip_grid('last_real'))[1], rho) < 1e-10
    rh_new, rho_next = m.io_step('HIO', rho, True)
    plan.mtip_iterate(HIO, True, [0.5])
    assert rel_l2(N(plan.mtip_grid('last_real'))[0], rho_next) < 1e-6
    assert rel_l2(N(plan.mtip_grid('last_reciprocal'))[1], rh_new) < 1e-6
    hist, best = plan.mtip_errors()
    e_ref = m.results['errors']['real']['l2_projection_diff'][-1]
    assert abs(float(hist[1, 0]) - e_ref) < 1e-6 * e_ref
    orders, sweeps = plan.jacobi_sweeps()
    assert sweeps.max() < 40
    plan.close()


def test_worker_shift_to_center_matches_reference():
    """output_density_modifiers.shift_to_center (reconstruct.py:732-738): full reference loop with the modifier on."""
    from xframe_b200.worker import ProjectWorker
    g = load_golden('ref_small_shift')
    sd = golden_settings(g)
    assert sd['output_density_modifiers']['shift_to_center']
    sd['GPU'] = {'use': True, 'batch': 2, 'seed': 3}
    w = ProjectWorker(sd, golden_data(g), n_reconstructions=2, initial_densities=[g['rho0'], g['rho0']])
    res, _ = w.run()
    for r in res:
        assert rel_l2(r['error_dict']['main'], g['loop_main_error']) < 1e-6
        assert rel_l2(r['last_real_density'], g['loop_last_real_density']) < 1e-6
        assert rel_l2(r['real_density'], g['loop_real_density']) < 1e-6
        assert rel_l2(r['last_reciprocal_density'], g['loop_last_reciprocal_density']) < 1e-6
        assert rel_l2(r['reciprocal_density'], g['loop_reciprocal_density']) < 1e-6
        assert rel_l2(r['last_deg2_invariant'], g['loop_last_deg2']) < 1e-6


def test_results_do_not_depend_on_the_batch():
    """A run gives bit-identical densities (and errors equal to rounding) whether it is iterated alone or inside a batch (L=63 / N_r=128,
    several problems per CTA in the Jacobi work queue, multi-group cp.async rings in the Legendre kernels), and every
    Procrustes problem of the batch is solved."""
    import bench
    from xframe_b200.plan import HIO
    nb = 12
    bench.select_workload('l63')
    plan, sd, rho0 = bench.build_problem(nb, 0, [1000 + i for i in range(nb)])
    plan.mtip_init(rho0)
    for _ in range(3):
        plan.mtip_iterate(HIO, True, [0.5])
    orders, sw = plan.jacobi_sweeps()
    assert sw.shape == (nb, len(orders)) and sw.min() >= 1 and sw.max() < 40
    hist, _ = plan.mtip_errors()
    big = N(plan.mtip_grid('last_real'))
    hist = N(hist)
    plan.close()
    for k in (0, 7, nb - 1):
        p1, _, r1 = bench.build_problem(1, 0, [1000 + k])
        p1.mtip_init(r1)
        for _ in range(3):
            p1.mtip_iterate(HIO, True, [0.5])
        assert np.array_equal(N(p1.mtip_grid('last_real'))[0], big[k])
        # the error integrals are summed over a batch-dependent number of blocks: equal to rounding, not bitwise
        assert np.allclose(N(p1.mtip_errors()[0])[0], hist[k], rtol=1e-12, atol=0)
        p1.close()


def test_tma_fed_hankel_gives_the_same_bits(monkeypatch):
    """hankel3_tma_kernel (operand tiles by cp.async.bulk + mbarrier) against the cp.async ring kernel: same accumulation order, so
    bit-identical outputs -- with partially filled row tiles (rows of an order not a multiple of 64), both directions."""
    from xframe_b200.plan import Plan
    rng = np.random.default_rng(17)
    out = []
    x = None
    for tma in ('1', '0'):
        monkeypatch.setenv('XFB_HANKEL_TMA', tma)
        plan = Plan(15, 128, 0.322416, n_theta=16, n_phi=32, max_batch=3)
        if x is None:
            x = T(rng.standard_normal((3, 128, 256)) + 1j * rng.standard_normal((3, 128, 256)))
        out.append((N(plan.hankel(x)), N(plan.hankel(x, inverse=True))))
        plan.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


def test_graph_replay_gives_the_same_bits(monkeypatch):
    """Iterations replayed from the captured CUDA graph (third and later iterations of a kind; beta, history column and sub-loop iteration
    read from device memory) against eager launches: identical densities, error histories, best-density bookkeeping and unknowns."""
    import bench
    from xframe_b200.plan import HIO, ER
    nb = 3
    bench.select_workload('l63')
    out = []
    for graph in ('0', '1'):
        monkeypatch.setenv('XFB_GRAPH', graph)
        plan, sd, rho0 = bench.build_problem(nb, 0, [3000 + i for i in range(nb)])
        plan.mtip_init(rho0)
        plan.mtip_set_outer_iteration(1)
        plan.mtip_iterate(HIO, True, [0.5, 0.47, 0.44, 0.41, 0.38])
        plan.mtip_shrinkwrap(20.0, 0.09, 6e-3)
        plan.mtip_set_outer_iteration(2)
        plan.mtip_iterate(ER, True, [0.0] * 4)
        plan.mtip_iterate(HIO, False, [0.3, 0.3, 0.3])
        unk = plan.unknowns(1)
        launches = plan.launch_count()
        assert plan.graph_replays() == (0 if graph == '0' else 4 + 3 + 2)      # per kind of iteration: the first runs eagerly, the others are replays
        out.append((N(plan.mtip_grid('last_real')), N(plan.mtip_grid('best_real')), N(plan.mtip_errors()[0]), N(plan.mtip_errors()[1]),
                    N(plan.mtip_grid('best_support')), unk, launches))
        plan.close()
    a, b = out
    for i in range(5):
        assert np.array_equal(a[i], b[i]), i
    for u, v in zip(a[5], b[5]):
        assert np.array_equal(u, v)
    assert abs(a[6] - b[6]) <= 12          # the launch count is kept honest for replayed iterations (+1 parameter kernel per replay)


def test_two_stream_halves_give_the_same_bits():
    """xfb_mtip_iterate with the batch cut into two halves on two streams (the second half one projection behind, Jacobi launches on a
    share of the SMs) against the single-stream path: identical densities, error histories and unknowns for every run."""
    import bench
    from xframe_b200.plan import HIO, ER
    nb = 6
    bench.select_workload('l63')
    out = []
    for dual in (False, True):
        plan, sd, rho0 = bench.build_problem(nb, 0, [2000 + i for i in range(nb)])
        plan.set_dual_stream(dual, min_batch=2, big_sms=40, small_sms=16)
        plan.mtip_init(rho0)
        plan.mtip_iterate(HIO, True, [0.5, 0.45, 0.4])
        plan.mtip_shrinkwrap(20.0, 0.09, 6e-3)
        plan.mtip_iterate(ER, True, [0.0, 0.0])
        unk = plan.unknowns(nb - 1)
        out.append((N(plan.mtip_grid('last_real')), N(plan.mtip_grid('best_reciprocal')), N(plan.mtip_errors()[0]), N(plan.mtip_grid('last_support')), unk))
        plan.close()
    a, b = out
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[3], b[3])
    assert np.allclose(a[2], b[2], rtol=1e-12, atol=0)           # error integrals: block count per run depends on the launch batch
    for u, v in zip(a[4], b[4]):
        assert np.array_equal(u, v)


@pytest.mark.parametrize('l_max,n_r,n_theta,n_phi,ft_type', [(10, 33, 16, 32, 'midpoint'), (6, 20, 8, 16, 'trapz'), (21, 48, 24, 64, 'midpoint'),
                                                              (8, 24, 16, 32, 'gauss')])
def test_ragged_sizes_iterations_against_oracle(l_max, n_r, n_theta, n_phi, ft_type):
    """Odd / small / non-power-of-two radial sizes and the trapz / gauss radial rules: HIO_ft_stab, shrink wrap and plain ER
    iterations of a batch of 3 against the oracle (odd N_r takes the non-cp.async Hankel kernel, n_phi = 32 / 16 the generic
    Stockham FFT without the fused pointwise variants, n_theta = 24 the 8-row FFT tiles)."""
    from xframe_b200.plan import Plan, HIO, ER
    from xframe_b200 import setup_host as S
    from xframe_b200.settings import tutorial_settings
    max_q = 2.0 * n_r / 794.0
    sd = tutorial_settings(grid={'max_q': max_q, 'max_order': l_max, 'n_phi': n_phi, 'n_theta': n_theta, 'n_radial_points': n_r},
                           fourier_transform={'type': ft_type, 'reciprocity_coefficient': 2.0})
    sd['projections']['reciprocal']['used_order_ids'] = np.arange(l_max + 1)
    plan = Plan(l_max, n_r, max_q, n_theta=n_theta, n_phi=n_phi, ft_type=ft_type, max_batch=3)
    data = S.invariants_from_density(plan, S.six_sphere_density(plan))
    m = O.MTIP(sd, data)
    S.ProjectionSetup(plan.qs, data, l_max, sd['projections']['reciprocal']).apply_to(plan)
    popt = sd['projections']['real']['projections']
    plan.set_real(popt['apply'], S.initial_support(plan, popt['support']['initial_support']), popt['value_threshold']['threshold'],
                  popt['limit_imag']['threshold'])
    rho0 = np.stack([m.density_guess(np.random.default_rng(5 + i)) for i in range(3)])
    plan.mtip_init(T(rho0))
    m.results['errors'] = {'real': {'l2_projection_diff': []}, 'reciprocal': {}, 'main': []}
    rho = m.ift(m.ft(rho0[1]))
    for it, (name, code, stab, beta) in enumerate([('HIO', HIO, True, 0.5), ('HIO', HIO, False, 0.45), ('ER', ER, True, 0.0)]):
        m.beta = beta
        rh, rho = m.io_step(name, rho, stab)
        plan.mtip_iterate(code, stab, [beta])
        assert rel_l2(N(plan.mtip_grid('last_real'))[1], rho) < 1e-7, (it, name)
        assert rel_l2(N(plan.mtip_grid('last_reciprocal'))[1], rh) < 1e-7, (it, name)
    hist, _ = plan.mtip_errors()
    assert np.allclose(N(hist)[1], m.results['errors']['real']['l2_projection_diff'], rtol=1e-7, atol=0)
    m.sw.set_sigma(15.0)
    m.sw.set_threshold(0.09)
    mask = m.shrink_wrap(rho)
    plan.mtip_shrinkwrap(15.0, 0.09, 6e-3)
    assert (N(plan.mtip_grid('last_support'))[1] != mask).mean() < 5e-3      # threshold ties may flip single voxels of these tiny grids
    plan.close()
