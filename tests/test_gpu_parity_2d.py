"""Parity of the 2-D (polar) CUDA path, through the C-ABI, with golden vectors of the UNMODIFIED reference run with
`dimensions: 2` (tests/golden/ref2d_*.npz) and with the 2-D oracle at the size of BASELINE.json configs[4]
(max_order 63 -> n_phi = 127, N_r = 128).  Relative L2: transforms / Hankel / FT <= 1e-12, projection <= 1e-12,
loop error history and densities <= 1e-6."""
import numpy as np
import pytest
import torch

from helpers import load_golden, rel_l2
from oracle import mtip as O
from oracle import mtip2d as O2
from test_oracle_golden_2d import settings_2d, data_2d, CASES

pytestmark = pytest.mark.gpu


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to('cuda')


def N(t):
    return t.cpu().numpy()


def make_plan(g, sd, max_batch=3):
    from xframe_b200.plan import Plan
    from xframe_b200 import setup_host as S
    plan = Plan(int(g['m_max']), int(g['n_r']), float(g['max_q']), max_batch=max_batch, dimensions=2)
    ps = S.ProjectionSetup2D(plan.qs, data_2d(g), plan.l_max, sd['projections']['reciprocal'])
    ps.apply_to(plan)
    popt = sd['projections']['real']['projections']
    plan.set_real(popt['apply'], S.initial_support(plan, popt['support']['initial_support']), popt['value_threshold']['threshold'],
                  popt['limit_imag']['threshold'])
    return plan, ps


@pytest.fixture(scope='module', params=CASES)
def case(request):
    g = load_golden(request.param)
    sd = settings_2d(g)
    plan, ps = make_plan(g, sd)
    yield g, sd, plan, ps
    plan.close()


def test_grids_and_setup(case):
    g, sd, plan, ps = case
    assert np.array_equal(plan.rs, g['rs']) and np.array_equal(plan.qs, g['qs']) and np.allclose(plan.phis, g['phis'], atol=1e-15)
    assert plan.grid_shape == (int(g['n_r']), int(g['n_phi']))
    assert rel_l2(ps.projection_matrices, g['projection_matrices_final']) < 1e-12
    assert np.array_equal(ps.radial_mask, g['radial_mask'])
    assert abs(ps.integrated_intensity - g['integrated_intensity']) < 1e-12 * abs(g['integrated_intensity'])


def test_circular_transforms(case):
    g, sd, plan, ps = case
    x = T(g['x_grid'])
    assert rel_l2(N(plan.sht_forward(x)), g['cht_complex_forward']) < 1e-12
    assert rel_l2(N(plan.sht_inverse(x)), g['cht_complex_inverse']) < 1e-12
    from xframe_b200.harmonic_transforms import HarmonicTransform
    opt = {'dimensions': 2, 'max_order': int(g['m_max'])}
    cht, rht = HarmonicTransform('complex', opt), HarmonicTransform('real', opt)
    assert np.allclose(cht.grid_param['phis'], g['phis'], atol=1e-15) and set(cht.transforms_by_indices) == {'m'}
    assert rel_l2(cht.forward(g['x_grid']), g['cht_complex_forward']) < 1e-12
    assert rel_l2(cht.inverse(g['x_grid']), g['cht_complex_inverse']) < 1e-12
    rf = rht.forward(g['x_grid'])
    assert rf.shape == g['cht_real_forward'].shape and rel_l2(rf, g['cht_real_forward']) < 1e-12
    ri = rht.inverse(g['cht_real_forward'])
    assert ri.dtype == np.float64 and rel_l2(ri, g['cht_real_inverse']) < 1e-12


def test_hankel_and_ft(case):
    g, sd, plan, ps = case
    c = T(g['cht_complex_forward'])[None]
    assert rel_l2(N(plan.hankel(c))[0], g['hankel_fwd']) < 1e-12
    assert rel_l2(N(plan.hankel(c, inverse=True))[0], g['hankel_inv']) < 1e-12
    b3 = np.stack([g['x_grid'], 2 * g['x_grid'], g['x_grid']])
    f = N(plan.ft(T(b3)))
    assert rel_l2(f[0], g['ft_x']) < 1e-12 and rel_l2(f[1], 2 * g['ft_x']) < 1e-12
    assert rel_l2(N(plan.ift(T(b3)))[2], g['ift_x']) < 1e-12
    # reference-facing factories (hankel_transforms.generate_ht / fourier_transforms.generate_ft with dimensions = 2)
    from xframe_b200.harmonic_transforms import HarmonicTransform
    from xframe_b200.hankel_transforms import generate_weightDict, generate_ht
    from xframe_b200.fourier_transforms import generate_ft
    M, n_r = int(g['m_max']), int(g['n_r'])
    wd = generate_weightDict(M, n_r, reciprocity_coefficient=2.0, dimensions=2, mode='midpoint')
    assert np.allclose(wd['weights'], g['hankel_weights'], rtol=1e-13, atol=1e-300)
    r_max = float(np.max(g['rs']))
    zht, izht = generate_ht(wd['weights'], wd['posHarmOrders'], r_max, reciprocity_coefficient=2.0, dimensions=2, mode='midpoint')
    assert rel_l2(zht(g['cht_complex_forward']), g['hankel_fwd']) < 1e-12 and rel_l2(izht(g['cht_complex_forward']), g['hankel_inv']) < 1e-12
    ft, ift = generate_ft(r_max, wd, HarmonicTransform('complex', {'dimensions': 2, 'max_order': M}), 2, reciprocity_coefficient=2.0,
                          mode='midpoint')
    assert rel_l2(ft(g['x_grid']), g['ft_x']) < 1e-12 and rel_l2(ift(g['x_grid']), g['ift_x']) < 1e-12


def test_projection_and_pointwise(case):
    g, sd, plan, ps = case
    M = int(g['m_max'])
    sq = O.square_grid(g['rho_hat0'])
    Ifull = plan.sht_forward(T(sq)[None])                                     # full DFT of the real intensity
    assert rel_l2(N(Ifull)[0][:, :M + 1], g['I_m']) < 1e-12
    Ip = N(plan.project_invariants(Ifull))[0]
    assert rel_l2(Ip[:, :M + 1], g['Iproj_m']) < 1e-12
    assert rel_l2(Ip[:, M + 1:], np.conj(Ip[:, 1:M + 1])[:, ::-1]) < 1e-15   # Hermitian completion for the complex inverse
    assert rel_l2(plan.unknowns(0), g['unknowns']) < 1e-12
    mod = N(plan.modify_intensity(T(g['rho_hat0'])[None], T(g['I_proj_grid'].astype(complex))[None]))[0]
    assert rel_l2(mod, g['rho_hat_mod']) < 1e-14
    sup = torch.from_numpy(plan.initial_support)[None].cuda()
    nxt, err = plan.real_update(1, 0.0, T(g['rho_new'])[None], T(g['rho0'])[None], sup)
    assert rel_l2(N(nxt)[0], g['rho_proj']) < 1e-15
    e = N(err)[0]
    assert abs(e[0] / e[1] - g['real_err']) < 1e-12 * abs(g['real_err'])       # PolarIntegrator weights
    assert np.array_equal(N(plan.shrinkwrap(T(g['rho0'])[None], 12.5, 0.09))[0], g['sw_mask'])


def test_average_center_and_ignored_names_2d(case):
    """The real projection chain in 2-D (fxs_Projections.py:96-130, dimension 2 branch: mean over the angular axis of the first shells)."""
    import copy
    g, sd, plan, ps = case
    from xframe_b200.plan import Plan, ER
    from xframe_b200 import setup_host as S
    popt = copy.deepcopy(sd['projections']['real']['projections'])
    popt['apply'] = ['average_center', 'support', 'assert_real', 'value_threshold']
    popt['average_center'] = {'max_radial_id': 3}
    p2 = Plan(int(g['m_max']), int(g['n_r']), float(g['max_q']), max_batch=2, dimensions=2)
    try:
        sup0 = S.initial_support(p2, popt['support']['initial_support'])
        p2.set_real(popt['apply'], sup0, popt['value_threshold']['threshold'], popt['limit_imag']['threshold'], average_center_shells=3)
        assert p2.real_projections == ('average_center', 'support', 'value_threshold')
        m = O2.MTIP2D(sd, data_2d(g))
        rp = O.RealProjection(popt, m.real_grid)
        rng = np.random.default_rng(4)
        x = np.stack([g['x_grid'], g['x_grid'] * (1 + 0.2 * rng.standard_normal(g['x_grid'].shape))])
        sup = torch.ones((2,) + p2.grid_shape, dtype=torch.uint8, device='cuda')
        nxt, _ = p2.real_update(ER, 0.0, T(x), T(x), sup)
        for b in range(2):
            want, _ = rp.projection(x[b].copy())
            assert rel_l2(N(nxt)[b], want) < 1e-13
    finally:
        p2.close()


def test_full_loop_against_reference(case):
    g, sd, plan, ps = case
    if 'so_freedom' in g and bool(g['so_freedom']):
        # the 2-D defaults: SO_freedom pin during phasing + shift_to_center / fix_orientation output modifiers -> through the worker
        from xframe_b200.worker import ProjectWorker
        sdw = dict(sd)
        sdw['GPU'] = {'use': True, 'batch': 2, 'seed': 1}
        res, _ = ProjectWorker(sdw, data_2d(g), n_reconstructions=2, initial_densities=[g['rho0'], g['rho0']]).run()
        for r in res:
            assert rel_l2(r['error_dict']['main'], g['loop_main_error']) < 1e-6
            assert rel_l2(r['fxs_unknowns'], g['loop_unknowns']) < 1e-6
            assert rel_l2(r['last_real_density'], g['loop_last_real_density']) < 1e-6
            assert rel_l2(r['real_density'], g['loop_real_density']) < 1e-6
            assert rel_l2(r['last_reciprocal_density'], g['loop_last_reciprocal_density']) < 1e-6
            assert rel_l2(r['last_deg2_invariant'], g['loop_last_deg2']) < 1e-6
        return
    from xframe_b200.reconstruct import run_schedule
    res = run_schedule(plan, sd, T(np.stack([g['rho0'], g['rho0']])))
    assert rel_l2(res['errors'][0], g['loop_main_error']) < 1e-6 and np.array_equal(res['errors'][0], res['errors'][1])
    assert rel_l2(res['last_real'][1], g['loop_last_real_density']) < 1e-6
    assert rel_l2(res['best_real'][0], g['loop_real_density']) < 1e-6
    assert rel_l2(res['last_reciprocal'][0], g['loop_last_reciprocal_density']) < 1e-6
    assert np.array_equal(res['last_support'][0], g['loop_last_support_mask'])
    assert abs(res['best_error'][0] - g['loop_final_error']) < 1e-8 * abs(g['loop_final_error'])
    assert rel_l2(plan.unknowns(1), g['loop_unknowns']) < 1e-6


def test_worker_2d_schema():
    from xframe_b200.worker import ProjectWorker
    g = load_golden('ref2d_small_ftstab')
    sd = settings_2d(g)
    sd['GPU'] = {'use': True, 'batch': 2, 'seed': 5}
    res, _ = ProjectWorker(sd, data_2d(g), n_reconstructions=3).run()
    shape = (int(g['n_r']), int(g['n_phi']))
    assert len(res) == 3
    for r in res:
        assert r['real_density'].shape == shape and r['real_density'].dtype == np.complex128 and np.isfinite(r['real_density']).all()
        assert r['support_mask'].shape == shape and r['support_mask'].dtype == bool
        assert r['fxs_unknowns'].shape == (int(g['m_max']) + 1,) and np.allclose(np.abs(r['fxs_unknowns']), 1.0)
        assert r['last_deg2_invariant'].shape == (int(g['m_max']) + 1, shape[0], shape[0])
        assert r['grid_pair']['real_grid'].shape == shape + (2,)


def test_config5_size_iteration_against_oracle():
    """max_order 63 (n_phi = 127), N_r = 128, six-disc model: one HIO_ft_stab and one plain ER iteration of a batch."""
    from xframe_b200.plan import Plan, HIO, ER
    from xframe_b200 import setup_host as S
    from xframe_b200.settings import tutorial_settings
    M, NR, MQ = 63, 128, 0.322416
    sd = tutorial_settings(dimensions=2, grid={'max_q': MQ, 'max_order': M, 'n_radial_points': NR})
    plan = Plan(M, NR, MQ, max_batch=8, dimensions=2)
    data = S.invariants_from_density_2d(plan, S.disk_model_density(plan))
    m = O2.MTIP2D(sd, data)
    ref = O2.invariants_from_density_2d(O2.disk_model_density(m.real_grid), m.ft, m.qs, m.real_grid[0, :, 1])
    assert rel_l2(data['data_projection_matrices'], ref['data_projection_matrices']) < 1e-10
    S.ProjectionSetup2D(plan.qs, data, M, sd['projections']['reciprocal']).apply_to(plan)
    popt = sd['projections']['real']['projections']
    plan.set_real(popt['apply'], S.initial_support(plan, popt['support']['initial_support']), popt['value_threshold']['threshold'],
                  popt['limit_imag']['threshold'])
    rho0 = np.stack([m.density_guess(np.random.default_rng(1000 + i)) for i in range(8)])
    plan.mtip_init(T(rho0))
    m.results['errors'] = {'real': {'l2_projection_diff': []}, 'reciprocal': {}, 'main': []}
    m.beta = 0.5
    rho = m.ift(m.ft(rho0[3]))
    assert rel_l2(N(plan.mtip_grid('last_real'))[3], rho) < 1e-11
    rh1, rho1 = m.io_step('HIO', rho, True)
    plan.mtip_iterate(HIO, True, [0.5])
    assert rel_l2(N(plan.mtip_grid('last_real'))[3], rho1) < 1e-9
    rh2, rho2 = m.io_step('ER', rho1, False)
    plan.mtip_iterate(ER, False, [0.0])
    assert rel_l2(N(plan.mtip_grid('last_real'))[3], rho2) < 1e-9
    assert rel_l2(N(plan.mtip_grid('last_reciprocal'))[3], rh2) < 1e-9
    hist, _ = plan.mtip_errors()
    ref_e = m.results['errors']['real']['l2_projection_diff']
    assert abs(float(hist[3, 0]) - ref_e[0]) < 1e-9 * ref_e[0] and abs(float(hist[3, 1]) - ref_e[1]) < 1e-9 * ref_e[1]
    plan.close()
