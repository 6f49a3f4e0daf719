"""The 2-D oracle (oracle/mtip2d.py) replayed against golden vectors produced by the UNMODIFIED reference with
`dimensions: 2` (tests/golden/make_golden_2d.py -> ref2d_*.npz).  Operators <= 1e-12, loop error history <= 1e-7."""
import numpy as np
import pytest

from helpers import load_golden, rel_l2
from oracle import mtip as O
from oracle import mtip2d as O2

CASES = ['ref2d_small_ftstab', 'ref2d_medium_plain', 'ref2d_medium_so']


def settings_2d(g):
    from make_golden_2d import settings_dict_2d
    return settings_dict_2d(int(g['n_r']), int(g['m_max']), float(g['max_q']), bool(g['ft_stab']),
                            so_freedom=bool(g['so_freedom']) if 'so_freedom' in g else False)


def data_2d(g):
    return {'dimensions': 2, 'xray_wavelength': 1.23984, 'average_intensity': g['avg_intensity'], 'data_radial_points': g['data_q'],
            'data_angular_points': g['phis'], 'max_order': int(g['m_max']), 'data_projection_matrices': g['pm'], 'number_of_particles': 1}


@pytest.fixture(scope='module', params=CASES)
def case(request):
    g = load_golden(request.param)
    return g, O2.MTIP2D(settings_2d(g), data_2d(g))


def test_grids_and_weights(case):
    g, m = case
    assert np.array_equal(m.rs, g['rs']) and np.array_equal(m.qs, g['qs'])
    assert np.allclose(m.real_grid[0, :, 1], g['phis'], atol=1e-15)
    assert np.allclose(m.weights, g['hankel_weights'], rtol=1e-13, atol=1e-300)
    assert m.n_phi == int(g['n_phi']) == 2 * int(g['m_max']) + 1


def test_transforms(case):
    g, m = case
    x = g['x_grid']
    assert rel_l2(O2.cht_complex_forward(x), g['cht_complex_forward']) < 1e-14
    assert rel_l2(O2.cht_complex_inverse(x), g['cht_complex_inverse']) < 1e-14
    assert rel_l2(O2.cht_real_forward(x), g['cht_real_forward']) < 1e-14
    assert rel_l2(O2.cht_real_inverse(g['cht_real_forward'], m.n_phi), g['cht_real_inverse']) < 1e-14
    zht, izht = O2.generate_polar_ht(O2.assemble_weights_2d(m.weights, np.max(m.rs), 2.0))
    assert rel_l2(zht(g['cht_complex_forward']), g['hankel_fwd']) < 1e-13
    assert rel_l2(izht(g['cht_complex_forward']), g['hankel_inv']) < 1e-13
    assert rel_l2(m.ft(x), g['ft_x']) < 1e-12 and rel_l2(m.ift(x), g['ift_x']) < 1e-12
    assert abs(m.integrator.integrate((x * x.conj()).real) - g['integral_of_x2']) < 1e-12 * abs(g['integral_of_x2'])


def test_projection_chain(case):
    g, m = case
    assert abs(m.rp.integrated_intensity - g['integrated_intensity']) < 1e-12 * abs(g['integrated_intensity'])
    assert rel_l2(m.rp.projection_matrices, g['projection_matrices_final']) < 1e-12
    assert np.array_equal(m.rp.radial_mask, g['radial_mask'])
    rho_hat = m.ft(g['rho0'])
    assert rel_l2(rho_hat, g['rho_hat0']) < 1e-12
    sq = O.square_grid(g['rho_hat0'])
    I = O2.cht_real_forward(sq)
    assert rel_l2(I, g['I_m']) < 1e-13
    unk = m.rp.approximate_unknowns(g['I_m'])
    assert rel_l2(unk, g['unknowns']) < 1e-12
    Ip = m.rp.mtip_projection(g['I_m'], g['unknowns'])
    assert rel_l2(Ip, g['Iproj_m']) < 1e-13
    assert rel_l2(O2.cht_real_inverse(g['Iproj_m'], m.n_phi), g['I_proj_grid']) < 1e-13
    assert rel_l2(m.rp.project_to_modified_intensity(g['rho_hat0'], np.array(sq), g['I_proj_grid']), g['rho_hat_mod']) < 1e-13
    rn = m.ift(g['rho_hat_mod'])
    assert rel_l2(rn, g['rho_new']) < 1e-12
    proj = m.real_pr.projection(np.array(g['rho_new']))
    assert rel_l2(proj[0], g['rho_proj']) < 1e-15
    e = O.l2_projection_diff(m.integrator, g['rho_new'], proj, m.real_pr.initial_support)
    assert abs(e - g['real_err']) < 1e-12 * abs(g['real_err'])
    assert abs(m.sw.default_sigma - g['sw_default_sigma']) < 1e-15 * g['sw_default_sigma']
    m.sw.set_sigma(12.5)
    m.sw.set_threshold(0.09)
    assert np.array_equal(m.shrink_wrap(g['rho0']), g['sw_mask'])


def test_full_loop(case):
    g, _ = case
    m = O2.MTIP2D(settings_2d(g), data_2d(g))
    res = m.run(rho0=g['rho0'].copy())
    assert rel_l2(res['error_dict']['main'], g['loop_main_error']) < 1e-7
    assert rel_l2(res['last_real_density'], g['loop_last_real_density']) < 1e-7
    assert rel_l2(res['real_density'], g['loop_real_density']) < 1e-7
    assert rel_l2(res['last_reciprocal_density'], g['loop_last_reciprocal_density']) < 1e-7
    assert np.array_equal(res['last_support_mask'], g['loop_last_support_mask'])
    assert abs(res['final_error'] - g['loop_final_error']) < 1e-9 * abs(g['loop_final_error'])
    assert rel_l2(res['fxs_unknowns'], g['loop_unknowns']) < 1e-7
    assert rel_l2(res['last_deg2_invariant'], g['loop_last_deg2']) < 1e-7
