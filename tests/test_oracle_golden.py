"""Pins oracle/ against vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from helpers import load_golden, golden_settings, golden_data, rel_l2
from oracle import mtip as O

CASES = ['ref_small_ftstab', 'ref_small_plain', 'ref_medium_ops']


@pytest.fixture(scope='module', params=CASES)
def case(request):
    g = load_golden(request.param)
    m = O.MTIP(golden_settings(g), golden_data(g))
    return g, m


def test_grids(case):
    g, m = case
    assert np.array_equal(m.rs, g['rs']) and np.array_equal(m.qs, g['qs'])
    assert np.allclose(m.sh.theta, g['thetas'], rtol=0, atol=1e-15)
    assert np.allclose(m.sh.phi, g['phis'], rtol=0, atol=1e-15)


def test_hankel_weights_and_apply(case):
    g, m = case
    assert rel_l2(m.weights, g['hankel_weights']) < 1e-15
    w = O.assemble_weights(m.weights, np.max(m.rs), 2.0)
    zht, izht = O.generate_spherical_ht_direct(w, m.l_max)
    assert rel_l2(zht(g['hankel_in_direct']), g['hankel_fwd_direct']) < 1e-13
    assert rel_l2(izht(g['hankel_in_direct']), g['hankel_inv_direct']) < 1e-13


def test_sht_and_ft(case):
    g, m = case
    cl = m.sh.forward_l(g['x_grid'])
    assert rel_l2(np.concatenate(cl, axis=1), g['sht_forward_direct']) < 1e-13
    band = m.sh.inverse_l(cl)
    assert rel_l2(band, g['sht_inverse_of_forward']) < 1e-13
    assert rel_l2(m.ft(band), g['ft_x']) < 1e-12
    assert rel_l2(m.ift(band), g['ift_x']) < 1e-12
    # direct flavour (the reference's GPU ordering) gives the same transform
    ft_d, ift_d = O.generate_ft(m.sh, m.weights, np.max(m.rs), 2.0, m.l_max, 'midpoint', 'direct')
    assert rel_l2(ft_d(band), g['ft_x']) < 1e-12


def test_reciprocal_projection(case):
    g, m = case
    assert abs(m.rp.integrated_intensity - g['integrated_intensity']) <= 1e-13 * abs(g['integrated_intensity'])
    assert np.array_equal(m.rp.radial_mask, g['radial_mask'])
    assert rel_l2(m.rp.deg2_invariants, g['deg2_ref']) < 1e-13
    rho_hat = m.ft(g['rho0'])
    assert rel_l2(rho_hat, g['rho_hat0']) < 1e-12
    sq = O.square_grid(g['rho_hat0'])
    assert rel_l2(sq, g['square0']) < 1e-15
    I = m.sh.forward_l(g['square0'])
    assert rel_l2(np.concatenate(I, axis=1), g['I_direct']) < 1e-12
    splits = np.arange(1, m.l_max + 1) ** 2
    I_ref = np.split(g['I_direct'], splits, axis=1)
    unk = m.rp.approximate_unknowns(I_ref)
    Ip = m.rp.mtip_projection(I_ref, unk)
    # the unknowns themselves are only defined up to LAPACK's choice in null spaces; the projected I'_lm is observable
    assert rel_l2(np.concatenate(Ip, axis=1), g['Iproj_direct']) < 1e-10
    out = m.rp.project_to_modified_intensity(g['rho_hat0'], g['square0'], g['I_proj_grid'])
    assert rel_l2(out, g['rho_hat_mod']) < 1e-14


def test_real_side(case):
    g, m = case
    rn = g['rho_new'].copy()
    rn_copy = rn.copy()
    proj = m.real_pr.projection(rn)
    assert np.array_equal(proj[0], g['rho_proj'])
    assert np.array_equal(proj[1]['all'], g['mask_all'])
    hio = O.hybrid_input_output(float(g['hio_beta']), rn_copy, proj, g['rho0'])
    assert rel_l2(hio, g['hio_out']) < 1e-15
    e = O.l2_projection_diff(m.integrator, rn_copy, proj, m.real_pr.initial_support)
    assert abs(e - g['real_err']) <= 1e-12 * abs(g['real_err'])


def test_shrink_wrap(case):
    g, m = case
    assert abs(m.sw.default_sigma - g['sw_default_sigma']) < 1e-13
    m.sw.set_sigma(12.5)
    m.sw.set_threshold(0.09)
    assert rel_l2(m.sw.gaussian_values, g['sw_gauss']) < 1e-15
    mask = m.shrink_wrap(g['rho0'].copy())
    assert (mask != g['sw_mask']).mean() < 1e-3   # threshold ties may flip single voxels


def test_density_guess(case):
    g, m = case
    rho0 = m.density_guess(np.random.default_rng(7))
    assert rel_l2(rho0, g['rho0']) < 1e-14


@pytest.mark.parametrize('tag', ['ref_small_ftstab', 'ref_small_plain', 'ref_small_shift'])   # the last one: shift_to_center output modifier
def test_full_loop(tag):
    g = load_golden(tag)
    m = O.MTIP(golden_settings(g), golden_data(g))
    res = m.run(rho0=g['rho0'].copy())
    assert res['loop_iterations'] == int(g['loop_iterations'])
    assert len(res['error_dict']['main']) == len(g['loop_main_error'])
    assert np.allclose(res['error_dict']['main'], g['loop_main_error'], rtol=1e-7, atol=0)
    assert abs(res['final_error'] - float(g['loop_final_error'])) <= 1e-7 * float(g['loop_final_error'])
    assert rel_l2(res['initial_density'], g['loop_initial_density']) < 1e-12
    assert rel_l2(res['last_real_density'], g['loop_last_real_density']) < 1e-7
    assert rel_l2(res['real_density'], g['loop_real_density']) < 1e-7
    assert rel_l2(res['last_reciprocal_density'], g['loop_last_reciprocal_density']) < 1e-7
    assert rel_l2(res['reciprocal_density'], g['loop_reciprocal_density']) < 1e-7
    assert (res['last_support_mask'] != g['loop_last_support_mask']).mean() < 1e-3
    assert (res['support_mask'] != g['loop_support_mask']).mean() < 1e-3
    assert rel_l2(res['last_deg2_invariant'], g['loop_last_deg2']) < 1e-7


# ----------------------------------------------------------------------------------------------
# radial-transform modes (midpoint, trapz, gauss; 3-D and 2-D) against the reference's own weight workers, assembly and
# CPU Hankel transforms -- tests/golden/make_golden_hankel_modes.py
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize('mode', ['midpoint', 'trapz', 'gauss'])
def test_hankel_modes_against_reference(mode):
    import os
    from oracle import mtip2d as O2
    from xframe_b200 import tables
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_hankel_modes.npz'))
    l_max, n_r, rc, q_max = int(g['l_max']), int(g['n_r']), float(g['rc']), float(g['q_max'])
    rs, qs = O.radial_grids(mode, q_max, n_r, rc)
    assert np.allclose(rs, g[f'{mode}_rs'], rtol=1e-15, atol=0) and np.allclose(qs, g[f'{mode}_qs'], rtol=1e-15, atol=0)
    r_max = rs.max()
    # 3-D
    w = O.hankel_weights(l_max, n_r, rc, mode)
    assert rel_l2(w, g[f'{mode}_3_weights']) < 1e-15
    aw = O.assemble_weights(w, r_max, rc, mode)
    assert rel_l2(aw['forward'], g[f'{mode}_3_forward']) < 1e-15 and rel_l2(aw['inverse'], g[f'{mode}_3_inverse']) < 1e-15
    zht, izht = O.generate_spherical_ht(aw, l_max, mode)
    cm = [g[f'{mode}_3_in_{i}'] for i in range(2 * l_max + 1)]
    f, b = zht(cm), izht(cm)
    for i in range(2 * l_max + 1):
        assert rel_l2(f[i], g[f'{mode}_3_zht_{i}']) < 1e-14 and rel_l2(b[i], g[f'{mode}_3_izht_{i}']) < 1e-14
    # the device-side split (real weights x scalar prefactor, phase in the kernel epilogue) reproduces the assembled arrays
    fs, iscale = tables.hankel_scales(r_max, n_r, rc, mode)
    tw = np.moveaxis(tables.hankel_weights(l_max, n_r, rc, mode), 0, 2)
    ls = np.arange(l_max + 1)
    assert rel_l2(tw * fs * (-1j) ** ls, g[f'{mode}_3_forward']) < 1e-15 and rel_l2(tw * iscale * (1j) ** ls, g[f'{mode}_3_inverse']) < 1e-15
    # 2-D
    w2 = O2.polar_hankel_weights(l_max, n_r, rc, mode)
    assert rel_l2(w2, g[f'{mode}_2_weights']) < 1e-15
    aw2 = O2.assemble_weights_2d(w2, r_max, rc, mode)
    assert rel_l2(aw2['forward'], g[f'{mode}_2_forward']) < 1e-15 and rel_l2(aw2['inverse'], g[f'{mode}_2_inverse']) < 1e-15
    z2, iz2 = O2.generate_polar_ht(aw2, mode)
    assert rel_l2(z2(g[f'{mode}_2_in']), g[f'{mode}_2_zht']) < 1e-14 and rel_l2(iz2(g[f'{mode}_2_in']), g[f'{mode}_2_izht']) < 1e-14
    fs2, is2 = tables.polar_hankel_scales(r_max, n_r, rc, mode)
    m_all = np.concatenate((ls, -ls[:0:-1]))
    tw2 = np.moveaxis(tables.polar_hankel_device_weights(tables.polar_hankel_weights(l_max, n_r, rc, mode)), 0, 2)
    assert rel_l2(tw2 * fs2 * (-1j) ** m_all, g[f'{mode}_2_forward']) < 1e-15 and rel_l2(tw2 * is2 * (1j) ** m_all, g[f'{mode}_2_inverse']) < 1e-15
