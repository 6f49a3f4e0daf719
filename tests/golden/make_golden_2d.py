"""Golden vectors for the 2-D (polar) flavour, from the UNMODIFIED reference in /root/reference.

    python tests/golden/make_golden_2d.py

Same harness as make_golden.py (stubbed xframe._version, hand-built settings.project / database.project); with
`dimensions: 2` no third-party transform is involved (numpy FFT only).  Runs the reference's own operator table and
its full MTIP.phasing_loop on seeded inputs and stores inputs + outputs under tests/golden/ref2d_*.npz.
"""
import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden import import_reference, settings_dict, FakeDB          # noqa: E402


def settings_dict_2d(n_r, m_max, max_q, ft_stab, particle_radius=250.0, so_freedom=False):
    sd = settings_dict(n_r, m_max, 1, 2 * m_max + 1, max_q, ft_stab, particle_radius)
    sd['dimensions'] = 2
    sd['grid'] = {'max_q': float(max_q), 'max_order': m_max, 'n_radial_points': n_r}
    sd['output_density_modifiers'] = {'shift_to_center': False, 'fix_orientation': bool(so_freedom)}
    sd['projections']['reciprocal']['SO_freedom'] = {'use': bool(so_freedom), 'radial_high_pass': 0.2}
    return sd


def build_case_2d(xframe, n_r, m_max, ft_stab, tag, seed=11, so_freedom=False):
    from xframe.library.pythonLibrary import DictNamespace, RecipeFactory
    from xframe.library.gridLibrary import SampledFunction, NestedArray
    from xframe import settings
    import xframe.database as database
    from xframe.library import mathLibrary as mLib
    from oracle import mtip as O, mtip2d as O2

    max_q = 2.0 * n_r / 794.0
    sd = settings_dict_2d(n_r, m_max, max_q, ft_stab, so_freedom=so_freedom)
    settings.project = DictNamespace.dict_to_dictnamespace(copy.deepcopy(sd))
    settings.general.cache_aware = False
    settings.general.n_control_workers = 0
    # synthetic invariants on a FINER data q-grid so that the reference's cubic regridding runs (as for the 3-D cases)
    sd_data = settings_dict_2d(2 * n_r, m_max, max_q, ft_stab)
    qs_data = O.radial_grids('midpoint', max_q, 2 * n_r, 2.0)[1]
    om = O2.MTIP2D(sd_data, {'data_radial_points': qs_data, 'average_intensity': np.ones(2 * n_r), 'max_order': m_max,
                             'data_projection_matrices': np.zeros((m_max + 1, 2 * n_r), complex)})
    inv = O2.invariants_from_density_2d(O2.disk_model_density(om.real_grid), om.ft, om.qs, om.real_grid[0, :, 1])

    def fresh_data():
        data = dict(inv)
        data['data_projection_matrices'] = inv['data_projection_matrices'].copy()
        data['average_intensity'] = SampledFunction(NestedArray(inv['data_radial_points'][:, None].copy(), 1),
                                                    inv['average_intensity'].copy(), coord_sys='cartesian')
        return data
    database.project = FakeDB(fresh_data())

    import importlib
    if 'xframe.projects.fxs.reconstruct' in sys.modules:
        rec = importlib.reload(sys.modules['xframe.projects.fxs.reconstruct'])
    else:
        rec = importlib.import_module('xframe.projects.fxs.reconstruct')
    os.chdir(ROOT)
    rec.set_globals()
    rec.MTIP.preinit()
    m = rec.MTIP(RecipeFactory({}))
    m.generate_phasing_loop()
    ops = m.process_factory.operatorDict

    rng = np.random.default_rng(seed)
    gshape = m.grid_pair.realGrid[:].shape[:-1]
    out = {'n_r': n_r, 'm_max': m_max, 'n_phi': gshape[1], 'max_q': max_q, 'ft_stab': ft_stab, 'so_freedom': bool(so_freedom),
           'avg_intensity': inv['average_intensity'], 'data_q': inv['data_radial_points'], 'pm': inv['data_projection_matrices'],
           'phis': m.grid_pair.realGrid[0, :, 1], 'rs': m.grid_pair.realGrid[:, 0, 0], 'qs': m.grid_pair.reciprocalGrid[:, 0, 0]}
    x = rng.normal(size=gshape) + 1j * rng.normal(size=gshape)
    out['x_grid'] = x
    out['cht_complex_forward'] = ops['complex_harmonic_transform'](x.copy())
    out['cht_complex_inverse'] = ops['complex_inverse_harmonic_transform'](x.copy())
    out['cht_real_forward'] = ops['harmonic_transform'](x.copy())
    out['cht_real_inverse'] = ops['inverse_harmonic_transform'](out['cht_real_forward'].copy())
    out['ft_x'] = np.array(ops['fourier_transform'](x.copy()))
    out['ift_x'] = np.array(ops['inverse_fourier_transform'](x.copy()))
    from xframe.projects.fxs.projectLibrary.hankel_transforms import generate_ht
    w = rec.MTIP.fourier_transform_weights
    out['hankel_weights'] = w['weights']
    r_max = np.max(rec.MTIP.real_radial_points)
    zht, izht = generate_ht(w['weights'], np.arange(m_max + 1), r_max, reciprocity_coefficient=2.0, dimensions=2, use_gpu=False, mode='midpoint')
    cm = out['cht_complex_forward']
    out['hankel_fwd'] = np.array(zht(cm.copy()))
    out['hankel_inv'] = np.array(izht(cm.copy()))

    A = 1 + 1 / 2 * np.random.default_rng(seed).random(gshape)
    bump = mLib.get_test_function(support=[-250.0, 250.0], slope=0.3)
    integ = mLib.PolarIntegrator(m.grid_pair.realGrid[:])
    dens0 = A * bump(m.grid_pair.realGrid[..., 0])
    dens0 = (dens0 * np.sqrt(m.rprojection.integrated_intensity / integ.integrate((dens0 * dens0.conj()).real))).astype(complex)
    out['rho0'] = dens0
    out['integrated_intensity'] = m.rprojection.integrated_intensity
    out['integral_of_x2'] = integ.integrate((x * x.conj()).real)
    rho_hat = np.array(ops['fourier_transform'](dens0.copy()))
    out['rho_hat0'] = rho_hat
    sq = np.array(ops['square_grid'](rho_hat))
    I = np.array(ops['harmonic_transform'](sq.copy()))
    unk = np.array(ops['approximate_unknowns'](I))
    Ip = np.array(ops['mtip_projection'](I, unk))
    out['I_m'], out['unknowns'], out['Iproj_m'] = I, unk, Ip
    I_proj_grid = np.array(ops['inverse_harmonic_transform'](Ip.copy()))
    out['I_proj_grid'] = I_proj_grid
    out['rho_hat_mod'] = np.array(ops['project_to_modified_intensity'](rho_hat, sq.copy(), I_proj_grid))
    rn = np.array(ops['inverse_fourier_transform'](out['rho_hat_mod'].copy()))
    out['rho_new'] = rn.copy()
    proj = ops['real_projection'](rn)
    out['rho_proj'] = np.array(proj[0])
    out['real_err'] = ops['real_errors'](out['rho_new'].copy(), proj)['l2_projection_diff']
    out['sw_default_sigma'] = m.projection_objects['sw'].default_sigma
    m.projection_objects['sw'].gaussian_sigma = 12.5
    m.projection_objects['sw'].threshold = 0.09
    out['sw_mask'] = m.routines['SW'].run(dens0.copy())
    out['radial_mask'] = np.array(m.rprojection.radial_mask)
    out['projection_matrices_final'] = np.array(m.rprojection.projection_matrices)

    orig = rec.MTIP.generate_density_guess_method
    rec.MTIP.generate_density_guess_method = lambda self, spec, grid: (lambda: dens0.copy())
    try:
        rec.MTIP.mtip_data = fresh_data()
        m3 = rec.MTIP(RecipeFactory({}))
        m3.generate_phasing_loop()
        res = m3.phasing_loop()
    finally:
        rec.MTIP.generate_density_guess_method = orig
    out['loop_main_error'] = res['error_dict']['main']
    out['loop_real_density'] = res['real_density']
    out['loop_last_real_density'] = res['last_real_density']
    out['loop_last_reciprocal_density'] = res['last_reciprocal_density']
    out['loop_last_support_mask'] = res['last_support_mask']
    out['loop_final_error'] = res['final_error']
    out['loop_unknowns'] = np.array(res['fxs_unknowns'])
    out['loop_last_deg2'] = res['last_deg2_invariant']
    path = os.path.join(HERE, f'{tag}.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    xf = import_reference()
    build_case_2d(xf, n_r=16, m_max=7, ft_stab=True, tag='ref2d_small_ftstab')
    build_case_2d(xf, n_r=24, m_max=15, ft_stab=False, tag='ref2d_medium_plain')
    build_case_2d(xf, n_r=24, m_max=15, ft_stab=True, tag='ref2d_medium_so', so_freedom=True)   # the 2-D defaults: SO_freedom + fix_orientation
