"""Generate golden vectors by running the UNMODIFIED reference from /root/reference.

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

What it does
  * stubs the setuptools_scm artefact ``xframe._version`` (pyproject.toml:64-67),
    points HOME at a scratch dir (xframe/logger.py:6-9 writes ~/.xframe/log.txt),
  * injects ``oracle.sht.sh`` at the reference's own plugin slot
    ``xframe.library.mathLibrary.shtns`` (startup_routines.py:57) -- shtns is the
    one third-party piece that is absent here,
  * hand-builds ``settings.project`` / ``database.project`` (no ruamel/h5py here),
  * runs the reference's own operators (hankel_transforms, fourier_transforms,
    fxs_Projections, fxs_IO_methods, misk, mathLibrary) and its own
    ``MTIP.phasing_loop`` on seeded inputs and stores inputs + outputs under
    tests/golden/*.npz.
"""
import os
import sys
import types
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = '/root/reference'


def import_reference():
    os.environ['HOME'] = tempfile.mkdtemp(prefix='xf_home_')
    m = types.ModuleType('xframe._version')
    m.__version__ = m.version = '0.0.0'
    sys.modules['xframe._version'] = m
    sys.path.insert(0, REF)
    import warnings
    warnings.filterwarnings('ignore')
    import xframe
    from oracle.sht import sh
    import xframe.library.mathLibrary as mLib
    mLib.shtns = sh
    return xframe


def settings_dict(n_r, l_max, n_theta, n_phi, max_q, ft_stab, particle_radius=250.0):
    return {
        'dimensions': 3, 'structure_name': 'golden', 'particle_radius': particle_radius,
        'grid': {'max_q': float(max_q), 'max_order': l_max, 'n_phi': n_phi, 'n_theta': n_theta, 'n_radial_points': n_r},
        'fourier_transform': {'type': 'midpoint', 'reciprocity_coefficient': 2.0,
                              'allow_weight_calculation': True, 'allow_weight_saving': False},
        'density_guess': {'type': 'bump', 'bump': {'slope': 0.3}, 'radius': particle_radius,
                          'amplitude_function': 'random', 'random': {'SNR': 2}},
        'projections': {
            'real': {
                'projections': {
                    'apply': ['support', 'value_threshold', 'limit_imag'],
                    'value_threshold': {'threshold': [0, False]},
                    'limit_imag': {'threshold': 2},
                    'support': {'initial_support': {'type': 'max_radius', 'max_radius': particle_radius},
                                'enforce_initial_support': {'apply': True, 'if_error_bigger_than': 6e-3}},
                },
                'shrink_wrap': {'sigmas': [[20, [False, 5], -2], False], 'thresholds': [0.09, 0.09]},
                'HIO': {'beta': [[0.5, 0.4, -1 / 250, 500], [0.01, 0.002, -1 / 200, 200]],
                        'considered_projections': ['all']},
            },
            'reciprocal': {
                'number_of_particles': {'initial': 1.0, 'estimate': False},
                'regrid': {'interpolation': 'cubic'},
                'used_order_ids': np.arange(l_max + 1),
                'odd_orders_to_0': True, 'use_averaged_intensity': True,
                'q_mask': {'type': 'none'}, 'SO_freedom': {'use': False},
            },
        },
        'output_density_modifiers': {'shift_to_center': False},
        'main_loop': {
            'error': {'methods': {
                'real': {'calculate': ['l2_projection_diff'], 'l2_projection_diff': {'inside_initial_support': True}},
                'reciprocal': {'calculate': []},
                'main': {'metrics': {'real': ['l2_projection_diff'], 'reciprocal': []}, 'type': 'mean'}},
                'limits': {'use': False}, 'gain_limits': {'use': False}},
            'sub_loops': {
                'order': ['main', 'refinement'],
                'main': {'methods': {'HIO': {'iterations': 3, 'ft_stab': ft_stab},
                                     'ER': {'iterations': 2, 'ft_stab': ft_stab}, 'SW': 1},
                         'order': ['HIO', 'SW', 'ER'], 'iterations': 2,
                         'best_density_not_in_first_n_iterations': np.inf},
                'refinement': {'methods': {'ER': {'iterations': 3, 'ft_stab': ft_stab}, 'SW': 1},
                               'order': ['SW', 'ER'], 'iterations': 1,
                               'best_density_not_in_first_n_iterations': np.inf},
            }},
        'GPU': {'use': False, 'n_gpu_workers': 1},
        'multi_process': {'use': False, 'n_parallel_reconstructions': 1},
        'profiling': {'enable': False, 'reconstruction_process_id': 1},
    }


class FakeDB:
    """Minimal stand-in for projects/fxs/_database_.py used by MTIP.preinit (reconstruct.py:241-285)."""

    def __init__(self, data):
        self.data = data

    def load(self, name, **kw):
        if name == 'invariants':
            return self.data
        raise FileNotFoundError(name)

    def save(self, *a, **k):
        pass

    def update_settings(self, *a, **k):
        pass


def build_case(xframe, n_r, l_max, n_theta, n_phi, ft_stab, tag, run_loop=True, seed=7, shift_to_center=False):
    from xframe.library.pythonLibrary import DictNamespace
    from xframe.library.gridLibrary import SampledFunction, NestedArray
    from xframe import settings
    import xframe.database as database
    from oracle import mtip as O

    r_max_domain = 794.0
    max_q = 2.0 * n_r / r_max_domain
    sd = settings_dict(n_r, l_max, n_theta, n_phi, max_q, ft_stab)
    sd['output_density_modifiers'] = {'shift_to_center': bool(shift_to_center)}
    settings.project = DictNamespace.dict_to_dictnamespace(sd)
    settings.general.cache_aware = False
    settings.general.n_control_workers = 0

    # synthetic invariants with the oracle's helper, sampled on a FINER data q-grid (2*n_r midpoints over the
    # same [0,max_q]) so that the reference's cubic regridding runs (its no-regrid branch raises
    # UnboundLocalError: fxs_Projections.py:644-676 leaves `low_res` unset); the reference consumes the record
    # through its own loader contract
    sd_data = settings_dict(2 * n_r, l_max, n_theta, n_phi, max_q, ft_stab)
    om = O.MTIP(sd_data, {'data_radial_points': O.radial_grids('midpoint', max_q, 2 * n_r, 2.0)[1],
                          'average_intensity': np.ones(2 * n_r), 'max_order': l_max,
                          'data_projection_matrices': [np.zeros((2 * n_r, min(2 * n_r, 2 * l + 1)), complex) for l in range(l_max + 1)]})
    dens = O.six_sphere_density(om.real_grid)
    inv = O.invariants_from_density(dens, om.ft, om.sh, om.qs)
    def fresh_data():
        # the reference regrids average_intensity in place (fxs_Projections.py:668), so every MTIP gets its own record
        data = dict(inv)
        pm = np.empty(len(inv['data_projection_matrices']), dtype=object)
        for i, p in enumerate(inv['data_projection_matrices']):
            pm[i] = p.copy()
        data['data_projection_matrices'] = pm
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
    from xframe.library.pythonLibrary import RecipeFactory
    m = rec.MTIP(RecipeFactory({}))
    m.generate_phasing_loop()
    ops = m.process_factory.operatorDict

    rng = np.random.default_rng(seed)
    gshape = m.grid_pair.realGrid[:].shape[:-1]
    out = {'n_r': n_r, 'l_max': l_max, 'n_theta': n_theta, 'n_phi': n_phi, 'max_q': max_q, 'ft_stab': ft_stab, 'shift_to_center': bool(shift_to_center),
           'avg_intensity': inv['average_intensity'], 'data_q': inv['data_radial_points']}
    for l, p in enumerate(inv['data_projection_matrices']):
        out[f'pm_{l}'] = p
    out['thetas'] = m.grid_pair.realGrid[0, :, 0, 1]
    out['phis'] = m.grid_pair.realGrid[0, 0, :, 2]
    out['rs'] = m.grid_pair.realGrid[:, 0, 0, 0]
    out['qs'] = m.grid_pair.reciprocalGrid[:, 0, 0, 0]

    # --- operator-level vectors from the reference's own operator table (reconstruct.py:303-316) ---
    x = (rng.normal(size=gshape) + 1j * rng.normal(size=gshape))
    out['x_grid'] = x
    cl = ops['harmonic_transform'](x.copy())
    out['sht_forward_direct'] = np.concatenate(cl, axis=1)
    band = ops['inverse_harmonic_transform'](cl)
    out['sht_inverse_of_forward'] = band
    out['ft_x'] = ops['fourier_transform'](band.copy())
    out['ift_x'] = ops['inverse_fourier_transform'](band.copy())
    # Hankel on its own, reference CPU flavour (hankel_transforms.py:642-658) on m-ordered lists
    from xframe.projects.fxs.projectLibrary.hankel_transforms import generate_ht
    from xframe.projects.fxs.projectLibrary.misk import _get_reciprocity_coefficient
    w = rec.MTIP.fourier_transform_weights
    out['hankel_weights'] = w['weights']
    r_max = np.max(rec.MTIP.real_radial_points)
    zht, izht = generate_ht(w['weights'], np.arange(l_max + 1), r_max, reciprocity_coefficient=2.0, dimensions=3, use_gpu=False, mode='midpoint')
    cm = m.projection_objects and None
    from xframe.library import mathLibrary as mLib
    sh_obj = mLib.get_spherical_harmonic_transform_obj(l_max, mode='complex', n_phi=n_phi, n_theta=n_theta)
    cm = sh_obj.forward_m(band.copy())
    hm = zht(cm)
    full = np.zeros((n_r, (l_max + 1) ** 2), complex)
    full_i = np.zeros_like(full)
    him = izht(cm)
    for mid, idx in enumerate(sh_obj.cplx_m_indices):
        full[:, idx] = hm[mid]
        full_i[:, idx] = him[mid]
    out['hankel_in_direct'] = sh_obj.forward_d(band.copy())
    out['hankel_fwd_direct'] = full
    out['hankel_inv_direct'] = full_i

    # intensity-like input: projection step
    rho = m.generate_density_guess_method(settings.project.density_guess, m.grid_pair.realGrid)
    # deterministic amplitude instead of os.urandom (reconstruct.py:1119): same formula with an injected generator
    A = 1 + 1 / 2 * np.random.default_rng(seed).random(gshape)
    bump = mLib.get_test_function(support=[-250.0, 250.0], slope=0.3)
    integ = mLib.SphericalIntegrator(m.grid_pair.realGrid[:])
    dens0 = A * bump(m.grid_pair.realGrid[..., 0])
    dens0 = dens0 * np.sqrt(m.rprojection.integrated_intensity / integ.integrate((dens0 * dens0.conj()).real))
    dens0 = dens0.astype(complex)
    out['rho0'] = dens0
    out['integrated_intensity'] = m.rprojection.integrated_intensity
    rho_hat = ops['fourier_transform'](dens0.copy())
    out['rho_hat0'] = rho_hat
    sq = ops['square_grid'](rho_hat)
    out['square0'] = np.array(sq)
    I = ops['harmonic_transform'](np.array(sq))
    unk = ops['approximate_unknowns'](I)
    Ip = ops['mtip_projection'](I, unk)
    out['I_direct'] = np.concatenate(I, axis=1)
    out['Iproj_direct'] = np.concatenate(Ip, axis=1)
    for l in range(l_max + 1):
        out[f'unk_{l}'] = np.array(unk[l])
    I_proj_grid = ops['inverse_harmonic_transform'](Ip)
    out['I_proj_grid'] = I_proj_grid
    out['rho_hat_mod'] = np.array(ops['project_to_modified_intensity'](rho_hat, np.array(sq), I_proj_grid))
    # real-space side
    rn = ops['inverse_fourier_transform'](out['rho_hat_mod'].copy())
    out['rho_new'] = np.array(rn)
    rn_copy = np.array(rn)
    proj = ops['real_projection'](rn)
    out['rho_proj'] = np.array(proj[0])
    out['mask_all'] = np.array(proj[1]['all'])
    m.projection_objects['hio'].beta = 0.37
    out['hio_beta'] = 0.37
    out['hio_out'] = ops['hybrid_input_output'](rn_copy, proj, dens0)
    out['real_err'] = ops['real_errors'](rn_copy, proj)['l2_projection_diff']
    # shrink wrap
    out['sw_default_sigma'] = m.projection_objects['sw'].default_sigma
    m.projection_objects['sw'].gaussian_sigma = 12.5
    m.projection_objects['sw'].threshold = 0.09
    out['sw_mask'] = m.routines['SW'].run(dens0.copy())
    out['sw_gauss'] = np.array(m.projection_objects['sw'].gaussian_values)
    out['radial_mask'] = np.array(m.rprojection.radial_mask)
    out['deg2_ref'] = np.array(m.rprojection.deg2_invariants)

    if run_loop:
        # full reference loop from an injected initial density
        # swap the guess closure: main_loop captures real_density_guess_method at assembly time
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
        out['loop_real_error'] = res['error_dict']['real']['l2_projection_diff']
        out['loop_real_density'] = res['real_density']
        out['loop_last_real_density'] = res['last_real_density']
        out['loop_reciprocal_density'] = res['reciprocal_density']
        out['loop_last_reciprocal_density'] = res['last_reciprocal_density']
        out['loop_support_mask'] = res['support_mask']
        out['loop_last_support_mask'] = res['last_support_mask']
        out['loop_final_error'] = res['final_error']
        out['loop_iterations'] = res['loop_iterations']
        out['loop_last_deg2'] = res['last_deg2_invariant']
        out['loop_initial_density'] = res['initial_density']
    path = os.path.join(HERE, f'{tag}.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    xf = import_reference()
    build_case(xf, n_r=16, l_max=7, n_theta=8, n_phi=16, ft_stab=True, tag='ref_small_ftstab')
    build_case(xf, n_r=16, l_max=7, n_theta=8, n_phi=16, ft_stab=False, tag='ref_small_plain')
    build_case(xf, n_r=32, l_max=15, n_theta=16, n_phi=32, ft_stab=True, tag='ref_medium_ops', run_loop=False)
    build_case(xf, n_r=16, l_max=7, n_theta=8, n_phi=16, ft_stab=True, tag='ref_small_shift', shift_to_center=True)
