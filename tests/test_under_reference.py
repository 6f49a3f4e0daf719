"""The reference's own code driving / driven beside xframe_b200.

Needs the reference package (`baseline/_ref`, installed by `__graft_entry__.build()`, or /root/reference in the build
container); skipped where neither exists.  Settings are the reference's OWN defaults (settings.default_settings() mirrors
settings/reconstruct/default_0.01.yaml, including `apply: [support, value_threshold, assert_real]`) with the overrides of
its reconstruct integration test (tests/test_fxs_integration.py:326-351: N_r 8, max_order 15, rc 2.0) and shortened loops.

  * CPU: reference MTIP.phasing_loop (oracle.sht.sh at the shtns slot) == oracle MTIP on the same input   (pins the oracle)
  * GPU: the reference's unmodified generate_spherical_ht_gpu -> xframe_b200.gpu_access.ClProcess == its CPU Hankel;
         reference MTIP.phasing_loop with xframe_b200's CUDA `sh` plugin and CUDA GPU layer == xframe_b200.worker.ProjectWorker
Tolerances: relative L2 <= 1e-6 on error histories and densities (projection step), 1e-12 on the Hankel.
"""
import copy

import numpy as np
import pytest

import ref_harness as RH
from helpers import rel_l2
from oracle import mtip as O
from xframe_b200 import settings as XS

needs_ref = pytest.mark.skipif(RH.reference_root() is None, reason='reference package not available (baseline/_ref)')


def reference_test_settings(gpu, n_r=8):
    over = {'structure_name': 'test', 'dimensions': 3, 'particle_radius': 250,
            'grid': {'n_radial_points': n_r, 'max_order': 15, 'n_theta': 16, 'n_phi': 32},
            'projections': {'reciprocal': {'used_order_ids': np.arange(16)}},
            'fourier_transform': {'reciprocity_coefficient': 2.0, 'allow_weight_saving': False},
            'multi_process': {'use': False, 'n_parallel_reconstructions': 1},
            'GPU': {'use': bool(gpu), 'n_gpu_workers': 1},
            'main_loop': {'sub_loops': {
                'main': {'iterations': 2, 'methods': {'HIO': {'iterations': 4, 'ft_stab': True}, 'ER': {'iterations': 3, 'ft_stab': True}, 'SW': 1}},
                'refinement': {'iterations': 1, 'methods': {'ER': {'iterations': 3, 'ft_stab': True}, 'SW': 1}}}}}
    sd = XS.finalize(XS.merge(XS.default_settings(), over))
    assert sd['projections']['real']['projections']['apply'] == ['support', 'value_threshold', 'assert_real']   # the reference default
    return sd


def synthetic_invariants(sd):
    """Invariants of the six-sphere model on a FINER q grid than the reconstruction grid, so that the reference's cubic
    regridding runs (its no-regrid branch raises UnboundLocalError, fxs_Projections.py:644-676)."""
    g = sd['grid']
    n_r, l_max = 2 * g['n_radial_points'], g['max_order']
    max_q = 2.0 * g['n_radial_points'] / 794.0
    sdd = copy.deepcopy(sd)
    sdd['grid'].update(n_radial_points=n_r, max_q=float(max_q))
    sdd['projections']['real']['projections']['apply'] = ['support', 'value_threshold', 'limit_imag']
    sdd['projections']['reciprocal']['q_mask'] = {'type': 'none'}
    om = O.MTIP(sdd, {'data_radial_points': O.radial_grids('midpoint', max_q, n_r, 2.0)[1], 'average_intensity': np.ones(n_r),
                      'max_order': l_max,
                      'data_projection_matrices': [np.zeros((n_r, min(n_r, 2 * l + 1)), complex) for l in range(l_max + 1)]})
    inv = O.invariants_from_density(O.six_sphere_density(om.real_grid), om.ft, om.sh, om.qs)
    inv.update({'dimensions': 3, 'xray_wavelength': 1.23984, 'data_angular_points': om.sh.phi, 'number_of_particles': 1})
    return inv


def initial_density(m_oracle, seed=11):
    return m_oracle.density_guess(np.random.default_rng(seed)).astype(complex)


@needs_ref
@pytest.mark.parametrize('n_r', [8, 16])       # 8: the reference's own test config (zero error inside the two support shells); 16: non-trivial errors
def test_reference_loop_with_default_apply_list_matches_oracle(n_r):
    from oracle.sht import sh
    sd = reference_test_settings(gpu=False, n_r=n_r)
    inv = synthetic_invariants(sd)
    RH.import_reference(sh_class=sh)
    mo = O.MTIP(sd, dict(inv))
    rho0 = initial_density(mo)
    rec, m = RH.make_mtip(sd, inv, rho0=rho0)
    ref = m.phasing_loop()
    got = mo.run(rho0=rho0.copy())
    if n_r > 8:
        assert min(ref['error_dict']['main']) > 0
    assert np.allclose(got['error_dict']['main'], ref['error_dict']['main'], rtol=1e-7, atol=0)
    assert rel_l2(got['last_real_density'], ref['last_real_density']) < 1e-7
    assert got['loop_iterations'] == ref['loop_iterations']


def sketch_tail_settings(gpu, variant='all'):
    """SW_center, the non-FXS methods (fixed-intensity projection) and a finite best_density_not_in_first_n_iterations
    (reconstruct.py:529-534,606-613,886-904,945-949).  variant 'all': everything in one schedule (pins the exact semantics, incl.
    which pair the reference's `hist` names, on the CPU); 'non_fxs' / 'sw_center': one feature per schedule for the device
    parity test -- SW_center exchanges the real and reciprocal densities (reference quirk), after which FT(FT(rho)) ~ 1e15 flows
    through the fixed-intensity division and rounding differences between implementations are amplified beyond 1e-6."""
    sd = reference_test_settings(gpu, n_r=16)
    if variant == 'non_fxs':
        sd['main_loop']['sub_loops'] = {
            'order': ['main', 'refinement'],
            'main': {'iterations': 2, 'order': ['HIO', 'HIO_non_FXS', 'SW', 'ER_non_FXS', 'ER'], 'best_density_not_in_first_n_iterations': 0,
                     'methods': {'HIO': {'iterations': 3, 'ft_stab': True}, 'SW': 1, 'HIO_non_FXS': {'iterations': 2, 'ft_stab': True},
                                 'ER_non_FXS': {'iterations': 2, 'ft_stab': False}, 'ER': {'iterations': 2, 'ft_stab': True}}},
            'refinement': {'iterations': 1, 'order': ['ER_non_FXS', 'SW', 'ER'], 'best_density_not_in_first_n_iterations': np.inf,
                           'methods': {'ER_non_FXS': {'iterations': 2, 'ft_stab': True}, 'SW': 1, 'ER': {'iterations': 2, 'ft_stab': True}}}}
        return sd
    if variant == 'sw_center':
        sd['main_loop']['sub_loops'] = {
            'order': ['main'],
            'main': {'iterations': 2, 'order': ['HIO', 'SW_center', 'ER'], 'best_density_not_in_first_n_iterations': 1,
                     'methods': {'HIO': {'iterations': 3, 'ft_stab': True}, 'SW_center': {'iterations': 1}, 'ER': {'iterations': 2, 'ft_stab': True}}}}
        return sd
    sd['main_loop']['sub_loops'] = {
        'order': ['main', 'refinement'],
        'main': {'iterations': 2, 'order': ['HIO', 'SW_center', 'HIO_non_FXS', 'ER_non_FXS', 'ER'], 'best_density_not_in_first_n_iterations': 0,
                 'methods': {'HIO': {'iterations': 3, 'ft_stab': True}, 'SW_center': {'iterations': 2}, 'HIO_non_FXS': {'iterations': 2, 'ft_stab': True},
                             'ER_non_FXS': {'iterations': 2, 'ft_stab': False}, 'ER': {'iterations': 2, 'ft_stab': True}}},
        'refinement': {'iterations': 1, 'order': ['ER_non_FXS', 'SW', 'ER'], 'best_density_not_in_first_n_iterations': np.inf,
                       'methods': {'ER_non_FXS': {'iterations': 2, 'ft_stab': True}, 'SW': 1, 'ER': {'iterations': 2, 'ft_stab': True}}}}
    return sd


@needs_ref
@pytest.mark.parametrize('variant', ['all', 'non_fxs', 'sw_center'])
def test_reference_sketch_tail_matches_oracle(variant):
    from oracle.sht import sh
    sd = sketch_tail_settings(gpu=False, variant=variant)
    inv = synthetic_invariants(sd)
    RH.import_reference(sh_class=sh)
    mo = O.MTIP(sd, dict(inv))
    rho0 = initial_density(mo)
    rec, m = RH.make_mtip(sd, inv, rho0=rho0)
    ref = m.phasing_loop()
    got = mo.run(rho0=rho0.copy())
    assert min(ref['error_dict']['main']) > 0 and len(ref['error_dict']['main']) == {'all': 22, 'non_fxs': 22, 'sw_center': 10}[variant]
    assert np.allclose(got['error_dict']['main'], ref['error_dict']['main'], rtol=1e-7, atol=0)
    for k in ('last_real_density', 'real_density', 'last_reciprocal_density', 'reciprocal_density'):
        assert rel_l2(got[k], ref[k]) < 1e-7, k
    assert np.array_equal(got['support_mask'], ref['support_mask']) and np.array_equal(got['last_support_mask'], ref['last_support_mask'])


@needs_ref
def test_invariant_extraction_matches_reference():
    """N3: density -> B_l -> V_l.  oracle.invariants_from_density (the restatement setup_host.invariants_from_density is
    tested against on the GPU) against the reference's own density_to_deg2_invariants (fxs_invariant_tools.py:889-898) and
    deg2_invariant_to_projection_matrices_3d (:1171-1207).  Eigenvectors are defined up to sign / rotations inside degenerate
    eigenspaces, so the projection matrices are compared through V_l V_l^H."""
    from oracle.sht import sh
    RH.import_reference(sh_class=sh)
    from xframe.projects.fxs.projectLibrary import fxs_invariant_tools as FI
    from xframe.projects.fxs.projectLibrary.harmonic_transforms import HarmonicTransform
    sd = reference_test_settings(gpu=False, n_r=16)
    l_max, n_r = 15, 16
    max_q = 2.0 * n_r / 794.0
    sdd = copy.deepcopy(sd)
    sdd['grid'].update(max_q=float(max_q))
    sdd['projections']['real']['projections']['apply'] = ['support', 'value_threshold', 'limit_imag']
    om = O.MTIP(sdd, {'data_radial_points': O.radial_grids('midpoint', max_q, n_r, 2.0)[1], 'average_intensity': np.ones(n_r), 'max_order': l_max,
                      'data_projection_matrices': [np.zeros((n_r, min(n_r, 2 * l + 1)), complex) for l in range(l_max + 1)]})
    dens = O.six_sphere_density(om.real_grid)
    inv = O.invariants_from_density(dens, om.ft, om.sh, om.qs)
    cht = HarmonicTransform('complex', {'dimensions': 3, 'max_order': l_max, 'n_phi': 32, 'n_theta': 16, 'anti_aliazing_degree': 2})
    Bl_ref = FI.density_to_deg2_invariants(dens.astype(complex), om.ft, 3, cht=cht)
    assert rel_l2(inv['deg_2_invariant'], Bl_ref) < 1e-12
    for l in range(l_max + 1):
        pm_ref, ev = FI.deg2_invariant_to_projection_matrices_3d(np.array(Bl_ref[l]), [[0, n_r]], l, 0)
        ours = 2 * inv['data_projection_matrices'][l]            # the oracle hands V_l / 2 (cancels fxs_Projections.py:711-713)
        assert ours.shape == pm_ref.shape
        a, b = ours @ ours.conj().T, pm_ref @ pm_ref.conj().T
        # even orders: to rounding.  Odd orders vanish for a real density (Friedel symmetry): what is left is rounding noise with
        # an imaginary part the oracle drops (the device path needs real V_l; reconstruct zeroes odd orders anyway) -- compared
        # on the scale of the even orders
        scale = np.linalg.norm(b) if l % 2 == 0 else np.linalg.norm(Bl_ref[0])
        assert np.linalg.norm(a - b) <= 1e-10 * scale, l


@needs_ref
@pytest.mark.parametrize('q_mask', [
    {'type': 'none'},
    {'type': 'from_projection_matrices'},
    {'type': 'manual', 'manual': {'type': 'region', 'region': [0.01, 0.03]}},
    {'type': 'manual', 'manual': {'type': 'region', 'region': [False, 0.025]}},
    {'type': 'manual', 'manual': {'type': 'order_dependent_line', 'order_dependent_line': [[2, 0.004], [14, 0.03]]}},
])
def test_radial_mask_types_match_reference(q_mask):
    """generate_radial_mask (fxs_Projections.py:578-629) for every q_mask type against setup_host.ProjectionSetup (host-side setup,
    numpy like the reference's): the mask that selects which radial points of an order the projection overwrites."""
    from oracle.sht import sh
    from xframe_b200 import setup_host as S
    sd = reference_test_settings(gpu=False, n_r=16)
    sd['projections']['reciprocal']['q_mask'] = q_mask
    inv = synthetic_invariants(sd)
    n_q = len(inv['data_radial_points'])
    inv['data_projection_matrices_q_id_limits'] = {'I1I1': np.array([[l % 3, n_q - (l % 4)] for l in range(16)])}
    RH.import_reference(sh_class=sh)
    rec, m = RH.make_mtip(sd, inv)
    ref_mask = np.asarray(m.rprojection.radial_mask)
    ps = S.ProjectionSetup(np.asarray(rec.MTIP.reciprocal_radial_points), dict(inv), 15, sd['projections']['reciprocal'])
    assert ps.radial_mask.shape == (16, 16)
    assert np.array_equal(ps.radial_mask, np.broadcast_to(ref_mask, (16, 16)))
    assert np.array_equal(np.broadcast_to(O.MTIP(sd, dict(inv)).rp.radial_mask, (16, 16)), ps.radial_mask)        # the oracle's restatement
    if q_mask['type'] != 'none':
        assert not ps.radial_mask.all() and ps.radial_mask.any()


@needs_ref
def test_oracle_deg2_invariant_l2_diff_matches_reference():
    """oracle.deg2_invariant_l2_diff against the reference's generate_deg2_invariant_l2_diff (fxs_IO_methods.py:331-346,412-447)."""
    from oracle.sht import sh
    RH.import_reference(sh_class=sh)
    from xframe.projects.fxs.projectLibrary.fxs_IO_methods import generate_deg2_invariant_l2_diff
    rng = np.random.default_rng(4)
    n_r, l_max = 12, 6
    V = [rng.standard_normal((n_r, min(n_r, 2 * l + 1))) + 0j for l in range(l_max + 1)]
    V[3][:] = 0
    ref_inv = O.harmonic_coeff_to_deg2_invariants_3d(V)
    rmask = np.ones((l_max + 1, n_r), bool)
    rmask[:, :2] = False
    grid = np.zeros((n_r, 4, 8, 3))
    grid[..., 0] = np.linspace(0.01, 0.3, n_r)[:, None, None]

    class GP:
        reciprocalGrid = grid
    fn = generate_deg2_invariant_l2_diff(GP, deg2_invariants=ref_inv, used_orders={l: l for l in range(l_max + 1)}, n_particles=[2.0],
                                         invariant_mask=rmask[:, :, None] * rmask[:, None, :])
    Ilm = [rng.standard_normal((n_r, 2 * l + 1)) + 1j * rng.standard_normal((n_r, 2 * l + 1)) for l in range(l_max + 1)]
    want = fn(None, None, Ilm)
    got = O.deg2_invariant_l2_diff(ref_inv, rmask, 2.0, Ilm)
    assert want[3] == -1 and np.allclose(got, want, rtol=1e-13, atol=0)


@needs_ref
@pytest.mark.gpu
def test_reference_gpu_hankel_runs_on_the_cuda_layer():
    """generate_ht(..., use_gpu=True) of the reference, unmodified, through Multiprocessing.openCL_plugin.ClProcess and
    comm_module.add_gpu_process -- both routed to xframe_b200.gpu_access."""
    from xframe_b200.harmonic_transforms import sh
    RH.import_reference(sh_class=sh, cuda_gpu_layer=True)
    from xframe.projects.fxs.projectLibrary.hankel_transforms import generate_ht, generate_weightDict
    n_r, l_max = 16, 7
    w = generate_weightDict(l_max, n_r, reciprocity_coefficient=2.0, dimensions=3, mode='midpoint')
    orders = np.arange(l_max + 1)
    zht_g, izht_g = generate_ht(w['weights'], orders, 100.0, reciprocity_coefficient=2.0, dimensions=3, use_gpu=True, mode='midpoint')
    zht_c, izht_c = generate_ht(w['weights'], orders, 100.0, reciprocity_coefficient=2.0, dimensions=3, use_gpu=False, mode='midpoint')
    rng = np.random.default_rng(3)
    x = rng.standard_normal((n_r, (l_max + 1) ** 2)) + 1j * rng.standard_normal((n_r, (l_max + 1) ** 2))
    # CPU flavour works on m-ordered lists (hankel_transforms.py:642-658): compare through the direct layout
    from oracle.sht import sh as osh
    s = osh(l_max, n_phi=16, n_theta=8)
    cm = [x[:, idx] for idx in s.cplx_m_indices]
    for g_fn, c_fn in ((zht_g, zht_c), (izht_g, izht_c)):
        want = np.zeros_like(x)
        for mid, idx in enumerate(s.cplx_m_indices):
            want[:, idx] = c_fn([c.copy() for c in cm])[mid]
        assert rel_l2(g_fn(x.copy()), want) < 1e-12


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize('n_r', [8, 16])
def test_reference_loop_on_cuda_plugins_matches_project_worker(n_r):
    from xframe_b200.harmonic_transforms import sh
    from xframe_b200.worker import ProjectWorker
    sd = reference_test_settings(gpu=True, n_r=n_r)
    inv = synthetic_invariants(sd)
    RH.import_reference(sh_class=sh, cuda_gpu_layer=True)
    rho0 = initial_density(O.MTIP(sd, dict(inv)))
    rec, m = RH.make_mtip(sd, inv, rho0=rho0)
    ref = m.phasing_loop()                                   # reference loop; SHT and Hankel run in xframe_b200's CUDA library
    w = ProjectWorker(sd, dict(inv), n_reconstructions=1, initial_densities=[rho0])
    res, _ = w.run()
    got = res[0]
    assert w.plan.real_projections == ('support', 'value_threshold')          # 'assert_real' ignored like fxs_Projections.py:113-118
    assert np.allclose(got['error_dict']['main'], ref['error_dict']['main'], rtol=1e-6, atol=1e-300)
    if n_r > 8:
        assert min(ref['error_dict']['main']) > 0
    assert rel_l2(got['last_real_density'], ref['last_real_density']) < 1e-6
    assert rel_l2(got['real_density'], ref['real_density']) < 1e-6
    assert rel_l2(got['last_deg2_invariant'], ref['last_deg2_invariant']) < 1e-6
    assert got['loop_iterations'] == ref['loop_iterations']
    assert (got['support_mask'] != ref['support_mask']).mean() < 1e-3
    for k in ('real_density', 'reciprocal_density', 'initial_density', 'support_mask', 'last_deg2_invariant'):
        assert np.asarray(got[k]).shape == np.asarray(ref[k]).shape and np.asarray(got[k]).dtype == np.asarray(ref[k]).dtype, k
    for a, b in zip(got['fxs_unknowns'], ref['fxs_unknowns']):
        assert a.shape == b.shape


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize('variant', ['non_fxs', 'sw_center'])
def test_reference_sketch_tail_on_cuda_matches_project_worker(variant):
    """SW_center, HIO_non_FXS / ER_non_FXS and a finite best_density_not_in_first_n_iterations: the reference's loop (on the CUDA
    plugins) against the device-resident loop of xframe_b200."""
    from xframe_b200.harmonic_transforms import sh
    from xframe_b200.worker import ProjectWorker
    sd = sketch_tail_settings(gpu=True, variant=variant)
    inv = synthetic_invariants(sd)
    RH.import_reference(sh_class=sh, cuda_gpu_layer=True)
    rho0 = initial_density(O.MTIP(sd, dict(inv)))
    rec, m = RH.make_mtip(sd, inv, rho0=rho0)
    ref = m.phasing_loop()
    w = ProjectWorker(sd, dict(inv), n_reconstructions=2, initial_densities=[rho0, rho0 * 1.05])
    res, _ = w.run()
    got = res[0]
    assert np.allclose(got['error_dict']['main'], ref['error_dict']['main'], rtol=1e-6, atol=0)
    for k in ('last_real_density', 'real_density', 'last_reciprocal_density', 'reciprocal_density'):
        assert rel_l2(got[k], ref[k]) < 1e-6, k
    assert (got['support_mask'] != ref['support_mask']).mean() < 1e-3 and (got['last_support_mask'] != ref['last_support_mask']).mean() < 1e-3
    assert got['loop_iterations'] == ref['loop_iterations'] and abs(got['final_error'] - ref['final_error']) < 1e-6 * ref['final_error']
    w.plan.close()


@needs_ref
@pytest.mark.gpu
def test_reference_plugin_worker_saves_the_reference_record():
    """xframe_b200.reference_plugin.ProjectWorker(): the reference's no-argument constructor contract (settings.project /
    database.project globals, reconstruct.py:89-110), run() -> (result, locals()), post_processing -> db.save('reconstructions', record)
    with the layout the reference's integration test asserts (tests/test_fxs_integration.py:388-421)."""
    from xframe_b200.harmonic_transforms import sh
    RH.import_reference(sh_class=sh, cuda_gpu_layer=True)
    from xframe.library.pythonLibrary import DictNamespace
    from xframe import settings
    import xframe.database as database
    sd = reference_test_settings(gpu=True, n_r=8)
    sd['multi_process'] = {'use': True, 'n_parallel_reconstructions': 3}
    sd['GPU']['seed'] = 5
    inv = synthetic_invariants(sd)
    settings.project = DictNamespace.dict_to_dictnamespace(sd)
    database.project = RH.FakeDB(RH.fresh_data(inv))
    from xframe_b200.reference_plugin import ProjectWorker
    w = ProjectWorker()
    result, _ = w.run()
    assert result.dtype == object and len(result) == 3
    rec = database.project.saved['reconstructions']
    assert set(rec) == {'configuration', 'reconstruction_results', 'projection_matrices', 'stats'}
    assert set(rec['configuration']) == {'internal_grid', 'xray_wavelength', 'reciprocity_coefficient'}
    n_r, l_max = 8, 15
    shape = (n_r, 16, 32)
    assert rec['configuration']['internal_grid']['real_grid'].shape == shape + (3,)
    assert list(rec['reconstruction_results']) == [str(i) for i in np.argsort([r['error_dict']['main'][-1] for r in result])]
    n_it = len(result[0]['error_dict']['main'])
    for r in rec['reconstruction_results'].values():
        assert 'grid_pair' not in r and 'projection_matrices' not in r
        for k in ('real_density', 'last_real_density', 'reciprocal_density', 'last_reciprocal_density', 'initial_density'):
            assert r[k].shape == shape and r[k].dtype == complex
        for k in ('support_mask', 'last_support_mask', 'initial_support'):
            assert r[k].shape == shape and r[k].dtype == bool
        assert r['last_deg2_invariant'].shape == (l_max + 1, n_r, n_r) and r['last_deg2_invariant'].dtype == complex
        assert r['error_dict']['real']['l2_projection_diff'].shape == (n_it,) and r['n_particles'].shape == (n_it, 1)
        assert [u.shape for u in r['fxs_unknowns']] == [(min(2 * i + 1, 2 * n_r), 2 * i + 1) for i in range(l_max + 1)]   # data on 2 N_r q points
        assert isinstance(r['final_error'], float) and not np.isnan(r['real_density']).any()
    assert [p.shape[0] for p in rec['projection_matrices']] == [n_r] * (l_max + 1)
    w.worker.plan.close()

