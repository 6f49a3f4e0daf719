"""Parity of the reference-facing interface mirrors (harmonic transform plugin, HarmonicTransform, Hankel factory,
Fourier composer, GPU-access layer, fxs_unknowns) with the oracle.  All calls go through the C-ABI library.
Tolerances: relative L2 <= 1e-12 for transforms (FP64 path); unknowns see the test."""
import numpy as np
import pytest
import torch

from helpers import load_golden, golden_settings, golden_data, rel_l2
from oracle import mtip as O
from oracle import sht as OS

pytestmark = pytest.mark.gpu
L, NT, NP, NR = 15, 16, 32, 24


@pytest.fixture(scope='module')
def pair():
    from xframe_b200.harmonic_transforms import sh
    return sh(L, n_phi=NP, n_theta=NT), OS.sh(L, n_phi=NP, n_theta=NT)


def _field(rng, shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def test_sh_plugin_all_orderings(pair):
    ours, ref = pair
    x = _field(np.random.default_rng(1), (NR, NT, NP))
    assert rel_l2(ours.forward_d(x), ref.forward_d(x)) < 1e-12
    for a, b in zip(ours.forward_l(x), ref.forward_l(x)):
        assert a.shape == b.shape and rel_l2(a, b) < 1e-12
    for a, b in zip(ours.forward_m(x), ref.forward_m(x)):
        assert a.shape == b.shape and rel_l2(a, b) < 1e-12
    cl, cm, cd = ref.forward_l(x), ref.forward_m(x), ref.forward_d(x)
    want = ref.inverse_d(cd)
    assert rel_l2(ours.inverse_d(cd), want) < 1e-12
    assert rel_l2(ours.inverse_l(cl), want) < 1e-12
    assert rel_l2(ours.inverse_m(cm), want) < 1e-12
    assert rel_l2(ours.test(want.real), ref.test(want.real)) < 1e-12          # shtns_plugin.py:263-267
    # leading batch axes and device tensors stay on the device
    xb = torch.from_numpy(np.stack([x, 2 * x])).cuda()
    cb = ours.forward_d(xb)
    assert cb.is_cuda and cb.shape == (2, NR, (L + 1) ** 2) and rel_l2(cb[1].cpu().numpy(), 2 * cd) < 1e-12


def test_sh_plugin_chunks_large_inputs(pair):
    ours, ref = pair
    ours._get_plan()
    cap, ours._cap = ours._cap, 16                                              # force several device calls
    try:
        x = _field(np.random.default_rng(2), (40, NT, NP))
        assert rel_l2(ours.forward_d(x), ref.forward_d(x)) < 1e-12
    finally:
        ours._cap = cap


def test_harmonic_transform_class():
    from xframe_b200.harmonic_transforms import HarmonicTransform
    ht = HarmonicTransform('complex', {'dimensions': 3, 'max_order': L, 'n_phi': NP, 'n_theta': NT, 'indices': 'lm'})
    ref = OS.sh(L, n_phi=NP, n_theta=NT)
    assert ht.max_order == L and ht.n_coeff == (L + 1) ** 2
    assert np.allclose(ht.grid_param['thetas'], ref.theta) and np.allclose(ht.grid_param['phis'], ref.phi)
    assert set(ht.transforms_by_indices) == {'lm', 'ml', 'direct'}
    x = _field(np.random.default_rng(3), (5, NT, NP))
    c = ht.forward(x)
    assert len(c) == L + 1 and c[3].shape == (5, 7)
    assert rel_l2(ht.inverse(c), ref.inverse_l(ref.forward_l(x))) < 1e-12
    ht2 = HarmonicTransform.from_data_array('complex', np.zeros((4, NT, NP)))    # harmonic_transforms.py:23-31
    assert ht2.max_order == NT - 1


@pytest.mark.parametrize('mode', ['midpoint', 'trapz', 'gauss'])
def test_generate_ht_and_ft(mode):
    from xframe_b200.harmonic_transforms import HarmonicTransform
    from xframe_b200.hankel_transforms import generate_weightDict, generate_ht
    from xframe_b200.fourier_transforms import generate_ft
    rc, max_q = 2.0, 0.4
    rs, qs = O.radial_grids(mode, max_q, NR, rc)
    r_max = float(np.max(rs))
    wd = generate_weightDict(L, NR, reciprocity_coefficient=rc, dimensions=3, mode=mode)
    assert np.array_equal(wd['weights'], O.hankel_weights(L, NR, rc, mode))
    zht, izht = generate_ht(wd['weights'], wd['posHarmOrders'], r_max, reciprocity_coefficient=rc, dimensions=3, use_gpu=True, mode=mode)
    w = O.assemble_weights(wd['weights'], r_max, rc, mode)
    fwd, inv = O.generate_spherical_ht_direct(w, L, mode)
    c = _field(np.random.default_rng(4), (NR, (L + 1) ** 2))
    assert rel_l2(zht(c), fwd(c)) < 1e-12 and rel_l2(izht(c), inv(c)) < 1e-12
    ht = HarmonicTransform('complex', {'dimensions': 3, 'max_order': L, 'n_phi': NP, 'n_theta': NT})
    ft, ift = generate_ft(r_max, wd, ht, 3, use_gpu=True, reciprocity_coefficient=rc, mode=mode)
    osh = OS.sh(L, n_phi=NP, n_theta=NT)
    oft, oift = O.generate_ft(osh, wd['weights'], r_max, rc, L, mode, 'direct')
    x = osh.inverse_d(_field(np.random.default_rng(5), (NR, (L + 1) ** 2)))
    assert rel_l2(ft(x), oft(x)) < 1e-12 and rel_l2(ift(x), oift(x)) < 1e-12
    assert rel_l2(ht.forward(x)[2], osh.forward_l(x)[2]) < 1e-12                 # the transform object shares the plan


def test_assemble_weights_layout():
    from xframe_b200.hankel_transforms import assemble_weights
    w = O.hankel_weights(L, NR, 2.0, 'midpoint')
    a, b = assemble_weights(w, np.arange(L + 1), 50.0, 2.0, 3, 'midpoint'), O.assemble_weights(w, 50.0, 2.0)
    assert a['forward'].shape == (NR, NR, L + 1)
    assert np.allclose(a['forward'], b['forward'], rtol=1e-15) and np.allclose(a['inverse'], b['inverse'], rtol=1e-15)


def test_gpu_access_layer_runs_the_hankel_process():
    """The kernel_dict the reference builds in generate_spherical_ht_gpu (hankel_transforms.py:742-766) runs unchanged."""
    from xframe_b200 import gpu_access as GA
    from xframe_b200._lib import XfbError
    assert GA.get_number_of_gpus() >= 1 and GA.CudaPlugin.get_number_of_gpus() == GA.get_number_of_gpus()
    w = O.assemble_weights(O.hankel_weights(L, NR, 2.0, 'midpoint'), 50.0, 2.0)
    nlm = (L + 1) ** 2
    for key in ('forward', 'inverse'):
        kd = {'kernel': '__kernel void apply_weights(...) {...}', 'name': key + '_hankel',
              'functions': ({'name': 'apply_weights', 'dtypes': (complex, complex, complex, np.int64, np.int64, np.int64),
                             'shapes': ((NR, nlm), w[key].shape, (NR, nlm), None, None, None),
                             'arg_roles': ('output', 'const_input', 'input', 'const_input', 'const_input', 'const_input'),
                             'const_inputs': (None, w[key], None, np.int64(NR), np.int64(nlm), np.int64(L + 1)),
                             'global_range': (NR, nlm), 'local_range': None},)}
        proc = GA.openCL_plugin.ClProcess(kd)
        assert proc.n_inputs == 1 and proc.n_outputs == 1 and proc.input_shapes == [(NR, nlm)]
        fn = GA.comm_module.add_gpu_process(proc)
        rho = _field(np.random.default_rng(6), (NR, nlm))
        ls = np.floor(np.sqrt(np.arange(nlm))).astype(int)
        want = np.einsum('qij,qj->ij', w[key][:, :, ls], rho)                   # the OpenCL kernel's sum (:684-694)
        assert rel_l2(fn(rho), want) < 1e-12
    kd['functions'][0]['name'] = 'matvec'
    with pytest.raises(XfbError):
        GA.ClProcess(kd)


def test_gpu_access_layer_runs_the_apply_matrix_demo():
    """The GPU demo of the reference's framework test (tests/test_framework_integration.py:230-309): its kernel_dict for `apply_matrix`,
    verbatim, through ClProcess / add_gpu_process; the reference asserts (result == matrix @ vects).all() for 10 x 10 times 10 x 5."""
    from xframe_b200 import gpu_access as GA
    rng = np.random.default_rng(12)
    nq, nvec = 10, 5
    matrix, vects = rng.random((nq, nq)), rng.random((nq, nvec))
    kd = {'kernel': '__kernel void apply_matrix(...) {...}', 'name': 'gpu_func',
          'functions': ({'name': 'apply_matrix', 'dtypes': (float, float, float, np.int64, np.int64),
                         'shapes': ((nq, nvec), matrix.shape, (nq, nvec), None, None, None),
                         'arg_roles': ('output', 'const_input', 'input', 'const_input', 'const_input'),
                         'const_inputs': (None, matrix, None, np.int64(nq), np.int64(nvec)),
                         'global_range': (nq, nvec), 'local_range': None},)}
    fn = GA.comm_module.add_gpu_process(GA.openCL_plugin.ClProcess(kd))
    result = fn(vects)
    expected = matrix @ vects
    assert result.shape == expected.shape and result.dtype == np.float64
    assert np.allclose(result, expected, rtol=4e-16, atol=0)            # q ascending, one FMA per term: at most an ulp from BLAS
    # a larger case against a sequential sum
    nq, nvec = 64, 7
    matrix, vects = rng.standard_normal((nq, nq)), rng.standard_normal((nq, nvec))
    kd['functions'][0].update(shapes=((nq, nvec), matrix.shape, (nq, nvec), None, None, None), const_inputs=(None, matrix, None, np.int64(nq), np.int64(nvec)))
    fn = GA.comm_module.add_gpu_process(GA.openCL_plugin.ClProcess(kd))
    assert rel_l2(fn(vects), matrix @ vects) < 1e-15


def test_unknowns_match_the_reference_formula():
    """fxs_unknowns (approximate_unknowns, fxs_Projections.py:752-767).  Only V_l unk_l is observable; unk_l itself is
    compared where PD_l I_l is well conditioned, and must be a partial isometry everywhere."""
    from xframe_b200.plan import Plan
    g = load_golden('ref_medium_ops')
    sd = golden_settings(g)
    m = O.MTIP(sd, golden_data(g))
    plan = Plan(m.l_max, len(m.rs), float(g['max_q']), n_theta=int(g['n_theta']), n_phi=int(g['n_phi']), max_batch=2)
    plan.set_projection(m.rp.projection_matrices, m.rp.radial_mask, m.rp.number_of_particles[0])
    I = g['I_direct']
    rng = np.random.default_rng(7)
    I2 = I * (1 + 0.1 * rng.standard_normal(I.shape[0]))[:, None]
    plan.project_invariants(torch.from_numpy(np.stack([I, I2])).cuda())
    splits = np.arange(1, m.l_max + 1) ** 2
    for run, Id in enumerate((I, I2)):
        unk = plan.unknowns(run)
        ref = m.rp.approximate_unknowns(np.split(Id, splits, axis=1))
        assert len(unk) == len(ref)
        for l, (u, r) in enumerate(zip(unk, ref)):
            assert u.shape == r.shape and u.dtype == np.complex128
            V = m.rp.projection_matrices[l]
            assert rel_l2(V @ u, V @ r) < 1e-6
            s = np.linalg.svd(m.rp.PDs[l] @ np.split(Id, splits, axis=1)[l], compute_uv=False)
            if l % 2 == 0 and s[-1] > 1e-6 * s[0]:
                assert rel_l2(u, r) < 1e-8
            if l % 2 == 0:
                p = u @ u.conj().T                                               # partial isometry: projector
                assert rel_l2(p @ p, p) < 1e-9
    plan.close()
