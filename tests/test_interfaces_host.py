"""Host-only checks of the reference-facing interface mirrors (no GPU needed): index bookkeeping equals the
reference plugin's (restated in oracle/sht.py from shtns_plugin.py:105-114,269-274), weight assembly equals
hankel_transforms.py:426-452, and the device-only entry points fail loudly without a GPU."""
import numpy as np
import pytest
import torch

from oracle import mtip as O
from oracle import sht as OS


@pytest.mark.parametrize('l_max', [3, 15, 63])
def test_sh_index_bookkeeping(l_max):
    from xframe_b200.harmonic_transforms import sh
    a = sh(l_max)                                  # default sizes (the product needs n_phi >= 16)
    b = OS.sh(l_max, n_phi=a.n_phi, n_theta=a.n_theta)
    assert (a.n_theta, a.n_phi) == (b._sh.nlat, b._sh.nphi)
    assert np.array_equal(a.m, b.m) and np.array_equal(a.l, b.l) and a.n_coeff == b.n_coeff
    assert all(np.array_equal(x, y) for x, y in zip(a.cplx_m_indices, b.cplx_m_indices))
    assert a.cplx_l_indices == b.cplx_l_indices
    assert np.array_equal(a.cplx_l_split_indices, b.cplx_l_split_indices)
    assert np.array_equal(a.cplx_m_split_indices, b.cplx_m_split_indices)
    assert np.array_equal(a.cplx_m_indices_concat, b.cplx_m_indices_concat)
    assert np.allclose(a.theta, b.theta, atol=1e-15) and np.allclose(a.phi, b.phi, atol=1e-15)
    assert a.grid.shape == (a.n_theta, a.n_phi, 2)
    v = np.arange(a.n_coeff)
    assert np.array_equal(a.m_to_l_ordering(v[a.cplx_m_indices_concat]), v)     # shtns_plugin.py:240-247


def test_angular_size_helpers():
    from xframe_b200.harmonic_transforms import sh
    a = sh(63)
    assert a.n_angular_step_from_max_order(63) == {'n_phi': 256, 'n_theta': 128}     # shtns_plugin.py:94-101
    assert a.max_order_from_n_angular_steps(200) == 128 // 3


def test_weight_dict_and_assembly():
    from xframe_b200.hankel_transforms import generate_weightDict, assemble_weights
    for mode in ('midpoint', 'trapz', 'gauss'):
        wd = generate_weightDict(7, 12, reciprocity_coefficient=2.0, dimensions=3, mode=mode)
        assert wd['mode'] == mode and np.array_equal(wd['posHarmOrders'], np.arange(8))
        assert np.array_equal(wd['weights'], O.hankel_weights(7, 12, 2.0, mode))
        a, b = assemble_weights(wd['weights'], wd['posHarmOrders'], 33.0, 2.0, 3, mode), O.assemble_weights(wd['weights'], 33.0, 2.0, mode)
        assert np.allclose(a['forward'], b['forward'], rtol=1e-15) and np.allclose(a['inverse'], b['inverse'], rtol=1e-15)


def test_phase_split_of_complex_weights():
    from xframe_b200.gpu_access import _split_phase
    w = O.assemble_weights(O.hankel_weights(5, 8, 2.0, 'midpoint'), 10.0, 2.0)
    W = _split_phase(w['forward'], False)
    assert W is not None and W.shape == (6, 8, 8) and _split_phase(w['forward'] * (1 + 1j), False) is None
    assert np.allclose(np.moveaxis(W, 0, 2) * ((-1j) ** np.arange(6))[None, None, :], w['forward'], rtol=1e-15)
    assert _split_phase(w['inverse'], True) is not None


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU error path')
def test_no_cpu_fallback():
    from xframe_b200._lib import XfbError
    from xframe_b200.harmonic_transforms import sh, HarmonicTransform
    from xframe_b200.hankel_transforms import generate_ht
    with pytest.raises(XfbError):
        sh(7).forward_d(np.zeros((2, 8, 16), complex))
    with pytest.raises(XfbError):
        generate_ht(O.hankel_weights(3, 8, 2.0, 'midpoint'), np.arange(4), 10.0, use_gpu=False)
    with pytest.raises(XfbError):
        sh(7, mode_flag='real')
    ht = HarmonicTransform('complex', {'dimensions': 3, 'max_order': 7})
    assert set(ht.transforms_by_indices) == {'lm', 'ml', 'direct'} and ht.grid_param['thetas'].shape == (8,)


def test_result_record_round_trip(tmp_path):
    """results_io: the record of post_processing (reconstruct.py:160-183) written with the nested-dict rules of the reference's HDF5 saver
    (hdf5_plugin.py:53-131; .npz carrier here, h5py is absent from the image) and read back unchanged."""
    import numpy as np
    from xframe_b200 import results_io as RIO
    rng = np.random.default_rng(0)
    shape = (4, 8, 16)
    res = {str(i): {'real_density': rng.standard_normal(shape) + 1j * rng.standard_normal(shape), 'support_mask': rng.random(shape) > 0.5,
                    'final_error': float(i) * 0.1, 'loop_iterations': np.int64(7),
                    'error_dict': {'main': rng.random(5), 'real': {'l2_projection_diff': rng.random(5)}, 'reciprocal': {}},
                    'fxs_unknowns': tuple(rng.standard_normal((min(2 * l + 1, 4), 2 * l + 1)) + 0j for l in range(3)),
                    'n_particles_gradients': np.array([])} for i in (1, 0)}
    record = {'configuration': {'internal_grid': {'real_grid': rng.random(shape + (3,)), 'reciprocal_grid': rng.random(shape + (3,))},
                                'xray_wavelength': 1.23984, 'reciprocity_coefficient': 2.0},
              'reconstruction_results': res, 'projection_matrices': [rng.standard_normal((4, min(2 * l + 1, 4))) + 0j for l in range(3)],
              'stats': {'run_time': 1.5, 'structure_name': 'tutorial'}}
    path = RIO.save_reconstructions(record, str(tmp_path / 'run_0'))
    back = RIO.load_reconstructions(path)

    def same(a, b):
        if isinstance(a, dict):
            return isinstance(b, dict) and list(a) == list(b) and all(same(a[k], b[k]) for k in a)
        if isinstance(a, (list, tuple)):
            return type(a) is type(b) and len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
        if isinstance(a, str):
            return a == b
        return np.array_equal(np.asarray(a), np.asarray(b)) and np.asarray(a).dtype == np.asarray(b).dtype
    assert list(back['reconstruction_results']) == ['1', '0']            # ranking order is kept
    assert same(record, back)
