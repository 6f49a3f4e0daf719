"""Anchors oracle/sht.py (the restated shtns) on analytic known answers -- the shtns boundary has no fixture in
the reference (tests/test_fxs_integration.py:16-29 checks shapes only)."""
import numpy as np
import pytest
from scipy.special import sph_harm_y

from oracle.sht import sh, default_angular_sizes


@pytest.mark.parametrize('L', [7, 15, 31])
def test_unit_coefficient_is_sph_harm(L):
    s = sh(L)
    core = s._sh
    T, P = np.meshgrid(s.theta, s.phi, indexing='ij')
    rng = np.random.default_rng(L)
    for _ in range(12):
        l = int(rng.integers(0, L + 1))
        m = int(rng.integers(-l, l + 1))
        c = np.zeros((L + 1) ** 2, complex)
        c[l * (l + 1) + m] = 1
        ref = sph_harm_y(l, m, T, P)
        assert np.abs(core.synth_cplx(c) - ref).max() < 1e-12
        assert np.abs(core.analys_cplx(ref) - c).max() < 1e-12


@pytest.mark.parametrize('L', [15, 63])
def test_roundtrip_and_hermitian_symmetry(L):
    s = sh(L)
    core = s._sh
    rng = np.random.default_rng(1)
    c = rng.normal(size=(3, (L + 1) ** 2)) + 1j * rng.normal(size=(3, (L + 1) ** 2))
    assert np.abs(core.analys_batch(core.synth_batch(c)) - c).max() < 1e-11       # sh.test, shtns_plugin.py:263-267
    f = rng.normal(size=(1,) + core.spat_shape)
    cr = core.analys_batch(f)[0]
    for l in range(L + 1):                                                        # tests/old/fxs_lib/fourier.py:98-101
        for m in range(1, l + 1):
            assert abs(cr[l * (l + 1) - m] - (-1) ** m * np.conj(cr[l * (l + 1) + m])) < 1e-13


def test_orderings_match_plugin_layout():
    L = 5
    s = sh(L)
    rng = np.random.default_rng(2)
    x = rng.normal(size=(4, *s._sh.spat_shape)) + 0j
    d = s.forward_d(x)
    fl = s.forward_l(x)
    fm = s.forward_m(x)
    assert [a.shape for a in fl] == [(4, 2 * l + 1) for l in range(L + 1)]
    assert np.allclose(np.concatenate(fl, axis=1), d)
    assert [a.shape for a in fm] == [(4, L - abs(m) + 1) for m in s.m]
    for mid, idx in enumerate(s.cplx_m_indices):
        assert np.allclose(fm[mid], d[:, idx])
    assert np.allclose(s.inverse_m(fm), s.inverse_l(fl))
    assert default_angular_sizes(63) == (64, 128)
