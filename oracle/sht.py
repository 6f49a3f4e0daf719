"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement of the spherical-harmonic transform that the reference obtains
from the third-party C library ``shtns`` (``pyproject.toml:35``, unpinned, not
vendored under /root/reference).  The class below follows the reference's own
plugin wrapper line by line for orderings and shapes
(``xframe/externalLibraries/shtns_plugin.py:11-274``) and restates the
published algorithm of shtns for the arithmetic:

  * orthonormal Y_l^m with Condon-Shortley phase, mmax = lmax, mres = 1
    (``shtns.sht(l_max)`` defaults, shtns_plugin.py:20),
  * Gauss-Legendre colatitude grid ordered north -> south,
    ``theta = arccos(cos_theta)`` (shtns_plugin.py:130-133),
  * phi_k = 2 pi k / n_phi (shtns_plugin.py:132),
  * complex coefficient index l*(l+1)+m (shtns_plugin.py:110-112),
  * spatial layout (n_theta, n_phi), phi contiguous (shtns_plugin.py:142,157).

PARITY STATUS AT THE shtns BOUNDARY: *unpinned* -- the reference ships no
numerical fixture for the SHT (tests/test_fxs_integration.py:16-29 checks shapes
only) and shtns itself cannot be installed here.  The restatement is anchored
on analytic known answers instead (scipy.special.sph_harm_y), see
tests/test_oracle_sht.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
import numpy as np
from scipy.special import roots_legendre


def default_angular_sizes(l_max, n_theta=0, n_phi=0):
    """Grid sizes used when the settings leave n_theta / n_phi at 0.

    The reference forwards 0 to shtns, which auto-sizes depending on its SIMD
    build (shtns_plugin.py:121-131, harmonic_transforms.py:65-66); that choice
    is not reproducible without shtns, so it is fixed here (and in the CUDA
    product) to: n_theta = L+1 rounded up to a multiple of 8,
    n_phi = next power of two >= 2L+2.  For L=63: 64 x 128.
    """
    if (not isinstance(n_theta, (int, np.integer))) or isinstance(n_theta, bool) or n_theta <= 0:
        n_theta = ((l_max + 1 + 7) // 8) * 8
    if (not isinstance(n_phi, (int, np.integer))) or isinstance(n_phi, bool) or n_phi <= 0:
        n_phi = 1
        while n_phi < 2 * l_max + 2:
            n_phi *= 2
    return int(n_theta), int(n_phi)


def gauss_grid(n_theta):
    """cos(theta) nodes north -> south and their Gauss weights."""
    x, w = roots_legendre(n_theta)
    return x[::-1].copy(), w[::-1].copy()


def normalized_legendre(l_max, x):
    """Orthonormal associated Legendre functions with Condon-Shortley phase.

    Returns P[m][l-m, j] = N_lm P_l^m(x_j) for 0<=m<=l<=l_max such that
    Y_l^m(theta,phi) = P[m][l-m] * exp(i m phi).  Stable three-term recurrence
    in l starting from the sectoral terms.
    """
    x = np.asarray(x, dtype=np.float64)
    s = np.sqrt(np.maximum(0.0, 1.0 - x * x))
    out = []
    pmm = np.full_like(x, np.sqrt(1.0 / (4.0 * np.pi)))
    for m in range(l_max + 1):
        if m > 0:
            pmm = -np.sqrt((2.0 * m + 1.0) / (2.0 * m)) * s * pmm
        tab = np.zeros((l_max - m + 1, x.size))
        tab[0] = pmm
        if m < l_max:
            tab[1] = np.sqrt(2.0 * m + 3.0) * x * pmm
        for l in range(m + 2, l_max + 1):
            a = np.sqrt((4.0 * l * l - 1.0) / (l * l - m * m))
            b = np.sqrt(((l - 1.0) ** 2 - m * m) / (4.0 * (l - 1.0) ** 2 - 1.0))
            tab[l - m] = a * (x * tab[l - m - 1] - b * tab[l - m - 2])
        out.append(tab)
    return out


class ShtCore:
    """Stand-in for the ``shtns.sht`` object: analys_cplx / synth_cplx on one shell."""

    def __init__(self, l_max, n_theta, n_phi):
        self.lmax = l_max
        self.mres = 1
        self.nlat = n_theta
        self.nphi = n_phi
        self.spat_shape = (n_theta, n_phi)
        self.cos_theta, self.gauss_w = gauss_grid(n_theta)
        assert n_theta > l_max and n_phi > 2 * l_max, "grid too small for exact quadrature"
        self.P = normalized_legendre(l_max, self.cos_theta)  # list over m>=0 of [l-m, j]
        self.l = np.concatenate([np.full(2 * l + 1, l) for l in range(l_max + 1)])

    # batched versions (leading axis = shells); the per-shell API maps onto them
    def analys_batch(self, f):
        """f: [S, n_theta, n_phi] complex -> [S, (L+1)^2] complex."""
        L, nphi = self.lmax, self.nphi
        fm = np.fft.fft(f, axis=-1) * (2.0 * np.pi / nphi)  # [S, j, mi], e^{-i m phi}
        out = np.zeros((f.shape[0], (L + 1) ** 2), dtype=complex)
        ls = np.arange(L + 1)
        for m in range(L + 1):
            Pw = self.P[m] * self.gauss_w[None, :]  # [l-m, j]
            lm = ls[m:] * (ls[m:] + 1)
            out[:, lm + m] = fm[:, :, m] @ Pw.T
            if m > 0:
                out[:, lm - m] = ((-1) ** m) * (fm[:, :, nphi - m] @ Pw.T)
        return out

    def synth_batch(self, c):
        """c: [S, (L+1)^2] complex -> [S, n_theta, n_phi] complex."""
        L, nphi = self.lmax, self.nphi
        fm = np.zeros((c.shape[0], self.nlat, nphi), dtype=complex)
        ls = np.arange(L + 1)
        for m in range(L + 1):
            P = self.P[m]
            lm = ls[m:] * (ls[m:] + 1)
            fm[:, :, m] = c[:, lm + m] @ P
            if m > 0:
                fm[:, :, nphi - m] = ((-1) ** m) * (c[:, lm - m] @ P)
        return np.fft.ifft(fm, axis=-1) * nphi

    def analys_cplx(self, shell):
        return self.analys_batch(np.asarray(shell, dtype=complex)[None])[0]

    def synth_cplx(self, coeff):
        return self.synth_batch(np.asarray(coeff, dtype=complex)[None])[0]


class sh:
    """Same surface as the reference plugin class ``sh`` (shtns_plugin.py:11-274).

    Injectable at ``xframe.library.mathLibrary.shtns`` (startup_routines.py:57).
    """

    def __init__(self, l_max, mode_flag='complex', output_order='l', anti_aliazing_degree=2,
                 n_phi=False, n_theta=False):
        l_max = int(l_max)
        self.l_max = l_max
        self.anti_aliazing_degree = anti_aliazing_degree
        self.n_coeff = (l_max + 1) ** 2
        nt, nph = default_angular_sizes(l_max, n_theta, n_phi)
        self._sh = ShtCore(l_max, nt, nph)
        self._phi = 2 * np.pi * np.arange(nph) / nph            # shtns_plugin.py:132
        self._theta = np.arccos(self._sh.cos_theta)             # shtns_plugin.py:133
        self.mode = mode_flag
        # shtns_plugin.py:105-114
        ls = np.arange(l_max + 1, dtype=int)
        ms = np.concatenate((ls, -ls[:0:-1]))
        self.m, self.l = ms, ls
        self.cplx_m_indices = [ls[np.abs(m):] * (ls[np.abs(m):] + 1) + m for m in ms]
        self.cplx_l_indices = [slice(l ** 2, l ** 2 + 2 * l + 1) for l in range(l_max + 1)]
        self.cplx_m_indices_concat = np.concatenate(self.cplx_m_indices)
        self.cplx_l_split_indices = np.arange(1, l_max + 1) ** 2
        lp1 = np.arange(l_max + 2)
        index = (lp1 * (lp1 + 1) / 2).astype(int)            # shtns_plugin.py:269-274
        self.cplx_m_split_indices = np.concatenate((index[-1] - index[-2::-1], index[-1] + index[1:-2]))
        if mode_flag != 'complex':
            raise NotImplementedError("oracle restates the complex transform only "
                                      "(the 3-D fxs path uses mode 'complex', reconstruct.py:345-350)")

    phi = property(lambda self: self._phi)
    theta = property(lambda self: self._theta)

    # --- analysis: shtns_plugin.py:151-176,218-229 ---
    def _analysis(self, data):
        shape = data.shape[:-2]
        val = self._sh.analys_batch(np.asarray(data, dtype=complex).reshape(-1, *self._sh.spat_shape))
        return np.moveaxis(val.reshape(*shape, -1), -1, 0)      # [(L+1)^2, ...]

    def forward_l(self, data):
        val = self._analysis(data)
        return [np.moveaxis(val[idx], 0, -1) for idx in self.cplx_l_indices]

    def forward_m(self, data):
        val = self._analysis(data)
        return [np.moveaxis(val[idx], 0, -1) for idx in self.cplx_m_indices]

    def forward_d(self, data):                                   # shtns_plugin.py:250-255
        return self._sh.analys_batch(np.asarray(data, dtype=complex))

    # --- synthesis: shtns_plugin.py:179-194,230-238,257-261 ---
    def inverse_l(self, data):
        return self._sh.synth_batch(np.concatenate(data, axis=1))

    def inverse_m(self, data):
        r_shape = data[0].shape[:-1]
        full = np.zeros(r_shape + (self.n_coeff,), dtype=complex)
        for m_id, index in enumerate(self.cplx_m_indices):
            full[..., index] = data[m_id]
        return self._sh.synth_batch(full)

    def inverse_d(self, data):
        return self._sh.synth_batch(np.asarray(data, dtype=complex))

    def test(self, data):                                        # shtns_plugin.py:263-267
        return self._sh.synth_batch(self._sh.analys_batch(data + 0.j))
