"""Host-side constant tables for the xfb200 plan (one-off setup, numpy/scipy).

Every formula cites the reference line it reproduces (paths relative to
/root/reference/xframe).  The Legendre tables stand in for what the third-party
``shtns`` library builds internally (orthonormal, Condon-Shortley phase, Gauss
grid; shtns_plugin.py:20,130-133) -- computed here in extended precision and
rounded once to float64.
"""
import numpy as np
from scipy.special import jv, roots_legendre, spherical_jn


def default_angular_sizes(l_max, n_theta=0, n_phi=0):
    """Angular grid when settings leave grid.n_theta / grid.n_phi at 0 (default_0.01.yaml:14-19).

    The reference lets shtns auto-size (harmonic_transforms.py:65-66), which depends on the
    shtns build; here it is fixed to n_theta = L+1 rounded up to a multiple of 8 and
    n_phi = next power of two >= 2L+2 (L=63 -> 64 x 128).
    """
    def unset(v):
        return (not isinstance(v, (int, np.integer))) or isinstance(v, bool) or v <= 0
    if unset(n_theta):
        n_theta = ((l_max + 1 + 7) // 8) * 8
    if unset(n_phi):
        n_phi = 16
        while n_phi < 2 * l_max + 2:
            n_phi *= 2
    return int(n_theta), int(n_phi)


def gauss_grid(n_theta):
    """cos(theta_j) north -> south (shtns_plugin.py:133) and Gauss weights."""
    x, w = roots_legendre(n_theta)
    return x[::-1].copy(), w[::-1].copy()


def normalized_legendre(l_max, x):
    """P[m][l-m, j] = N_lm P_l^m(x_j), orthonormal with Condon-Shortley phase; long double recurrence."""
    ld = np.longdouble
    x = np.asarray(x, dtype=ld)
    s = np.sqrt(np.maximum(ld(0), ld(1) - x * x))
    out = []
    pmm = np.full_like(x, np.sqrt(ld(1) / (ld(4) * ld(np.pi))))
    pmm[:] = np.sqrt(ld(1) / (ld(4) * np.arccos(ld(-1))))
    for m in range(l_max + 1):
        if m > 0:
            pmm = -np.sqrt(ld(2 * m + 1) / ld(2 * m)) * s * pmm
        tab = np.zeros((l_max - m + 1, x.size), dtype=ld)
        tab[0] = pmm
        if m < l_max:
            tab[1] = np.sqrt(ld(2 * m + 3)) * x * pmm
        for l in range(m + 2, l_max + 1):
            a = np.sqrt(ld(4 * l * l - 1) / ld(l * l - m * m))
            b = np.sqrt(ld((l - 1) ** 2 - m * m) / ld(4 * (l - 1) ** 2 - 1))
            tab[l - m] = a * (x * tab[l - m - 1] - b * tab[l - m - 2])
        out.append(tab)
    return out


def pack_legendre(l_max, n_theta, n_phi):
    """Tables consumed by csrc/legendre.cuh: concat(FE, FO, IE, IO).

    K2 = n_theta/2 northern nodes, NP = same-parity degree count rounded up to 8.
      FE[m][j][i] = w_j (2 pi / n_phi) P_{m+2i}^m(x_j)      FO: l = m+1+2i   (analysis)
      IE[m][i][j] = P_{m+2i}^m(x_j)                          IO: l = m+1+2i   (synthesis)
    The southern hemisphere follows from P_l^m(-x) = (-1)^(l+m) P_l^m(x).
    """
    assert n_theta % 8 == 0 and n_theta > l_max
    x, w = gauss_grid(n_theta)
    K2 = n_theta // 2
    NP = ((l_max // 2 + 1) + 7) // 8 * 8
    P = normalized_legendre(l_max, x[:K2])
    wl = np.asarray(w[:K2], dtype=np.longdouble) * (2 * np.arccos(np.longdouble(-1)) / n_phi)
    FE = np.zeros((l_max + 1, K2, NP))
    FO = np.zeros_like(FE)
    IE = np.zeros((l_max + 1, NP, K2))
    IO = np.zeros_like(IE)
    for m in range(l_max + 1):
        pe = P[m][0::2]      # l = m, m+2, ...
        po = P[m][1::2]
        IE[m, :pe.shape[0], :] = pe.astype(np.float64)
        IO[m, :po.shape[0], :] = po.astype(np.float64)
        FE[m, :, :pe.shape[0]] = (pe * wl[None, :]).T.astype(np.float64)
        FO[m, :, :po.shape[0]] = (po * wl[None, :]).T.astype(np.float64)
    return np.concatenate([FE.ravel(), FO.ravel(), IE.ravel(), IO.ravel()]), NP


def radial_grids(ft_type, q_max, n_r, rc):
    """ft_grid_pairs.py:274-291 with r_max = rc*N/q_max (mathLibrary.py:1169-1176)."""
    r_max = rc * n_r / q_max
    if ft_type == 'midpoint':
        dr, dq = r_max / n_r, q_max / n_r
        rs = np.linspace(dr / 2, r_max - dr / 2, num=n_r, endpoint=True)
        qs = np.linspace(dq / 2, q_max - dq / 2, num=n_r, endpoint=True)
    elif ft_type in ('trapz', 'Zernike'):
        rs = np.linspace(0, r_max, n_r, endpoint=True)
        qs = np.linspace(0, q_max, n_r, endpoint=True)
    elif ft_type == 'gauss':                         # Gauss-Legendre nodes mapped to [0, r_max] (ft_grid_pairs.py:293-300)
        xs = roots_legendre(n_r)[0]
        rs = r_max / 2 * xs + r_max / 2
        qs = q_max / 2 * xs + q_max / 2
    else:
        raise ValueError(f"fourier_transform.type '{ft_type}' is not supported by xframe_b200 (midpoint, trapz, gauss)")
    return rs, qs


def hankel_weights(l_max, n_r, rc, mode='midpoint'):
    """w[l,p,k] = p^2 j_l(p k rc / N); p is summed. hankel_transforms.py:399-410 (midpoint), :322-333 (trapz);
    gauss (:477-490): p, k = Gauss-Legendre nodes on [0, 2], w = p^2 j_l(p k rc N / 4) w_gauss(p)."""
    ls = np.arange(l_max + 1)
    if mode == 'gauss':
        xi, wg = roots_legendre(n_r)
        ps = ks = xi + 1
        arg = ks[None, :] * ps[:, None] * rc * n_r / 4
        return np.ascontiguousarray(ps[None, :, None] ** 2 * spherical_jn(ls[:, None, None], arg[None, :, :]) * wg[None, :, None])
    if mode == 'midpoint':
        ps = np.arange(n_r) + 0.5
        ks = np.arange(n_r) + 0.5
    elif mode == 'trapz':
        ps = np.arange(1, n_r)
        ks = np.arange(n_r)
    else:
        raise ValueError(f"hankel mode '{mode}' is not supported by xframe_b200 (midpoint, trapz, gauss)")
    arg = ks[None, :] * ps[:, None] * rc / n_r
    return np.ascontiguousarray(ps[None, :, None] ** 2 * spherical_jn(ls[:, None, None], arg[None, :, :]))


def hankel_scales(r_max_grid, n_r, rc, mode='midpoint'):
    """(r_max/N)^3 sqrt(2/pi) and (q_max/N)^3 sqrt(2/pi) with q_max = rc N / r_max; r_max is the LARGEST GRID
    POINT as the reference passes it (reconstruct.py:329, hankel_transforms.py:432-445).  gauss: r_max/2 and q_max/2
    instead of the uniform steps (:523-531)."""
    q_max = rc * n_r / r_max_grid
    c = np.sqrt(2 / np.pi)
    div = 2 if mode == 'gauss' else n_r
    return (r_max_grid / div) ** 3 * c, (q_max / div) ** 3 * c


def integration_weights(rs, n_theta):
    """Weights of SphericalIntegrator.integrate (mathLibrary.py:1223-1235): trapz over r of
    r^2 * (pi/n_theta) * sum_theta w_theta * sum_phi f  ->  wt[r,theta] (phi weight 1)."""
    w = roots_legendre(n_theta)[1]
    rs = np.asarray(rs, dtype=np.float64)
    tw = np.zeros_like(rs)
    d = np.diff(rs)
    tw[:-1] += d / 2
    tw[1:] += d / 2
    return np.ascontiguousarray((tw * rs ** 2)[:, None] * (np.pi / n_theta) * w[None, :])


# ---------------------------------------------------------------------------------- 2-D (polar) tables
def polar_hankel_weights(m_max, n_r, rc, mode='midpoint'):
    """w[m,p,k] = p J_m(p k rc / N), m = 0..M; p is summed.  hankel_transforms.py:412-424 (midpoint), :335-347 (trapz)."""
    ms = np.arange(m_max + 1)
    if mode == 'gauss':                              # :492-503
        xi, wg = roots_legendre(n_r)
        ps = ks = xi + 1
        arg = ks[None, :] * ps[:, None] * rc * n_r / 4
        return np.ascontiguousarray(ps[None, :, None] * jv(ms[:, None, None], arg[None, :, :]) * wg[None, :, None])
    if mode == 'midpoint':
        ps = np.arange(n_r) + 0.5
        ks = np.arange(n_r) + 0.5
    elif mode == 'trapz':
        ps = np.arange(1, n_r)
        ks = np.arange(n_r)
    else:
        raise ValueError(f"hankel mode '{mode}' is not supported by xframe_b200 (midpoint, trapz, gauss)")
    arg = ks[None, :] * ps[:, None] * rc / n_r
    return np.ascontiguousarray(ps[None, :, None] * jv(ms[:, None, None], arg[None, :, :]))


def polar_hankel_device_weights(w):
    """[M+1, p, k] -> [2M+1, p, k] in DFT-index order (0..M, -M..-1) with w_{-m} = (-1)^m w_m
    (J_{-m} = (-1)^m J_m; hankel_transforms.py:441)."""
    m = np.arange(w.shape[0])
    neg = ((-1.0) ** m[:0:-1])[:, None, None] * w[:0:-1]
    return np.ascontiguousarray(np.concatenate((w, neg), axis=0))


def polar_hankel_scales(r_max_grid, n_r, rc, mode='midpoint'):
    """(r_max/N)^2 and (q_max/N)^2, q_max = rc N / r_max (hankel_transforms.py:433-439); gauss: r_max/2, q_max/2 (:517-519)."""
    q_max = rc * n_r / r_max_grid
    div = 2 if mode == 'gauss' else n_r
    return (r_max_grid / div) ** 2, (q_max / div) ** 2


def polar_integration_weights(rs, phis):
    """Weights of PolarIntegrator.integrate (mathLibrary.py:1254-1262): trapz over phi (NOT periodic: end points get
    half weight) then trapz over r of r * (...)  ->  wt[r, phi]."""
    def trapz_w(x):
        x = np.asarray(x, dtype=np.float64)
        w = np.zeros_like(x)
        d = np.diff(x)
        w[:-1] += d / 2
        w[1:] += d / 2
        return w
    rs = np.asarray(rs, dtype=np.float64)
    return np.ascontiguousarray((trapz_w(rs) * rs)[:, None] * trapz_w(phis)[None, :])
