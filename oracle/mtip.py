"""ORACLE (test infrastructure, never shipped on the product path).

numpy restatement of the reference's MTIP phasing path for ``xframe fxs
reconstruct`` (3-D).  Every function cites the reference file:line it follows
(paths relative to /root/reference/xframe).  It mirrors the reference's numpy
formulation (broadcast-multiply-sum Hankel, LAPACK SVD Procrustes, boolean-index
elementwise code) so that timing it on host cores is representative of the
reference CPU path (``GPU.use: False``); the SHT stage comes from
``oracle/sht.py`` (shtns restated -- the one piece the reference does not own).

Pinned against the reference itself: ``tests/golden/make_golden.py`` imports the
unmodified reference from /root/reference in the build container, injects
``oracle.sht.sh`` at the reference's own plugin slot and dumps operator- and
loop-level golden vectors to ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays them against this file.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
import numpy as np
from scipy.special import spherical_jn, roots_legendre

from .sht import sh as OracleSH

# --------------------------------------------------------------------------
# ramps -- library/mathLibrary.py:1033-1129
# --------------------------------------------------------------------------


class ExponentialRamp:
    def __init__(self, start, stop, exponent, stop_argument=1):
        self.start, self.stop, self.stop_argument = start, stop, stop_argument
        if stop < start:
            exponent *= exponent / abs(exponent) * -1
        else:
            exponent *= exponent / abs(exponent)
        self.exponent = exponent
        self.A = (start - stop) / (1 - np.exp(exponent * stop_argument))
        self.B = start - self.A

    def eval(self, x):
        v = self.A * np.exp(x * self.exponent) + self.B
        return np.maximum(v, self.stop) if self.start > self.stop else np.minimum(v, self.stop)

    __call__ = eval


def _is_number(v):
    return np.issubdtype(np.array(v).dtype, np.number)


class LinearRamp:
    def __init__(self, start, stop=False, slope=False, default_start=False, default_stop=False):
        self.default_stop, self.default_start = default_stop, default_start
        self.start = (start, 0) if not isinstance(start, (list, tuple)) else start
        self.undefined = False
        if not _is_number(self.start[0]):
            if default_start == False:  # noqa: E712  (sic, mathLibrary.py:1069)
                self.undefined = True
            else:
                self.start = (default_start, 0)
        self.stop, self.stop_is_defined = self._parse_stop(stop)
        self.slope_is_defined = not isinstance(slope, bool)
        self.slope = slope
        if not self.undefined:
            self._set()

    def _parse_stop(self, stop):
        ok = False
        if isinstance(stop, (list, tuple)):
            stop = list(stop)
            val_num = _is_number(stop[0])
            if not val_num and _is_number(self.default_stop):
                stop[0] = self.default_stop
                val_num = True
            if _is_number(stop[1]) and val_num and stop[1] >= self.start[1]:
                ok = True
        return (stop if ok else False), ok

    def _set(self):
        start, stop, slope = self.start, self.stop, self.slope
        if (not self.stop_is_defined) and (not self.slope_is_defined):
            self.A, self.B = 0, start[0]
            return
        if self.stop_is_defined:
            self.C = stop[0]
            self.A = 0 if (stop[1] - start[1]) == 0 else (stop[0] - start[0]) / (stop[1] - start[1])
            if self.slope_is_defined:
                self.A = slope
        elif slope == 0:
            self.C, self.A = np.nan, slope
        else:
            self.C, self.A = np.sign(slope) * np.inf, slope
        self.B = start[0] - self.A * start[1]

    def eval(self, x):
        if self.undefined:
            return np.nan
        val = self.A * x + self.B
        if self.A < 0:
            val = max(val, self.C)
        elif self.A > 0:
            val = min(val, self.C)
        return val

    __call__ = eval


# --------------------------------------------------------------------------
# grids -- projects/fxs/projectLibrary/ft_grid_pairs.py:274-291,380-392,515-554
# --------------------------------------------------------------------------

def radial_grids(ft_type, q_max, n_r, rc):
    r_max = rc * n_r / q_max                       # mathLibrary.py:1169-1176
    if ft_type == 'midpoint':
        dr, dq = r_max / n_r, q_max / n_r
        rs = np.linspace(dr / 2, r_max - dr / 2, num=n_r, endpoint=True)
        qs = np.linspace(dq / 2, q_max - dq / 2, num=n_r, endpoint=True)
    elif ft_type in ('trapz', 'Zernike'):
        rs = np.linspace(0, r_max, n_r, endpoint=True)
        qs = np.linspace(0, q_max, n_r, endpoint=True)
    elif ft_type == 'gauss':                       # radial_grid_gauss, ft_grid_pairs.py:293-300
        xs = roots_legendre(n_r)[0]
        rs = r_max / 2 * xs + r_max / 2
        qs = q_max / 2 * xs + q_max / 2
    else:
        raise AssertionError(f'ft type {ft_type} not restated')
    return rs, qs


def spherical_grid(radial, thetas, phis):
    """GridFactory.construct_grid('uniform',[r,theta,phi]) -> [N_r,n_theta,n_phi,3] (gridLibrary.py:940-949)."""
    g = np.stack(np.meshgrid(radial, thetas, phis, indexing='ij'), axis=-1)
    return g


class SphericalIntegrator:
    """library/mathLibrary.py:1212-1240 (note pi/n_theta phi weight, roots_legendre order)."""

    def __init__(self, grid):
        self.n_r, self.n_theta, self.n_phi = grid.shape[:-1]
        self.rs = grid[:, 0, 0, 0]
        self.max_r = np.max(self.rs)
        self.norm = 4 / 3 * np.pi * self.max_r ** 3
        self.w = roots_legendre(self.n_theta)[1]

    def integrate(self, values):
        w, rs, n = self.w, self.rs, self.n_theta
        w_shape = (1,) + w.shape + (1,) * (values.ndim - 3)
        rs_shape = rs.shape + (1,) * (values.ndim - 3)
        s2 = np.pi / n * np.sum(w.reshape(w_shape) * np.sum(values, axis=2), axis=1)
        return np.trapezoid(s2 * (rs ** 2).reshape(rs_shape), x=rs, axis=0)

    def integrate_normed(self, values):
        return self.integrate(values) / self.norm


# --------------------------------------------------------------------------
# Hankel -- projects/fxs/projectLibrary/hankel_transforms.py
# --------------------------------------------------------------------------

def hankel_weights(l_max, n_r, rc, mode='midpoint'):
    """w[l,p,k]; p summed. midpoint: :399-410, trapz: :322-333, gauss: :477-490."""
    ls = np.arange(l_max + 1)
    if mode == 'gauss':                            # Gauss-Legendre nodes on [0, 2] with their weights
        xi, wg = roots_legendre(n_r)
        ps = ks = xi + 1
        j = spherical_jn(ls[:, None, None], (ks[None, :] * ps[:, None] * rc * n_r / 4)[None, :, :])
        return ps[None, :, None] ** 2 * j * wg[None, :, None]
    if mode == 'midpoint':
        ps = np.arange(n_r) + 0.5
        ks = np.arange(n_r) + 0.5
    elif mode == 'trapz':
        ps = np.arange(1, n_r)
        ks = np.arange(n_r)
    else:
        raise AssertionError(f'hankel mode {mode} not restated')
    arg = ks[None, :] * ps[:, None] * rc / n_r
    j = spherical_jn(ls[:, None, None], arg[None, :, :])
    return ps[None, :, None] ** 2 * j


def assemble_weights(weights, r_max, rc, mode='midpoint'):
    """-> forward/inverse complex [p,k,l]; trapz :349-375 / midpoint :426-452 (identical prefactors), gauss :505-535
    (step r_max/2 of the [-1,1] -> [0,r_max] map instead of r_max/N)."""
    n_r = weights.shape[-1]
    q_max = rc * n_r / r_max
    orders = np.arange(weights.shape[0])
    div = 2 if mode == 'gauss' else n_r
    fwd = (-1.j) ** (orders[None, None, :]) * (r_max / div) ** 3 * np.sqrt(2 / np.pi)
    inv = (1.j) ** (orders[None, None, :]) * (q_max / div) ** 3 * np.sqrt(2 / np.pi)
    w = np.moveaxis(weights, 0, 2)
    return {'forward': w * fwd, 'inverse': w * inv}


def generate_spherical_ht(w, l_max, mode='midpoint'):
    """CPU flavour on the 'ml' ordered list; :642-658."""
    fw, iw = w['forward'], w['inverse']
    m_orders = np.concatenate((np.arange(l_max + 1, dtype=int), -np.arange(l_max, 0, -1, dtype=int)))
    skip = 1 if mode in ('trapz', 'Zernike') else 0

    def zht(c):
        return tuple(np.sum(fw[:, :, np.abs(m):] * c[m][skip:, None, :l_max - np.abs(m) + 1], axis=0) for m in m_orders)

    def izht(c):
        return tuple(np.sum(iw[:, :, np.abs(m):] * c[m][skip:, None, :l_max - np.abs(m) + 1], axis=0) for m in m_orders)
    return zht, izht


def generate_spherical_ht_direct(w, l_max, mode='midpoint'):
    """Semantics of the OpenCL flavour on the 'direct' [N_r,(L+1)^2] layout; :660-766."""
    fw, iw = w['forward'], w['inverse']
    skip = 1 if mode in ('trapz', 'Zernike') else 0
    l_of = np.floor(np.sqrt(np.arange((l_max + 1) ** 2))).astype(int)

    def apply(W, rho):
        out = np.empty((W.shape[1], rho.shape[1]), dtype=complex)
        for l in range(l_max + 1):
            sel = l_of == l
            out[:, sel] = W[:, :, l].T @ rho[skip:, sel]
        return out
    return (lambda c: apply(fw, c)), (lambda c: apply(iw, c))


def generate_ft(sh, weights, r_max, rc, l_max, mode='midpoint', flavour='ml'):
    """projects/fxs/projectLibrary/fourier_transforms.py:39-86."""
    w = assemble_weights(weights, r_max, rc, mode)
    if flavour == 'ml':
        hankel, ihankel = generate_spherical_ht(w, l_max, mode)
        ht, iht = sh.forward_m, sh.inverse_m
    else:
        hankel, ihankel = generate_spherical_ht_direct(w, l_max, mode)
        ht, iht = sh.forward_d, sh.inverse_d

    def ft(data):
        return iht(hankel(ht(data)))

    def ift(data):
        return iht(ihankel(ht(data)))
    return ft, ift


# --------------------------------------------------------------------------
# coordinate conversions -- library/mathLibrary.py:628-698
# --------------------------------------------------------------------------

def spherical_to_cartesian(grid):
    out = np.array(grid, dtype=float)
    if grid.shape[-1] == 2:
        r, phi = grid[..., 0], grid[..., 1]
        out[..., 0], out[..., 1] = r * np.cos(phi), r * np.sin(phi)
    else:
        r, theta, phi = grid[..., 0], grid[..., 1], grid[..., 2]
        xy = r * np.sin(theta)
        out[..., 0], out[..., 1], out[..., 2] = np.cos(phi) * xy, np.sin(phi) * xy, r * np.cos(theta)
    return out


def cartesian_to_spherical(grid):
    out = np.array(grid, dtype=float)
    if grid.shape[-1] == 2:
        x, y = grid[..., 0], grid[..., 1]
        phi = np.arctan2(y, x)
        out[..., 0], out[..., 1] = np.sqrt(x * x + y * y), np.where(phi < 0, phi + 2 * np.pi, phi)
    else:
        x, y, z = grid[..., 0], grid[..., 1], grid[..., 2]
        r = np.sqrt(x * x + y * y + z * z)
        theta = np.zeros(np.shape(r))
        nz = r != 0
        theta[nz] = np.arccos(np.asarray(z)[nz] / r[nz])
        phi = np.arctan2(y, x)
        out[..., 0], out[..., 1], out[..., 2] = r, theta, np.where(phi < 0, phi + 2 * np.pi, phi)
    return out


# --------------------------------------------------------------------------
# elementwise helpers -- projects/fxs/projectLibrary/misk.py
# --------------------------------------------------------------------------

def square_grid(data):                       # :159-168
    return data * data.conj()


def abs_value(data):                         # :221-225  (complex output buffer in the reference)
    return np.sqrt((data * data.conj()).real).astype(complex)


def add_above_zero_index(a, b):              # :325-329
    result = a + b
    result[0] = a[0]
    return result


def harmonic_coeff_to_deg2_invariants_3d(Ilm):   # fxs_invariant_tools.py:915-923
    return np.array(tuple(Il @ Il.T.conj() for Il in Ilm))


def deg2_invariant_l2_diff(reference_invariant, radial_mask, n_particles, Ilm):
    """_generate_deg2_invariant_diff_3d (fxs_IO_methods.py:412-447) for used_order_ids = arange(n): per-order error array."""
    ref = np.array(reference_invariant)
    rm = np.broadcast_to(np.asarray(radial_mask, dtype=bool), ref.shape[:2])
    mask = ~(rm[:, :, None] & rm[:, None, :])                       # invariant_mask = outer(radial_mask) (reconstruct.py:471)
    ref[mask] = 0
    norm = np.sum(ref * ref.conj(), axis=(1, 2))
    nz = norm != 0
    errors = np.full(len(norm), -1, dtype=float)
    Bl = harmonic_coeff_to_deg2_invariants_3d(Ilm)[:len(norm)].copy()
    Bl[mask] = 0
    ref = ref.copy()
    ref[0] = ref[0] / n_particles
    diff = ref - Bl
    errors[nz] = (np.sum((diff * diff.conj()).real, axis=(1, 2))[nz] / norm[nz]).real
    return errors


def gaussian_fourier_transformed_spherical(points, sigma):   # mathLibrary.py:616-624 (q**4, sic)
    a = 1 / (2 * sigma ** 2)
    return np.sqrt(np.pi / a) * np.exp(-np.pi ** 2 * np.square(points[..., 0]) ** 2 / a)


def get_test_function(support, slope):       # mathLibrary.py:1456-1466
    center = np.mean(support)
    size = support[1] - center

    def f(data):
        nz = (data > support[0]) & (data < support[1])
        v = np.zeros_like(data)
        v[nz] = np.exp(-slope * size ** 2 / (size ** 2 - (data[nz] - center) ** 2))
        return v
    return f


def midpoint_rule(samples, pts, **kw):       # mathLibrary.py:1492-1497
    return (pts[1] - pts[0]) * np.sum(samples, **kw)


# --------------------------------------------------------------------------
# reciprocal projection -- projects/fxs/projectLibrary/fxs_Projections.py:443-929
# --------------------------------------------------------------------------

class ReciprocalProjection:
    def __init__(self, qs, grid_shape, data, l_max, ropt):
        """qs: reconstruction radial q grid; data: invariants record (Appendix B of SURVEY.md)."""
        self.radial_points = qs
        self.grid_shape = grid_shape
        dq = np.asarray(data['data_radial_points'], dtype=float)
        avg = np.asarray(data['average_intensity'], dtype=float)
        self.data_min_q, self.data_max_q = dq.min(), dq.max()
        self.integrated_intensity = midpoint_rule(avg * dq ** 2, dq, axis=0) * 2 * np.sqrt(np.pi)   # :476
        self.positive_orders = np.arange(l_max + 1)
        used_ids = np.asarray(ropt['used_order_ids'])
        used_ids = used_ids[(used_ids <= data['max_order']) & (used_ids <= l_max)]                  # :548-559
        self.used_orders = {o: i for o, i in zip(self.positive_orders, used_ids)}                   # :491-492
        order_ids = list(self.used_orders.values())
        assert order_ids == list(range(len(order_ids))), 'only used_order_ids = arange(n) is consistent in the reference'
        self.number_of_particles = [ropt['number_of_particles']['initial']]
        pms = data['data_projection_matrices']
        same = dq.shape == qs.shape and (dq == qs).all()                                            # :642-651
        if not same:
            from scipy.interpolate import griddata
            kind = ropt.get('regrid', {}).get('interpolation', 'cubic')
            avg = griddata(dq, avg, qs, method=kind, fill_value=0.0)
            pms = [griddata(dq, np.asarray(p), qs, method=kind, fill_value=0.0) for p in pms]
        proj = [np.array(pms[i], dtype=complex) for i in order_ids]
        nq = len(qs)
        self.full_projection_matrices = [np.zeros((nq, min(nq, 2 * o + 1)), dtype=complex) for o in range(l_max + 1)]
        for oid, pm in zip(order_ids, proj):
            self.full_projection_matrices[oid] = pm
        # modify_projection_matrices :679-714
        proj = [p.copy() for p in proj]
        if ropt.get('odd_orders_to_0', False):
            for o in self.used_orders:
                if o % 2 == 1:
                    proj[self.used_orders[o]][:] = 0
        if ropt.get('use_averaged_intensity', False):
            proj[self.used_orders[0]] = avg.astype(complex)[:, None].real * 2 * np.sqrt(np.pi) + 0.j
        for pm in proj:
            pm[:] *= 2
        self.projection_matrices = proj
        # radial mask :578-629 (q_mask types 'none' and manual/region)
        mask = np.full((l_max + 1, nq), False)
        data_mask = mask | ((qs >= self.data_min_q) & (qs <= self.data_max_q))
        mopt = ropt.get('q_mask', {'type': 'none'})
        if mopt['type'] == 'none':
            mask = True
        elif mopt['type'] == 'manual' and mopt['manual']['type'] == 'region':
            region = mopt['manual']['region']
            if (region[0] == False) and (region[1] != False):     # noqa: E712
                mask[:] = (qs < region[1])[None, :]
            elif (region[0] != False) and (region[1] == False):   # noqa: E712
                mask[:] = (qs >= region[0])[None, :]
            elif (region[0] != False) and (region[1] != False):   # noqa: E712
                mask[:] = ((qs >= region[0]) & (qs < region[1]))[None, :]
            else:
                mask[:] = True
        elif mopt['type'] == 'from_projection_matrices':          # :593-598
            for part, lim in zip(mask, np.asarray(data['data_projection_matrices_q_id_limits']['I1I1'])):
                part[:] = (qs > dq[int(lim[0])]) & (qs < dq[int(lim[1]) - 1])
        elif mopt['type'] == 'manual' and mopt['manual']['type'] == 'order_dependent_line':     # :618-623, mathLibrary.py:1131-1137
            p1, p2 = (np.asarray(v, dtype=float) for v in mopt['manual']['order_dependent_line'])
            d = p2 - p1
            rot = np.array([[0, 1], [-1, 0]]) @ d
            grid = np.stack(np.meshgrid(np.arange(l_max + 1, dtype=float), qs, indexing='ij'), axis=-1) - p1
            mask = (-1 * np.sum(grid * rot[None, None, :], axis=-1)) >= 0
        else:
            raise AssertionError('q_mask type not restated')
        self.radial_mask = mask & data_mask
        D = np.diag(qs)
        self.PDs = tuple(self.projection_matrices[i].T.conj() @ D ** 2 for i in order_ids)           # :753-754
        self.unknowns = tuple(np.zeros((len(PD), 2 * o + 1), dtype=complex) for PD, o in zip(self.PDs, self.used_orders))
        self.deg2_invariants = harmonic_coeff_to_deg2_invariants_3d(self.projection_matrices)        # :631-637

    def approximate_unknowns(self, I):           # :761-767
        for unknown, PD, oid in zip(self.unknowns, self.PDs, self.used_orders.values()):
            u, _, vh = np.linalg.svd(PD @ I[oid], full_matrices=False)
            np.matmul(u, vh, out=unknown)
        return self.unknowns

    def mtip_projection(self, I, unknowns):      # :832-872
        rm, pm = self.radial_mask, self.projection_matrices
        out = [np.array(c) for c in I]
        for o_id in self.used_orders.values():
            tmp = pm[o_id] @ unknowns[o_id]
            out[o_id][rm[o_id], ...] = tmp[rm[o_id], ...]
        zero_id = self.used_orders.get(0, False)
        if not isinstance(zero_id, bool):
            out[zero_id][rm[zero_id], ...] = pm[zero_id][rm[zero_id], ...]
            out[zero_id][:] /= np.sqrt(self.number_of_particles[0])
        return out

    def project_to_modified_intensity(self, reciprocal_density, square, new_intensity):   # :899-909
        nz = (square.real >= 0) & (new_intensity.real >= 0)
        temp = np.zeros(reciprocal_density.shape)
        with np.errstate(divide='ignore', invalid='ignore'):
            temp[nz] = new_intensity.real[nz] / square[nz].real
            mult = np.sqrt(temp).astype(complex)
            mult[~nz] = 0
            return reciprocal_density * mult


    def project_to_fixed_intensity(self, reciprocal_density, square, fixed_intensity):   # :889-925, use_fixed_intensity=True
        nz = (square.real >= 0) & (fixed_intensity >= 0)
        temp = np.zeros(reciprocal_density.shape)
        with np.errstate(divide='ignore', invalid='ignore'):
            temp[nz] = fixed_intensity[nz] / square[nz].real
            mult = np.sqrt(temp).astype(complex)
            mult[~nz] = 0
            return reciprocal_density * mult


# --------------------------------------------------------------------------
# real projection / HIO / ER / errors / shrink wrap
# --------------------------------------------------------------------------

class RealProjection:
    """fxs_Projections.py:26-155 + pythonLibrary.py:1289-1318."""

    def __init__(self, popt, real_grid):
        self.opt = popt
        self.real_grid = real_grid
        self.enforce_initial_support = True
        sup = popt['support']['initial_support']
        assert sup['type'] == 'max_radius'
        self._initial_mask = ~np.where(real_grid[..., 0] < sup['max_radius'], True, False)     # :137-140
        self._mask = [self._initial_mask.copy()]

    @property
    def initial_support(self):
        return ~self._initial_mask.copy()

    @property
    def support(self):
        return ~self._mask[0]

    @support.setter
    def support(self, support):                 # :53-58
        self._mask[0] = (self._initial_mask | (~support)) if self.enforce_initial_support else ~support

    def projection(self, data):                 # :110-130 ; in place, like the reference
        mask = False
        masks = {}
        for key in self.opt['apply']:
            if key == 'support':
                m = self._mask[0]
                data[m] = 0
                p = m
            elif key == 'value_threshold':
                thr = self.opt['value_threshold'].get('threshold', 0.0)
                lo = isinstance(thr[0], (float, int)) and not isinstance(thr[0], bool)
                hi = isinstance(thr[1], (float, int)) and not isinstance(thr[1], bool)
                re = data.real
                p = False
                if lo:
                    ml = re < thr[0]
                if hi:
                    mh = re > thr[1]
                if lo:
                    re[ml] = thr[0]
                    p = ml
                if hi:
                    re[mh] = thr[1]
                    p = (mh | p) if lo else mh
            elif key == 'limit_imag':
                t = self.opt['limit_imag'].get('threshold', 0.0)
                im = data.imag
                p = np.abs(im) >= t
                im[p] = 0
            elif key == 'average_center':          # :96-110: the first shells are replaced by their angular mean; mask False
                thresh = int(self.opt['average_center'].get('max_radial_id', 1))
                axes = tuple(range(1, data.ndim))
                data[:thresh] = np.mean(data[:thresh], axis=axes).reshape((-1,) + (1,) * (data.ndim - 1))
                p = False
            else:                                  # :113-118: 'projection {} not known. Ignoring it.'
                continue
            masks[key] = p
            mask = mask | p
        masks['all'] = mask
        return [data, masks]


def hybrid_input_output(beta, without_projection, projection_out, _input, considered=('all',)):   # fxs_IO_methods.py:56-63
    out, md = projection_out
    m = md[considered[0]]
    for name in considered[1:]:
        m = m | md[name]
    return np.where(m, _input - beta * (without_projection - out), out)


def error_reduction(out_without_projection, out, _input):     # fxs_IO_methods.py:67-68
    return np.array(out[0])


def l2_projection_diff(integrator, values, projected_values, mask=True):   # fxs_IO_methods.py:97-128
    diff = values - projected_values[0]
    sd = (diff * diff.conj()).real
    sq = (values * values.conj()).real
    if mask is not True:
        sd[~mask] = 0
        sq[~mask] = 0
    d, v = integrator.integrate(sd), integrator.integrate(sq)
    return d / v if v != 0 else np.inf


class ShrinkWrap:
    """fxs_Projections.py:178-298 (threshold mode)."""

    def __init__(self, reciprocal_grid):
        self.reciprocal_grid = reciprocal_grid
        self.default_sigma = np.pi / reciprocal_grid[:, 0, 0, 0].max()
        self._sigma = self.default_sigma
        self._threshold = 0.06
        self.gaussian_values = gaussian_fourier_transformed_spherical(reciprocal_grid, self._sigma)

    def set_threshold(self, v):                 # :218-227
        self._threshold = 0 if v < 0 else (1 if v >= 1 else v)

    def set_sigma(self, value):                 # :233-243
        ok_type = not ((not _is_number(value)) or isinstance(value, bool))
        ok = bool(ok_type and value > 0)
        self._sigma = value if ok else self.default_sigma
        self.gaussian_values = gaussian_fourier_transformed_spherical(self.reciprocal_grid, self._sigma)

    def get_new_mask(self, conv):               # :245-258
        c = conv.real.copy()
        c[c < 0] = 0
        mx, mn = c.max(), c.min()
        return c >= mn + self._threshold * (mx - mn)


# --------------------------------------------------------------------------
# the loop -- projects/fxs/reconstruct.py:515-619,768-1036,1115-1258
# --------------------------------------------------------------------------

class MTIP:
    def __init__(self, opt, data, sht_factory=OracleSH, ft_flavour='ml'):
        self.opt = opt
        g = opt['grid']
        l_max = int(g['max_order'])
        self.l_max = l_max
        self.sh = sht_factory(l_max, mode_flag='complex', n_phi=g.get('n_phi', 0), n_theta=g.get('n_theta', 0))
        fto = opt['fourier_transform']
        rc = fto.get('reciprocity_coefficient', np.pi)
        max_q = g['max_q']
        if not isinstance(max_q, float):           # reconstruct.py:258-261
            max_q = float(np.max(data['data_radial_points']))
        n_r = int(g['n_radial_points'])
        self.rs, self.qs = radial_grids(fto['type'], max_q, n_r, rc)
        self.real_grid = spherical_grid(self.rs, self.sh.theta, self.sh.phi)
        self.reciprocal_grid = spherical_grid(self.qs, self.sh.theta, self.sh.phi)
        self.weights = hankel_weights(l_max, n_r, rc, fto['type'])
        r_max = np.max(self.rs)                    # reconstruct.py:329 (max grid point, not domain radius)
        self.ft, self.ift = generate_ft(self.sh, self.weights, r_max, rc, l_max, fto['type'], ft_flavour)
        self.rp = ReciprocalProjection(self.qs, self.reciprocal_grid.shape[:-1], data, l_max, opt['projections']['reciprocal'])
        popt = dict(opt['projections']['real']['projections'])
        self.real_pr = RealProjection(popt, self.real_grid)
        self.sw = ShrinkWrap(self.reciprocal_grid)
        self.integrator = SphericalIntegrator(self.real_grid)
        self.results = {}
        err = opt['main_loop']['error']['methods']
        self.err_inside = err['real'].get('l2_projection_diff', {}).get('inside_initial_support', False)
        self.beta = None

    # -- sketches (reconstruct.py:518-528,576-605) --
    def mtip_start(self, rho_hat):
        rp = self.rp
        sq = square_grid(rho_hat)
        I = self.sh.forward_l(sq)
        unk = rp.approximate_unknowns(I)
        self.results['fxs_unknowns'] = unk
        Ip = rp.mtip_projection(I, unk)
        I_proj = self.sh.inverse_l(Ip)
        return rp.project_to_modified_intensity(rho_hat, np.array(sq), I_proj)

    def io_step(self, method, rho, ft_stab, fixed_intensity=None):
        rho_hat = self.ft(rho)
        if fixed_intensity is None:
            rho_hat_new = self.mtip_start(rho_hat)
        else:                                    # MTIP_start_non_FXS sketch (reconstruct.py:529-534)
            rho_hat_new = self.rp.project_to_fixed_intensity(rho_hat, np.array(square_grid(rho_hat)), fixed_intensity)
        if ft_stab:
            rho_rt = self.ift(rho_hat)
            rho_new = add_above_zero_index(self.ift(rho_hat_new), rho - rho_rt)
        else:
            rho_new = self.ift(rho_hat_new)
        rho_new_copy = np.array(rho_new)
        proj = self.real_pr.projection(rho_new)
        if method.startswith('HIO'):
            considered = self.opt['projections']['real']['HIO'].get('considered_projections', ['all'])
            rho_next = hybrid_input_output(self.beta, rho_new_copy, proj, rho, considered)
        else:
            rho_next = error_reduction(rho_new_copy, proj, rho)
        mask = self.real_pr.initial_support if self.err_inside else True
        e = l2_projection_diff(self.integrator, rho_new_copy, proj, mask)
        self.results['errors']['real']['l2_projection_diff'].append(e)
        return rho_hat_new, rho_next

    def shrink_wrap(self, rho):
        c = self.ift(self.ft(abs_value(rho)) * self.sw.gaussian_values)
        return self.sw.get_new_mask(c)

    def density_guess(self, rng):                # reconstruct.py:1115-1174 with an injected generator
        dopt = self.opt['density_guess']
        radius = dopt.get('radius', self.opt['particle_radius'])
        if isinstance(radius, bool):
            radius = self.opt['particle_radius']
        if radius < 0:
            radius = np.max(self.real_grid[..., 0])
        A = 1 + 1 / dopt['random']['SNR'] * rng.random(self.real_grid.shape[:-1])
        assert dopt['type'] == 'bump'
        density = A * get_test_function([-radius, radius], dopt['bump']['slope'])(self.real_grid[..., 0])
        tot = self.integrator.integrate((density * density.conj()).real)
        density = density * np.sqrt(self.rp.integrated_intensity / tot)
        return density.astype(complex)

    def _sw_ramps(self):                         # reconstruct.py:1212-1258
        sw_opt = self.opt['projections']['real']['shrink_wrap']
        names = self.opt['main_loop']['sub_loops']['order']
        sig, thr = [], []
        for lid in range(len(names)):
            s = sw_opt['sigmas'][lid] if len(sw_opt['sigmas']) - 1 >= lid else False
            if not isinstance(s, (list, tuple)):
                s = [s]
            sig.append(LinearRamp(*s, default_start=self.sw.default_sigma, default_stop=self.sw.default_sigma))
            t = sw_opt['thresholds'][lid] if len(sw_opt['thresholds']) - 1 >= lid else 0.1
            if not isinstance(t, (list, tuple)):
                t = [t]
            thr.append(LinearRamp(*t))
        return sig, thr

    def run(self, rho0=None, rng=None):
        opt = self.opt
        loops = opt['main_loop']['sub_loops']
        hio_opt = opt['projections']['real']['HIO']
        sup_opt = opt['projections']['real']['projections']['support']['enforce_initial_support']
        sig_ramps, thr_ramps = self._sw_ramps()
        self.results['errors'] = {'real': {'l2_projection_diff': []}, 'reciprocal': {}, 'main': []}
        errors = self.results['errors']
        if rho0 is None:
            rho0 = self.density_guess(rng if rng is not None else np.random.default_rng())
        rho_hat = self.ft(rho0)                  # reconstruct.py:957-962
        rho = self.ift(rho_hat)
        pair = (rho_hat, rho)
        initial = tuple(d.copy() for d in pair)
        state = dict(pair=pair, mask=self.real_pr.initial_support, best_pair=pair, best_error=np.inf,
                     best_iteration=0, best_mask=self.real_pr.initial_support)
        enforce_list = []
        iterations = []

        def update_sw(it, lid):
            if not sig_ramps[lid].undefined:
                self.sw.set_sigma(sig_ramps[lid](it))
            if not thr_ramps[lid].undefined:
                self.sw.set_threshold(thr_ramps[lid](it))

        for lid, name in enumerate(loops['order']):
            lopt = loops[name]
            beta = hio_opt['beta'][lid] if len(hio_opt['beta']) - 1 >= lid else [0.5, 0.5, -1 / 700, 1600]
            beta_ramp = ExponentialRamp(*beta)
            limit = [sup_opt['if_error_bigger_than']] if sup_opt['apply'] else [np.inf]
            if 'SW' in lopt['order']:
                update_sw(0, lid)
            step, sw_step = 0, 0
            it = 0
            # `hist` of the reference's loop (reconstruct.py:853,911): bound at the start of the sub-loop and re-bound at the start
            # of every HIO/ER iteration, i.e. after an iteration it still names the pair that iteration STARTED from
            hist_last = state['pair']
            latest_intensity = None
            for it in range(1, lopt['iterations'] + 1):
                for key in lopt['order']:
                    mo = lopt['methods'][key]
                    repeats = mo['iterations'] if isinstance(mo, dict) else mo
                    if key == 'SW':
                        support = self.shrink_wrap(state['pair'][1])
                        enforce = errors['main'][-1:] > limit      # reconstruct.py:879 (list comparison, sic)
                        enforce_list.append(enforce)
                        self.real_pr.enforce_initial_support = enforce
                        self.real_pr.support = support
                        state['mask'] = self.real_pr.support
                        sw_step += 1
                        update_sw(sw_step, lid)
                    elif key == 'SW_center':             # reconstruct.py:886-897 with the sketch :606-613
                        # The sketch ends in ['calculate_support_mask','id','id'] on (conv, x, FT(x), x); get_new_mask takes ONE
                        # argument (fxs_Projections.py:247), so the process returns (support, x, FT(x)) and the loop binds
                        # `support, ft_density, density` to it: the history pair becomes (reciprocal = x, real = FT(x)) -- the
                        # real and reciprocal densities are exchanged (a quirk of the reference, reproduced here and on the device).
                        enforce = np.float64(errors['main'][-1]) > limit    # scalar against the one-element list (:887)
                        self.real_pr.enforce_initial_support = enforce
                        enforce_list.append(enforce)
                        for _ in range(repeats):
                            rho_c = state['pair'][1]
                            support = self.shrink_wrap(rho_c)
                            self.real_pr.support = support
                            state['mask'] = self.real_pr.support
                            state['pair'] = (np.array(rho_c), self.ft(rho_c))
                            sw_step += 1
                            update_sw(sw_step, lid)
                    else:
                        if key in ('ER_non_FXS', 'HIO_non_FXS'):     # :899-904
                            if latest_intensity is None:
                                latest_intensity = np.abs(hist_last[0]).real
                        else:
                            latest_intensity = None
                        ft_stab = bool(mo.get('ft_stab', False)) if isinstance(mo, dict) else False
                        for _ in range(repeats):
                            self.beta = beta_ramp.eval(step)
                            hist_last = state['pair']
                            new_pair = self.io_step(key, state['pair'][1], ft_stab, latest_intensity)
                            new_pair = tuple(np.array(a) for a in new_pair)
                            state['pair'] = new_pair
                            main = float(np.mean([errors['real']['l2_projection_diff'][-1]]))
                            errors['main'].append(main)
                            if state['best_error'] > main:
                                state.update(best_error=main, best_pair=new_pair, best_iteration=it, best_mask=state['mask'])
                            step += 1
            if state['best_iteration'] > lopt.get('best_density_not_in_first_n_iterations', np.inf):
                state['pair'] = state['best_pair']
                self.real_pr.support = state['best_mask']
                state['mask'] = state['best_mask']
            iterations.append(it)
        best_out, last_out = state['best_pair'], state['pair']
        if opt.get('output_density_modifiers', {}).get('shift_to_center', False):      # reconstruct.py:988-993
            best_out, last_out = self._shift_to_center(*best_out[:2]), self._shift_to_center(*last_out[:2])
        last_inv = self._last_invariants(last_out[1])
        return {
            'real_density': best_out[1], 'reciprocal_density': best_out[0],
            'last_real_density': last_out[1], 'last_reciprocal_density': last_out[0],
            'final_error': state['best_error'], 'initial_density': initial[1],
            'initial_support': self.real_pr.initial_support, 'support_mask': state['best_mask'],
            'last_support_mask': state['mask'], 'loop_iterations': np.sum(iterations) + 1,
            'error_dict': {'main': np.array(errors['main']),
                           'real': {k: np.array(v) for k, v in errors['real'].items()}, 'reciprocal': {}},
            'fxs_unknowns': self.results.get('fxs_unknowns'),
            'last_deg2_invariant': last_inv,
        }

    # -- output modifiers (reconstruct.py:721-755): only shift_to_center is restated ----------------------------
    def _shift_to_center(self, rho_hat, rho):
        """shift_center sketch (:732-738): (rho_hat * e^{+i q.c}, IFT(FT(rho) * e^{+i q.c})), c = centre of mass of Re rho
        (misk.py:295-312), phases of generate_shift_by_operator(..., opposite_direction=True) (fxs_Projections.py:1419-1444)."""
        cart_r, cart_q = spherical_to_cartesian(self.real_grid), spherical_to_cartesian(self.reciprocal_grid)
        total = self.integrator.integrate(rho.real)
        if total == 0:
            total = 1
        center = self.integrator.integrate(cart_r * rho[..., None].real) / total
        center = spherical_to_cartesian(cartesian_to_spherical(center))
        phases = np.exp(1.j * (cart_q * center).sum(axis=-1))
        self.results['neg_center_pos'] = center
        return rho_hat * phases, self.ift(self.ft(rho) * phases)

    def _last_invariants(self, rho):             # reconstruct.py:757-765
        return harmonic_coeff_to_deg2_invariants_3d(self.sh.forward_l(square_grid(self.ft(rho))))


# --------------------------------------------------------------------------
# synthetic invariants (bench / test inputs) -- simulate_ccd.py:194-230,
# settings/simulate_ccd/tutorial.yaml:11-20, fxs_invariant_tools.py:1114-1207
# --------------------------------------------------------------------------

def six_sphere_density(real_grid, centers=None, radius=70.0, densities=(25, 50, 25, 50, 25, 50)):
    if centers is None:
        centers = [(0.0, 0.0, 0.0)] + [(140.0, np.pi / 2, k * 2 * np.pi / 5) for k in range(5)]
    r, t, p = real_grid[..., 0], real_grid[..., 1], real_grid[..., 2]
    xyz = np.stack([r * np.sin(t) * np.cos(p), r * np.sin(t) * np.sin(p), r * np.cos(t)], axis=-1)
    rho = np.zeros(r.shape)
    for (cr, ct, cp), d in zip(centers, densities):
        c = np.array([cr * np.sin(ct) * np.cos(cp), cr * np.sin(ct) * np.sin(cp), cr * np.cos(ct)])
        rho += d * (np.linalg.norm(xyz - c, axis=-1) < radius)
    return rho


def invariants_from_density(density, ft, sh, qs):
    """B_l -> V_l as `extract` would hand them to reconstruct, divided by 2 to cancel fxs_Projections.py:711-713."""
    fd = ft(density.astype(complex))
    I = sh.forward_l(fd * fd.conj())
    Bl = harmonic_coeff_to_deg2_invariants_3d(I)
    pms = []
    for l, b in enumerate(Bl):
        b = (b + b.T.conj()) / 2
        ev, evec = np.linalg.eigh(b.real)
        order = np.argsort(ev)[::-1]
        ev, evec = ev[order], evec[:, order]
        n = min(len(evec), 2 * l + 1)
        ev, evec = ev[:n].copy(), evec[:, :n].copy()
        neg = ev < 0
        ev[neg] = 0
        evec[:, neg] = 0
        pms.append(((evec @ np.diag(np.sqrt(ev))) / 2).astype(complex))
    avg = np.sqrt(np.diag(Bl[0]).real / (4 * np.pi))
    return {'dimensions': 3, 'xray_wavelength': 1.23984, 'average_intensity': avg, 'data_radial_points': qs.copy(),
            'data_angular_points': sh.phi.copy(), 'max_order': len(Bl) - 1, 'data_projection_matrices': pms,
            'deg_2_invariant': Bl, 'number_of_particles': 1}
