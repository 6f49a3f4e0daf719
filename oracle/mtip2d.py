"""ORACLE (test infrastructure, never shipped on the product path) -- 2-D flavour of the fxs MTIP path.

numpy restatement, in the reference's own formulation, of what `xframe fxs reconstruct` does with `dimensions: 2`:
circular-harmonic transforms (xframe/library/mathLibrary.py:469-496), polar Hankel weights and transform
(projects/fxs/projectLibrary/hankel_transforms.py:412-452,602-640), the polar grid pair (ft_grid_pairs.py:282-291,325-336),
PolarIntegrator (mathLibrary.py:1242-1265), the 2-D branches of ReciprocalProjection (fxs_Projections.py:471-537,679-706,
723-750,792-830,852-862) and of the loop assembly (reconstruct.py:345-350,1126-1127).  Everything dimension-agnostic
(real projection, HIO/ER, error metric, shrink wrap, loop driver) is inherited from oracle/mtip.py.

Pinned against the UNMODIFIED reference by tests/golden/make_golden_2d.py -> tests/golden/ref2d_*.npz
(tests/test_oracle_golden_2d.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import this.
"""
import numpy as np
from scipy.special import jv, roots_legendre

from . import mtip as O


# ------------------------------------------------------------------ transforms
def cht_complex_forward(x):                      # mathLibrary.py:469-475
    return np.fft.fft(np.copy(x), axis=1) / x.shape[-1]


def cht_complex_inverse(c):                      # :478-482
    return np.fft.ifft(np.copy(c) * c.shape[-1], axis=1)


def cht_real_forward(x):                         # :484-490
    return np.fft.rfft(np.copy(x.real)) / x.shape[-1]


def cht_real_inverse(c, size):                   # :492-496
    return np.fft.irfft(np.copy(c) * size, size)


def polar_hankel_weights(m_max, n_r, rc, mode='midpoint'):      # hankel_transforms.py:412-424 / :335-347
    ms = np.arange(m_max + 1)
    if mode == 'gauss':                          # :492-503
        xi, wg = roots_legendre(n_r)
        ps = ks = xi + 1
        return ps[None, :, None] * jv(ms[:, None, None], (ks[None, :] * ps[:, None] * rc * n_r / 4)[None]) * wg[None, :, None]
    if mode == 'midpoint':
        ps, ks = np.arange(n_r) + 0.5, np.arange(n_r) + 0.5
    elif mode == 'trapz':
        ps, ks = np.arange(1, n_r), np.arange(n_r)
    else:
        raise AssertionError('mode not restated')
    return ps[None, :, None] * jv(ms[:, None, None], (ks[None, :] * ps[:, None] * rc / n_r)[None])


def assemble_weights_2d(weights, r_max, rc, mode='midpoint'):     # hankel_transforms.py:426-452 (dimensions == 2), gauss :505-535
    n_r = weights.shape[-1]
    q_max = rc * n_r / r_max
    orders = np.arange(weights.shape[0])
    all_orders = np.concatenate((orders, -orders[:0:-1]))
    div = 2 if mode == 'gauss' else n_r
    fwd = (-1.j) ** (all_orders[None, None, :]) * (r_max / div) ** 2
    inv = (1.j) ** (all_orders[None, None, :]) * (q_max / div) ** 2
    w = np.concatenate((weights, (-1.0) ** orders[:0:-1, None, None] * weights[:0:-1]), axis=0)
    w = np.moveaxis(w, 0, 2)
    return {'forward': w * fwd, 'inverse': w * inv}


def generate_polar_ht(w, mode='midpoint'):       # hankel_transforms.py:602-640 (all orders used)
    fw, iw = w['forward'], w['inverse']
    skip = 1 if mode in ('trapz', 'Zernike') else 0
    return (lambda c: np.sum(fw * c[skip:, None, :], axis=0)), (lambda c: np.sum(iw * c[skip:, None, :], axis=0))


def generate_ft_2d(weights, r_max, rc, mode='midpoint'):        # fourier_transforms.py:53-85 with the 'm' transforms
    hankel, ihankel = generate_polar_ht(assemble_weights_2d(weights, r_max, rc, mode), mode)
    return (lambda d: cht_complex_inverse(hankel(cht_complex_forward(d)))), \
           (lambda d: cht_complex_inverse(ihankel(cht_complex_forward(d))))


def polar_grid(radial, phis):                    # gridLibrary.py:940-949 via ft_grid_pairs.py:325-336
    r, p = np.meshgrid(radial, phis, indexing='ij')
    return np.stack((r, p), axis=-1)


class PolarIntegrator:                           # mathLibrary.py:1242-1265
    def __init__(self, grid):
        self.rs, self.phis = grid[:, 0, 0], grid[0, :, 1]

    def integrate(self, values):
        trapz = getattr(np, 'trapezoid', None) or np.trapz
        s = trapz(values, x=self.phis, axis=1)
        rs = self.rs.reshape(self.rs.shape + (1,) * (values.ndim - 2))
        return trapz(s * rs, x=self.rs, axis=0)


# ------------------------------------------------------------------ reciprocal projection, 2-D branches
class ReciprocalProjection2D:
    def __init__(self, qs, grid_shape, data, m_max, ropt):
        self.radial_points, self.grid_shape = qs, grid_shape
        dq = np.asarray(data['data_radial_points'], dtype=float)
        avg = np.asarray(data['average_intensity'], dtype=float)
        self.data_min_q, self.data_max_q = dq.min(), dq.max()
        self.integrated_intensity = O.midpoint_rule(avg * dq, dq, axis=0) * 2 * np.sqrt(np.pi)        # :473-474
        self.positive_orders = np.arange(m_max + 1)
        used_ids = np.asarray(ropt['used_order_ids'])
        used_ids = used_ids[(used_ids <= data['max_order']) & (used_ids <= m_max)]
        self.used_orders = {o: i for o, i in zip(self.positive_orders, used_ids)}
        order_ids = list(self.used_orders.values())
        assert order_ids == list(range(len(order_ids)))
        self.number_of_particles = [ropt['number_of_particles']['initial']]
        pms = np.asarray(data['data_projection_matrices'])[order_ids]                               # [orders, q]
        same = dq.shape == qs.shape and (dq == qs).all()
        if not same:                                                                                # :657-661
            from scipy.interpolate import griddata
            kind = ropt.get('regrid', {}).get('interpolation', 'cubic')
            avg = griddata(dq, avg, qs, method=kind, fill_value=0.0)
            pms = np.array([griddata(dq, p, qs, method=kind, fill_value=0.0) for p in pms])
        pms = np.array(pms, dtype=complex)
        self.full_projection_matrices = np.zeros((m_max + 1, len(qs)), dtype=complex)
        self.full_projection_matrices[order_ids] = pms
        proj = pms.copy()                                                                           # :679-706
        if ropt.get('odd_orders_to_0', False):
            proj[np.array(tuple(self.used_orders)) % 2 == 1, :] = 0
        if ropt.get('use_averaged_intensity', False):
            proj[self.used_orders[0]] = avg.astype(complex)
        self.projection_matrices = proj
        mask = np.full((m_max + 1, len(qs)), False)
        data_mask = mask | ((qs >= self.data_min_q) & (qs <= self.data_max_q))
        assert ropt.get('q_mask', {'type': 'none'})['type'] == 'none'
        self.radial_mask = np.broadcast_to(True & data_mask, mask.shape).copy()
        so = ropt.get('SO_freedom', {'use': False})
        self.so_order_id = None
        self.use_SO_freedom = bool(so.get('use', False))
        self.radial_high_pass = so.get('radial_high_pass', 0.2)
        if self.use_SO_freedom:
            self.so_order_id = self.rank_projection_matrix_orders_2d()[0][0]                        # :964-971
        self.n_half = (grid_shape[1] + 1) // 2
        self.deg2_invariants = np.array(tuple(Im[:, None] * Im[None, :].conj() for Im in proj))      # fxs_invariant_tools.py:906-914

    def rank_projection_matrix_orders_2d(self):  # :933-962
        hp = int((len(self.radial_points) - 1) * self.radial_high_pass)
        radial_points = self.radial_points[hp:]
        orders = np.array(list(self.used_orders.keys()))
        order_mask = (orders % 2 == 0) & (orders != 0)
        vect = self.projection_matrices[order_mask, hp:].T * radial_points[:, None]
        metric = np.mean(np.abs(vect), axis=0)
        sorted_indices = np.argsort(metric)[::-1]
        so_ids = order_mask.nonzero()[0][sorted_indices]
        return so_ids, orders[so_ids], sorted_indices

    def approximate_unknowns(self, I):           # :727-748 ; I [N_r, M+1]
        ids = list(self.used_orders.values())
        pm = self.projection_matrices.T
        s = np.sum(I[:, ids] * np.conjugate(pm) * self.radial_points[:, None], axis=0)
        unk = np.ones(len(ids), dtype=complex)
        nz = s != 0
        unk[nz] = s[nz] / np.abs(s[nz])
        if self.use_SO_freedom:
            unk[self.so_order_id] = 1
        return unk

    def generate_remaining_SO_projection(self, n_angular_points):    # :1022-1095
        orders = np.array(list(self.used_orders.keys()))
        projection_orders = np.concatenate((np.arange(int(n_angular_points / 2) + 1),
                                            -1 * np.arange(int(n_angular_points / 2) + n_angular_points % 2)[:0:-1]))
        order_mask = (orders % 2 == 0) & (orders != 0)
        harmonic_orders = orders[order_mask]
        max_order = np.max(harmonic_orders)
        so_ids, so_orders, sorted_order_indices = self.rank_projection_matrix_orders_2d()
        remaining_rotations = current_order = so_orders[0]
        free_orders_mask = True
        order_indices, angle_coeffs, angles, gcds = (), (), (), ()
        while remaining_rotations > 2:
            multiples = np.arange(current_order, max_order + 1, current_order)
            multiple_indices = np.where(np.isin(harmonic_orders, multiples))
            free_orders_mask = free_orders_mask * ~np.isin(sorted_order_indices, multiple_indices)
            if not free_orders_mask.any():
                break
            current_order_index = sorted_order_indices[free_orders_mask][0]
            current_order = harmonic_orders[current_order_index]
            gcd = np.gcd(remaining_rotations, current_order)
            n_ind = remaining_rotations / gcd
            smallest_angle = 2 * np.pi / n_ind
            coeff = np.argmin((np.arange(1, n_ind) * current_order / gcd) % n_ind) + 1
            order_indices += (current_order_index,)
            angle_coeffs += (coeff,)
            angles += (smallest_angle,)
            gcds += (gcd,)
            remaining_rotations = gcd

        def apply_SO_freedom(harmonic_coefficients, fxs_unknowns):
            phases = (-1.j * np.log(fxs_unknowns[order_mask])).real
            rotation_phase = 0
            for oi, angle, ac, gcd in zip(order_indices, angles, angle_coeffs, gcds):
                rotation_phase -= (phases[oi] // angle) * ac * angle / gcd
            return harmonic_coefficients * np.exp(1.j * projection_orders * rotation_phase)
        return apply_SO_freedom

    def mtip_projection(self, I, unknowns):      # :804-826 + :852-862
        ids = np.array(list(self.used_orders.values()))
        pm = self.projection_matrices.T
        out = np.array(I, dtype=complex)
        mask = np.zeros(out.shape, dtype=bool)
        for o in ids:
            mask[:, o] = self.radial_mask[o]
        rm2 = self.radial_mask[ids].T
        out[mask] = (pm * unknowns[None, :])[rm2]
        zid = self.used_orders.get(0, False)
        if not isinstance(zid, bool):
            out[self.radial_mask[zid], zid] = pm[self.radial_mask[zid], 0]
        out[:, 0] /= np.sqrt(self.number_of_particles[0])
        return out

    project_to_modified_intensity = O.ReciprocalProjection.project_to_modified_intensity


class ShrinkWrap2D(O.ShrinkWrap):
    def __init__(self, reciprocal_grid):         # fxs_Projections.py:189-190
        self.reciprocal_grid = reciprocal_grid
        self.default_sigma = np.pi / reciprocal_grid[:, 0, 0].max()
        self._sigma = self.default_sigma
        self._threshold = 0.06
        self.gaussian_values = O.gaussian_fourier_transformed_spherical(reciprocal_grid, self._sigma)


class MTIP2D(O.MTIP):
    """The loop of oracle/mtip.py with the 2-D operator set (reconstruct.py:319-371 with dimensions == 2)."""

    def __init__(self, opt, data):
        self.opt = opt
        g = opt['grid']
        m_max = int(g['max_order'])
        self.l_max = m_max
        self.n_phi = 2 * m_max + 1                                                               # harmonic_transforms.py:44-47
        phis = np.arange(self.n_phi) / self.n_phi * 2 * np.pi
        fto = opt['fourier_transform']
        rc = fto.get('reciprocity_coefficient', np.pi)
        max_q = g['max_q']
        if not isinstance(max_q, float):
            max_q = float(np.max(data['data_radial_points']))
        n_r = int(g['n_radial_points'])
        self.rs, self.qs = O.radial_grids(fto['type'], max_q, n_r, rc)
        self.real_grid, self.reciprocal_grid = polar_grid(self.rs, phis), polar_grid(self.qs, phis)
        self.weights = polar_hankel_weights(m_max, n_r, rc, fto['type'])
        self.ft, self.ift = generate_ft_2d(self.weights, np.max(self.rs), rc, fto['type'])          # reconstruct.py:329
        self.rp = ReciprocalProjection2D(self.qs, self.reciprocal_grid.shape[:-1], data, m_max, opt['projections']['reciprocal'])
        self.real_pr = O.RealProjection(dict(opt['projections']['real']['projections']), self.real_grid)
        self.sw = ShrinkWrap2D(self.reciprocal_grid)
        self.integrator = PolarIntegrator(self.real_grid)
        self.results = {}
        err = opt['main_loop']['error']['methods']
        self.err_inside = err['real'].get('l2_projection_diff', {}).get('inside_initial_support', False)
        self.beta = None

    def mtip_start(self, rho_hat):               # reconstruct.py:518-528 with the 'real' harmonic transform (:347-348)
        sq = O.square_grid(rho_hat)
        I = cht_real_forward(sq)
        unk = self.rp.approximate_unknowns(I)
        self.results['fxs_unknowns'] = unk.copy()
        Ip = self.rp.mtip_projection(I, unk)
        I_proj = cht_real_inverse(Ip, self.n_phi)
        return self.rp.project_to_modified_intensity(rho_hat, np.array(sq), I_proj)

    def run(self, rho0=None, rng=None):
        # output modifiers (reconstruct.py:721-755): in 2-D `fix_orientation` (with SO_freedom) = shift_to_center + orientation fix
        mods = self.opt.get('output_density_modifiers', {})
        self._fix = bool(mods.get('fix_orientation', False)) and self.rp.use_SO_freedom
        if self._fix:
            self.opt = dict(self.opt)
            self.opt['output_density_modifiers'] = dict(mods, shift_to_center=True)
            self._apply_so = self.rp.generate_remaining_SO_projection(self.n_phi)
        return super().run(rho0=rho0, rng=rng)

    def _shift_to_center(self, rho_hat, rho):
        a, b = super()._shift_to_center(rho_hat, rho)
        if getattr(self, '_fix', False):         # fix_orientation sketch (:740-745)
            unk = self.results['fxs_unknowns']
            a = cht_complex_inverse(self._apply_so(cht_complex_forward(a), unk))
            b = cht_complex_inverse(self._apply_so(cht_complex_forward(b), unk))
        return a, b

    def _last_invariants(self, rho):             # fxs_invariant_tools.py:906-914
        I = cht_real_forward(O.square_grid(self.ft(rho)))
        return np.array(tuple(Im[:, None] * Im[None, :].conj() for Im in I.T))


# ------------------------------------------------------------------ synthetic 2-D inputs
def disk_model_density(real_grid, centers=None, radius=70.0, densities=(25, 50, 25, 50, 25, 50)):
    """2-D projection of the tutorial model: discs at (0,0) and 5 x (140, 2 pi k / 5) (simulate_ccd/tutorial.yaml:11-20)."""
    if centers is None:
        centers = [(0.0, 0.0)] + [(140.0, k * 2 * np.pi / 5) for k in range(5)]
    r, p = real_grid[..., 0], real_grid[..., 1]
    x, y = r * np.cos(p), r * np.sin(p)
    rho = np.zeros(r.shape)
    for (cr, cp), d in zip(centers, densities):
        rho += d * (np.sqrt((x - cr * np.cos(cp)) ** 2 + (y - cr * np.sin(cp)) ** 2) < radius)
    return rho


def invariants_from_density_2d(density, ft, qs, phis):
    """I_m(q) of |FT rho|^2 -> projection 'matrices' V_m(q) = I_m(q) (2-D: B_m = I_m I_m^*, rank one) and <I>(q) = I_0(q)."""
    I = cht_real_forward(O.square_grid(ft(density.astype(complex))))
    return {'dimensions': 2, 'xray_wavelength': 1.23984, 'average_intensity': I[:, 0].real.copy(), 'data_radial_points': qs.copy(),
            'data_angular_points': phis.copy(), 'max_order': I.shape[1] - 1, 'data_projection_matrices': np.ascontiguousarray(I.T),
            'number_of_particles': 1}
