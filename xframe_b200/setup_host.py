"""One-off host-side setup of a reconstruction (numpy/scipy): everything the reference computes once per run
before the phasing loop starts.  The loop itself never comes back here.

  * projection constants    ReciprocalProjection.__init__ / _regrid_data / modify_projection_matrices /
                            generate_radial_mask            (projectLibrary/fxs_Projections.py:471-537,578-714)
  * initial support         RealProjection.generate_initial_support_mask (fxs_Projections.py:133-155)
  * initial density guess   MTIP.generate_density_guess_method           (reconstruct.py:1115-1174)
  * synthetic invariants    simulate_ccd._bl_from_density + deg2_invariant_to_projection_matrices_3d
                            (simulate_ccd.py:194-230, fxs_invariant_tools.py:1114-1207) -- bench / test inputs
"""
import numpy as np

from ._lib import XfbError


def midpoint_rule(samples, pts):                               # mathLibrary.py:1492-1497
    return (pts[1] - pts[0]) * np.sum(samples, axis=0)


def radial_mask(qs, dq, n_orders, mopt, data):
    """generate_radial_mask (fxs_Projections.py:578-629): (custom mask [n_orders, N_q] or True, data mask); the projection uses their AND.
    Types: none; from_projection_matrices (per order the q range the data's projection matrices cover, `data_projection_matrices_q_id_limits`
    of the invariants record); manual / region; manual / order_dependent_line (half plane of the (order, q) grid left of a line)."""
    nq = len(qs)
    mask = np.full((n_orders, nq), False)
    data_mask = mask | ((qs >= dq.min()) & (qs <= dq.max()))                                           # :585-586
    mtype = mopt['type']
    if mtype == 'none':
        return True, data_mask
    if mtype == 'from_projection_matrices':                                                            # :593-598
        limits = data.get('data_projection_matrices_q_id_limits', False)
        if isinstance(limits, bool) or isinstance(limits.get('I1I1', False), bool):
            raise XfbError("q_mask: from_projection_matrices needs data_projection_matrices_q_id_limits['I1I1'] in the invariants record")
        for part, lim in zip(mask, np.asarray(limits['I1I1'])):
            part[:] = (qs > dq[int(lim[0])]) & (qs < dq[int(lim[1]) - 1])
        return mask, data_mask
    if mtype == 'manual' and mopt['manual']['type'] == 'region':                                       # :603-617
        lo, hi = mopt['manual']['region']
        if (lo == False) and (hi != False):      # noqa: E712
            mask[:] = (qs < hi)[None, :]
        elif (lo != False) and (hi == False):    # noqa: E712
            mask[:] = (qs >= lo)[None, :]
        elif (lo != False) and (hi != False):    # noqa: E712
            mask[:] = ((qs >= lo) & (qs < hi))[None, :]
        else:
            mask[:] = True
        return mask, data_mask
    if mtype == 'manual' and mopt['manual']['type'] == 'order_dependent_line':                         # :618-623, mathLibrary.py:1131-1137
        p1, p2 = (np.asarray(v, dtype=float) for v in mopt['manual']['order_dependent_line'])
        d = p2 - p1
        rot = np.array([d[1], -d[0]])
        grid = np.stack(np.meshgrid(np.arange(n_orders, dtype=float), qs, indexing='ij'), axis=-1) - p1
        return (-(grid * rot[None, None, :]).sum(axis=-1)) >= 0, data_mask
    raise XfbError(f"q_mask type '{mtype}' is not known (none, from_projection_matrices, manual / region, manual / order_dependent_line)")


class ProjectionSetup:
    """Final V_l, radial mask and integrated intensity for a reconstruction grid `qs`."""

    def __init__(self, qs, data, l_max, ropt):
        qs = np.asarray(qs, dtype=np.float64)
        dq = np.asarray(data['data_radial_points'], dtype=np.float64)
        avg = np.asarray(getattr(data['average_intensity'], 'data', data['average_intensity']), dtype=np.float64)
        self.integrated_intensity = float(midpoint_rule(avg * dq ** 2, dq) * 2 * np.sqrt(np.pi))      # :476
        used = np.asarray(ropt['used_order_ids'])
        used = used[(used <= int(data['max_order'])) & (used <= l_max)]                               # :548-559
        n_used = min(len(used), l_max + 1)
        if list(used[:n_used]) != list(range(n_used)):
            raise XfbError("xframe_b200 supports used_order_ids = arange(n) (the only setting under which the "
                           "reference indexes its projection matrices consistently, fxs_Projections.py:837-840)")
        self.n_used = n_used
        if ropt.get('SO_freedom', {}).get('use', False):
            raise XfbError("projections.reciprocal.SO_freedom.use is a 2-D option (default False in 3-D, default_0.01.yaml:185-191); "
                           "the reference's 3-D variant only drops one imaginary part (fxs_Projections.py:766-786) and is not implemented")
        self.number_of_particles = float(ropt['number_of_particles']['initial'])
        pms = [np.asarray(data['data_projection_matrices'][i]) for i in range(n_used)]
        same = dq.shape == qs.shape and bool((dq == qs).all())                                         # :642-651
        if not same:
            # the reference regrids with scipy griddata column by column (gridLibrary.py:635-651), fill value 0
            from scipy.interpolate import griddata
            kind = ropt.get('regrid', {}).get('interpolation', 'cubic')
            avg = griddata(dq[:, None], avg, qs[:, None], method=kind, fill_value=0.0).reshape(len(qs))
            pms = [np.stack([griddata(dq[:, None], col, qs[:, None], method=kind, fill_value=0.0).reshape(len(qs))
                             for col in np.moveaxis(p, 1, 0)], axis=1) for p in pms]
        proj = [np.array(p, dtype=complex) for p in pms]
        if ropt.get('odd_orders_to_0', False):                                                         # :693-698
            for l in range(1, n_used, 2):
                proj[l][:] = 0
        if ropt.get('use_averaged_intensity', False):                                                  # :700-708
            proj[0] = (avg[:, None] * 2 * np.sqrt(np.pi)).astype(complex)
        for p in proj:                                                                                 # :711-713
            p *= 2
        self.projection_matrices = proj
        nq = len(qs)
        mask, data_mask = radial_mask(qs, dq, l_max + 1, ropt.get('q_mask', {'type': 'none'}), data)
        self.radial_mask = np.ascontiguousarray(np.broadcast_to(mask & data_mask, (l_max + 1, nq)))

    def apply_to(self, plan, sv_cutoff=1e-15):
        plan.set_projection(self.projection_matrices, self.radial_mask, self.number_of_particles, sv_cutoff=sv_cutoff)

    def masked_projection_matrices(self):                                                              # reconstruct.py:998-1002
        out = []
        for l, p in enumerate(self.projection_matrices):
            t = np.array(p)
            t[~self.radial_mask[l]] = 0
            out.append(t)
        return out


class ProjectionSetup2D:
    """2-D flavour: V_m(q) [n_orders, N_r] complex, radial mask, integrated intensity
    (fxs_Projections.py:473-474,657-661,679-706,578-629 with dimensions == 2)."""

    def __init__(self, qs, data, m_max, ropt):
        qs = np.asarray(qs, dtype=np.float64)
        dq = np.asarray(data['data_radial_points'], dtype=np.float64)
        avg = np.asarray(getattr(data['average_intensity'], 'data', data['average_intensity']), dtype=np.float64)
        self.integrated_intensity = float(midpoint_rule(avg * dq, dq) * 2 * np.sqrt(np.pi))            # :473-474
        used = np.asarray(ropt['used_order_ids'])
        used = used[(used <= int(data['max_order'])) & (used <= m_max)]
        n_used = min(len(used), m_max + 1)
        if list(used[:n_used]) != list(range(n_used)):
            raise XfbError("xframe_b200 supports used_order_ids = arange(n)")
        self.n_used = n_used
        self.number_of_particles = float(ropt['number_of_particles']['initial'])
        pms = np.asarray(data['data_projection_matrices'])[:n_used]
        same = dq.shape == qs.shape and bool((dq == qs).all())
        if not same:
            from scipy.interpolate import griddata
            kind = ropt.get('regrid', {}).get('interpolation', 'cubic')
            avg = griddata(dq[:, None], avg, qs[:, None], method=kind, fill_value=0.0).reshape(len(qs))
            pms = np.array([griddata(dq[:, None], p, qs[:, None], method=kind, fill_value=0.0).reshape(len(qs)) for p in pms])
        proj = np.array(pms, dtype=complex)
        if ropt.get('odd_orders_to_0', False):
            proj[1::2, :] = 0
        if ropt.get('use_averaged_intensity', False):
            proj[0] = avg.astype(complex)
        self.projection_matrices = proj                                                                # no *2 in 2-D (:710-713 is 3-D only)
        mask, data_mask = radial_mask(qs, dq, m_max + 1, ropt.get('q_mask', {'type': 'none'}), data)
        self.radial_mask = np.ascontiguousarray(np.broadcast_to(mask & data_mask, (m_max + 1, len(qs))))
        so = ropt.get('SO_freedom', {'use': False})
        self.use_SO_freedom = bool(so.get('use', False))
        self.radial_high_pass = so.get('radial_high_pass', 0.2)
        self.qs = qs
        self.so_order_id = int(self.rank_orders()[0][0]) if self.use_SO_freedom else None           # :964-971

    def rank_orders(self):
        """rank_projection_matrix_orders_2d (fxs_Projections.py:933-962): even non-zero orders by mean |V_m(q)| q beyond the
        radial high pass."""
        hp = int((len(self.qs) - 1) * self.radial_high_pass)
        orders = np.arange(self.n_used)
        order_mask = (orders % 2 == 0) & (orders != 0)
        metric = np.mean(np.abs(self.projection_matrices[order_mask, hp:].T * self.qs[hp:, None]), axis=0)
        sorted_indices = np.argsort(metric)[::-1]
        so_ids = order_mask.nonzero()[0][sorted_indices]
        return so_ids, orders[so_ids], sorted_indices

    def remaining_so_rotation(self, n_angular_points):
        """generate_remaining_SO_projection_2D (fxs_Projections.py:1022-1095): returns f(unknowns) -> rotation phase; the
        harmonic coefficients of order m are then multiplied by exp(i m phase)."""
        orders = np.arange(self.n_used)
        order_mask = (orders % 2 == 0) & (orders != 0)
        harmonic_orders = orders[order_mask]
        max_order = np.max(harmonic_orders)
        _, so_orders, sorted_order_indices = self.rank_orders()
        remaining = current = so_orders[0]
        free = True
        steps = []
        while remaining > 2:
            multiples = np.arange(current, max_order + 1, current)
            free = free * ~np.isin(sorted_order_indices, np.where(np.isin(harmonic_orders, multiples)))
            if not np.any(free):
                break
            idx = sorted_order_indices[free][0]
            current = harmonic_orders[idx]
            gcd = np.gcd(remaining, current)
            n_ind = remaining / gcd
            angle = 2 * np.pi / n_ind
            coeff = np.argmin((np.arange(1, n_ind) * current / gcd) % n_ind) + 1
            steps.append((idx, angle, coeff, gcd))
            remaining = gcd
        projection_orders = np.concatenate((np.arange(int(n_angular_points / 2) + 1),
                                            -1 * np.arange(int(n_angular_points / 2) + n_angular_points % 2)[:0:-1]))

        def rotation_phase(unknowns):
            phases = (-1.j * np.log(np.asarray(unknowns)[order_mask])).real
            rot = 0.0
            for idx, angle, coeff, gcd in steps:
                rot -= (phases[idx] // angle) * coeff * angle / gcd
            return rot
        return rotation_phase, projection_orders

    def apply_to(self, plan, sv_cutoff=None):
        plan.set_projection_2d(self.projection_matrices, self.radial_mask, self.number_of_particles, self.so_order_id)

    def masked_projection_matrices(self):
        t = np.array(self.projection_matrices)
        t[~self.radial_mask[:len(t)]] = 0
        return t


def disk_model_density(plan, centers=None, radius=70.0, densities=(25, 50, 25, 50, 25, 50)):
    """2-D projection of the bundled tutorial model (settings/simulate_ccd/tutorial.yaml:11-20): six discs."""
    if centers is None:
        centers = [(0.0, 0.0)] + [(140.0, k * 2 * np.pi / 5) for k in range(5)]
    r, p = plan.rs[:, None], plan.phis[None, :]
    x, y = r * np.cos(p), r * np.sin(p)
    rho = np.zeros(plan.grid_shape)
    for (cr, cp), d in zip(centers, densities):
        rho += d * (np.sqrt((x - cr * np.cos(cp)) ** 2 + (y - cr * np.sin(cp)) ** 2) < radius)
    return rho


def invariants_from_density_2d(plan, density):
    """I_m(q) of |FT rho|^2 as the 2-D invariants record: V_m(q) = I_m(q), <I>(q) = I_0(q); transforms on the GPU."""
    import torch
    d = torch.from_numpy(np.ascontiguousarray(density.astype(complex)))[None].to(plan.device)
    fd = plan.ft(d)
    inten = (fd * fd.conj()).real.to(torch.complex128).contiguous()
    I = plan.sht_forward(inten)[0].cpu().numpy()[:, :plan.l_max + 1]        # rfft half
    return {'dimensions': 2, 'xray_wavelength': 1.23984, 'average_intensity': I[:, 0].real.copy(), 'data_radial_points': plan.qs.copy(),
            'data_angular_points': plan.phis.copy(), 'max_order': plan.l_max, 'data_projection_matrices': np.ascontiguousarray(I.T),
            'number_of_particles': 1}


def initial_support(plan, sup_opt):
    """max_radius initial support on the plan's real grid (fxs_Projections.py:137-140)."""
    if sup_opt['type'] != 'max_radius':
        raise XfbError(f"initial_support type '{sup_opt['type']}' is not supported by xframe_b200 (max_radius)")
    r = plan.rs.reshape((-1,) + (1,) * (len(plan.grid_shape) - 1))
    return np.broadcast_to(r < sup_opt['max_radius'], plan.grid_shape).copy()


def bump(r, radius, slope):                                      # mathLibrary.py:1456-1466
    v = np.zeros_like(r)
    nz = (r > -radius) & (r < radius)
    v[nz] = np.exp(-slope * radius ** 2 / (radius ** 2 - r[nz] ** 2))
    return v


def integrate(plan, values):
    """SphericalIntegrator.integrate (mathLibrary.py:1223-1235) / PolarIntegrator.integrate (:1254-1262) with the
    plan's quadrature weights."""
    if getattr(plan, 'dims', 3) == 2:
        return float(np.sum(plan.int_weight * values))
    return float(np.sum(plan.int_weight[:, :, None] * values))


def density_guess(plan, dopt, particle_radius, integrated_intensity, rng):
    """'bump' guess with a random amplitude (reconstruct.py:1117-1120,1155-1174); `rng` replaces os.urandom seeding."""
    if dopt['type'] != 'bump' or dopt['amplitude_function'] != 'random':
        raise XfbError("density_guess: xframe_b200 supports type 'bump' with amplitude_function 'random'")
    radius = dopt.get('radius', particle_radius)
    if isinstance(radius, bool):
        radius = particle_radius
    if radius < 0:
        radius = float(plan.rs.max())
    A = 1 + 1 / dopt['random']['SNR'] * rng.random(plan.grid_shape)
    r = np.broadcast_to(plan.rs.reshape((-1,) + (1,) * (len(plan.grid_shape) - 1)), plan.grid_shape)
    density = A * bump(np.ascontiguousarray(r), radius, dopt['bump']['slope'])
    total = integrate(plan, density * density)
    return (density * np.sqrt(integrated_intensity / total)).astype(complex)


def six_sphere_density(plan, centers=None, radius=70.0, densities=(25, 50, 25, 50, 25, 50)):
    """The bundled tutorial model (settings/simulate_ccd/tutorial.yaml:11-20)."""
    if centers is None:
        centers = [(0.0, 0.0, 0.0)] + [(140.0, np.pi / 2, k * 2 * np.pi / 5) for k in range(5)]
    r = plan.rs[:, None, None]
    t = plan.thetas[None, :, None]
    p = plan.phis[None, None, :]
    x, y, z = r * np.sin(t) * np.cos(p), r * np.sin(t) * np.sin(p), r * np.cos(t) + 0 * p
    rho = np.zeros(plan.grid_shape)
    for (cr, ct, cp), d in zip(centers, densities):
        c = (cr * np.sin(ct) * np.cos(cp), cr * np.sin(ct) * np.sin(cp), cr * np.cos(ct))
        rho += d * (np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) < radius)
    return rho


def invariants_from_density(plan, density):
    """density -> B_l -> V_l record in the layout `reconstruct` loads; transforms on the GPU.

    V_l is divided by 2 to cancel fxs_Projections.py:711-713, as SURVEY.md section 8d prescribes for synthetic inputs.
    """
    import torch
    d = torch.from_numpy(np.ascontiguousarray(density.astype(complex)))[None].to(plan.device)
    fd = plan.ft(d)
    inten = (fd * fd.conj()).contiguous()
    I = plan.sht_forward(inten)[0].cpu().numpy()                 # [N_r, (L+1)^2]
    pms, Bl = [], []
    for l in range(plan.l_max + 1):
        Il = I[:, l * l:(l + 1) * (l + 1)]
        b = Il @ Il.conj().T                                     # fxs_invariant_tools.py:915-923
        Bl.append(b)
        b = ((b + b.T.conj()) / 2).real
        ev, evec = np.linalg.eigh(b)
        order = np.argsort(ev)[::-1]
        ev, evec = ev[order], evec[:, order]
        n = min(len(evec), 2 * l + 1)
        ev, evec = ev[:n].copy(), evec[:, :n].copy()
        neg = ev < 0
        ev[neg] = 0
        evec[:, neg] = 0
        pms.append(((evec @ np.diag(np.sqrt(ev))) / 2).astype(complex))
    avg = np.sqrt(np.diag(Bl[0]).real / (4 * np.pi))             # simulate_ccd.py:227-230
    return {'dimensions': 3, 'xray_wavelength': 1.23984, 'average_intensity': avg, 'data_radial_points': plan.qs.copy(),
            'data_angular_points': plan.phis.copy(), 'max_order': plan.l_max, 'data_projection_matrices': pms,
            'number_of_particles': 1}
