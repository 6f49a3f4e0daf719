"""Plan: owns the device tables / workspaces of one (L, N_r, n_theta, n_phi, batch) configuration and exposes
the hot-path operators on torch CUDA tensors (torch is plumbing only: device memory + streams).

Shapes follow the reference:
  grid   [nb, N_r, n_theta, n_phi] complex128
  direct [nb, N_r, (L+1)^2]       complex128   (index l(l+1)+m; shtns_plugin.py:110-112,250-261)
"""
import ctypes as C
import logging

import numpy as np
import torch

from . import _lib, tables

HIO, ER = 0, 1
_OPS = {'support': 1, 'value_threshold': 2, 'limit_imag': 3, 'average_center': 4}
log = logging.getLogger('root')          # the reference's logger name (fxs_Projections.py:24)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Plan:
    def __init__(self, l_max, n_r, max_q, n_theta=0, n_phi=0, reciprocity_coefficient=2.0, ft_type='midpoint',
                 max_batch=1, device=None, hankel_weights=None, hankel_scales=None, dimensions=3):
        """hankel_weights / hankel_scales: optional caller-supplied radial weights [L+1, n_sum, N_r] (float64) and
        (forward, inverse) prefactors, as the reference's generate_ht receives them (hankel_transforms.py:540-559);
        by default they are computed for (ft_type, reciprocity_coefficient) like generate_weightDict + assemble_weights.
        dimensions=2: polar plan (settings `dimensions: 2`): grids are [.., N_r, n_phi] with n_phi = 2*l_max+1
        (harmonic_transforms.py:44-47,60), coefficients [.., N_r, n_phi] in 'm' order (0..M, -M..-1), hankel_weights
        [M+1, n_sum, N_r] (calc_polar_mid_weights)."""
        self.dims = int(dimensions)
        if self.dims not in (2, 3):
            raise _lib.XfbError(f"dimensions={dimensions} not supported")
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.XfbError("no CUDA device visible: xframe_b200 has no CPU fallback")
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        torch.cuda.set_device(self.device)
        _lib.check(self.lib.xfb_set_device(self.device.index))
        self.l_max, self.n_r = int(l_max), int(n_r)
        self.max_batch = int(max_batch)
        if self.dims == 2:
            self._init_polar(max_q, reciprocity_coefficient, ft_type, hankel_weights, hankel_scales)
            return
        self.n_theta, self.n_phi = tables.default_angular_sizes(self.l_max, n_theta, n_phi)
        self.n_lm = (self.l_max + 1) ** 2
        self.rc, self.ft_type, self.max_q = float(reciprocity_coefficient), ft_type, float(max_q)
        self.rs, self.qs = tables.radial_grids(ft_type, self.max_q, self.n_r, self.rc)
        self.cos_theta, self.gauss_w = tables.gauss_grid(self.n_theta)
        self.thetas = np.arccos(self.cos_theta)                         # shtns_plugin.py:133
        self.phis = 2 * np.pi * np.arange(self.n_phi) / self.n_phi      # shtns_plugin.py:132
        leg, self.NP = tables.pack_legendre(self.l_max, self.n_theta, self.n_phi)
        if hankel_weights is None:
            self.hankel_w = tables.hankel_weights(self.l_max, self.n_r, self.rc, ft_type)
        else:
            self.hankel_w = np.ascontiguousarray(hankel_weights, dtype=np.float64)
            if self.hankel_w.ndim != 3 or self.hankel_w.shape[0] != self.l_max + 1 or self.hankel_w.shape[2] != self.n_r \
                    or self.hankel_w.shape[1] not in (self.n_r, self.n_r - 1):
                raise ValueError(f"hankel_weights shape {self.hankel_w.shape} is not [L+1, N_r or N_r-1, N_r]")
        if hankel_scales is None:
            fs, iscale = tables.hankel_scales(float(np.max(self.rs)), self.n_r, self.rc, ft_type)
        else:
            fs, iscale = (float(v) for v in hankel_scales)
        self.int_weight = tables.integration_weights(self.rs, self.n_theta)
        d = _lib.PlanDesc()
        d.l_max, d.n_r, d.n_theta, d.n_phi, d.max_batch = self.l_max, self.n_r, self.n_theta, self.n_phi, self.max_batch
        d.hankel_skip = self.n_r - self.hankel_w.shape[1]      # trapz / Zernike weights drop the p = 0 row (hankel_transforms.py:326-333)
        keep = [np.ascontiguousarray(a, dtype=np.float64) for a in
                (self.cos_theta, self.gauss_w, leg, self.hankel_w, self.int_weight, self.rs, self.qs)]
        d.cos_theta, d.gauss_w, d.legendre = _dp(keep[0]), _dp(keep[1]), _dp(keep[2])
        d.legendre_len = keep[2].size
        d.hankel_w, d.hankel_n_sum = _dp(keep[3]), self.hankel_w.shape[1]
        d.hankel_fwd_scale, d.hankel_inv_scale = fs, iscale
        d.int_weight, d.r_points, d.q_points = _dp(keep[4]), _dp(keep[5]), _dp(keep[6])
        h = C.c_void_p()
        _lib.check(self.lib.xfb_plan_create(C.byref(h), C.byref(d)))
        self.h = h
        self.n_batch = 0
        self.initial_support = None

    def _init_polar(self, max_q, rc, ft_type, hankel_weights, hankel_scales):
        self.n_theta, self.n_phi = 1, 2 * self.l_max + 1
        self.n_lm = self.n_phi
        self.rc, self.ft_type, self.max_q = float(rc), ft_type, float(max_q)
        self.rs, self.qs = tables.radial_grids(ft_type, self.max_q, self.n_r, self.rc)        # ft_grid_pairs.py:274-291 (same radial rule)
        self.phis = np.arange(self.n_phi) / self.n_phi * 2 * np.pi                            # harmonic_transforms.py:59
        self.thetas = None
        w = tables.polar_hankel_weights(self.l_max, self.n_r, self.rc, ft_type) if hankel_weights is None \
            else np.ascontiguousarray(hankel_weights, dtype=np.float64)
        if w.ndim != 3 or w.shape[0] != self.l_max + 1 or w.shape[2] != self.n_r or w.shape[1] not in (self.n_r, self.n_r - 1):
            raise ValueError(f"hankel_weights shape {w.shape} is not [M+1, N_r or N_r-1, N_r]")
        self.hankel_w = w
        fs, iscale = tables.polar_hankel_scales(float(np.max(self.rs)), self.n_r, self.rc, ft_type) if hankel_scales is None \
            else (float(v) for v in hankel_scales)
        self.int_weight = tables.polar_integration_weights(self.rs, self.phis)
        d = _lib.PlanDesc()
        d.dimensions = 2
        d.l_max, d.n_r, d.n_theta, d.n_phi, d.max_batch = self.l_max, self.n_r, 1, self.n_phi, self.max_batch
        d.hankel_skip = self.n_r - w.shape[1]
        keep = [np.ascontiguousarray(a, dtype=np.float64) for a in
                (tables.polar_hankel_device_weights(w), self.int_weight, self.rs, self.qs)]
        d.hankel_w, d.hankel_n_sum = _dp(keep[0]), w.shape[1]
        d.hankel_fwd_scale, d.hankel_inv_scale = fs, iscale
        d.int_weight, d.r_points, d.q_points = _dp(keep[1]), _dp(keep[2]), _dp(keep[3])
        h = C.c_void_p()
        _lib.check(self.lib.xfb_plan_create(C.byref(h), C.byref(d)))
        self.h = h
        self.n_batch = 0
        self.initial_support = None

    def close(self):
        if getattr(self, 'h', None):
            self.lib.xfb_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _c128(self, t, shape_tail):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.complex128 and t.is_contiguous()):
            raise TypeError("expected a contiguous CUDA complex128 tensor")
        if tuple(t.shape[-len(shape_tail):]) != tuple(shape_tail):
            raise ValueError(f"trailing shape {tuple(t.shape)} does not end with {tuple(shape_tail)}")
        return t

    @property
    def grid_shape(self):
        return (self.n_r, self.n_phi) if self.dims == 2 else (self.n_r, self.n_theta, self.n_phi)

    @property
    def _ang_shape(self):
        return (self.n_phi,) if self.dims == 2 else (self.n_theta, self.n_phi)

    def set_fused_ft_stab(self, on=True):
        """True (default): one inverse transform per ft_stab iteration (linearity of IFT); False: literal sketch."""
        _lib.check(self.lib.xfb_plan_set_fused_ft_stab(self.h, int(bool(on))))

    def set_dual_stream(self, on=True, min_batch=0, big_sms=0, small_sms=0):
        """Two halves of the batch on two streams inside mtip_iterate (Jacobi of one half overlaps the transforms of the other)."""
        _lib.check(self.lib.xfb_plan_set_dual_stream(self.h, int(bool(on)), int(min_batch), int(big_sms), int(small_sms)))

    def set_sht_chunk(self, runs_per_chunk, streams=3):
        """L2-resident phi-Fourier intermediate: runs per transform chunk (0 = unchunked) and number of streams (1..4)."""
        _lib.check(self.lib.xfb_plan_set_sht_chunk(self.h, int(runs_per_chunk), int(streams)))

    def jacobi_sweeps(self):
        """Diagnostics: (orders, sweeps[n_batch, n_orders]) of the last invariant projection."""
        cap = 4096 * 64
        buf = (C.c_int32 * cap)()
        orders = (C.c_int32 * 1024)()
        n = C.c_int32(0)
        _lib.check(self.lib.xfb_debug_jacobi_sweeps(self.h, buf, cap, C.byref(n), orders))
        na = n.value
        arr = np.frombuffer(buf, dtype=np.int32)
        nb = max(1, self.n_batch)
        return list(orders[:na]), arr[:nb * na].reshape(nb, na).copy() if na else np.zeros((nb, 0), np.int32)

    def workspace_bytes(self):
        return int(self.lib.xfb_plan_workspace_bytes(self.h))

    def graph_replays(self):
        return int(self.lib.xfb_plan_graph_replays(self.h))

    def launch_count(self):
        return int(self.lib.xfb_plan_launch_count(self.h))

    # ------------------------------------------------------------------ transforms
    def sht_forward(self, grid):
        """sh.forward_d (shtns_plugin.py:250-255): [..., n_theta, n_phi] -> [..., (L+1)^2]."""
        self._c128(grid, self._ang_shape)
        lead = grid.shape[:-len(self._ang_shape)]
        n = int(np.prod(lead)) if lead else 1
        out = torch.empty(lead + (self.n_lm,), dtype=torch.complex128, device=grid.device)
        _lib.check(self.lib.xfb_sht_forward(self.h, _ptr(grid), _ptr(out), n, _stream()))
        return out

    def sht_inverse(self, direct):
        """sh.inverse_d (shtns_plugin.py:257-261)."""
        self._c128(direct, (self.n_lm,))
        lead = direct.shape[:-1]
        n = int(np.prod(lead)) if lead else 1
        out = torch.empty(lead + self._ang_shape, dtype=torch.complex128, device=direct.device)
        _lib.check(self.lib.xfb_sht_inverse(self.h, _ptr(direct), _ptr(out), n, _stream()))
        return out

    def _nb(self, t, tail):
        self._c128(t, tail)
        lead = t.shape[:-len(tail)]
        return int(np.prod(lead)) if lead else 1

    def hankel(self, direct, inverse=False):
        """zht / izht of generate_spherical_ht_gpu (hankel_transforms.py:660-766) on [nb, N_r, (L+1)^2]."""
        nb = self._nb(direct, (self.n_r, self.n_lm))
        out = torch.empty_like(direct)
        _lib.check(self.lib.xfb_hankel_apply(self.h, 1 if inverse else 0, _ptr(direct), _ptr(out), nb, _stream()))
        return out

    def ft(self, grid, inverse=False):
        """ft / ift of generate_ft (fourier_transforms.py:57-85)."""
        nb = self._nb(grid, self.grid_shape)
        out = torch.empty_like(grid)
        _lib.check(self.lib.xfb_ft(self.h, 1 if inverse else 0, _ptr(grid), _ptr(out), nb, _stream()))
        return out

    def ift(self, grid):
        return self.ft(grid, inverse=True)

    # ------------------------------------------------------------------ projection constants
    def set_projection(self, projection_matrices, radial_mask, number_of_particles=1.0, sv_cutoff=1e-15, max_sweeps=40):
        """projection_matrices: list over used orders of the FINAL V_l (after regrid / odd->0 / V_0 / *2,
        fxs_Projections.py:679-714), shape [N_r, n_l]; must be real (imag == 0)."""
        n = len(projection_matrices)
        vs, ncols = [], (C.c_int32 * n)()
        for l, v in enumerate(projection_matrices):
            v = np.asarray(v)
            if np.iscomplexobj(v):
                if np.abs(v.imag).max() > 0:
                    raise _lib.XfbError(f"projection matrix of order {l} has a non-zero imaginary part: the real-arithmetic "
                                        "Procrustes path of xframe_b200 needs real V_l (as produced by fxs extract)")
                v = v.real
            v = np.ascontiguousarray(v, dtype=np.float64)
            if v.shape[0] != self.n_r:
                raise ValueError(f"order {l}: expected {self.n_r} radial rows, got {v.shape[0]}")
            vs.append(v)
            ncols[l] = v.shape[1]
        arr = (C.POINTER(C.c_double) * n)(*[_dp(v) for v in vs])
        rm = np.ascontiguousarray(np.broadcast_to(np.asarray(radial_mask, dtype=bool), (self.l_max + 1, self.n_r)), dtype=np.uint8)
        d = _lib.ProjectionDesc()
        d.n_orders, d.n_cols, d.v = n, ncols, arr
        d.radial_mask = rm.ctypes.data_as(C.POINTER(C.c_uint8))
        d.sqrt_n_particles = float(np.sqrt(number_of_particles))
        d.sv_cutoff, d.max_sweeps = float(sv_cutoff), int(max_sweeps)
        _lib.check(self.lib.xfb_plan_set_projection(self.h, C.byref(d)))
        self._proj_ncols = [int(c) for c in ncols]

    def set_projection_2d(self, projection_matrices, radial_mask, number_of_particles=1.0, so_order_id=None):
        """projection_matrices: complex [n_orders, N_r], the FINAL V_m(q) (after regrid / odd->0 / V_0 = <I>,
        fxs_Projections.py:679-706); so_order_id pins u[so_order_id] = 1 (SO_freedom, :743-748)."""
        v = np.ascontiguousarray(np.asarray(projection_matrices), dtype=np.complex128)
        if v.ndim != 2 or v.shape[1] != self.n_r or v.shape[0] > self.l_max + 1:
            raise ValueError(f"projection matrices shape {v.shape} is not [n_orders <= {self.l_max + 1}, {self.n_r}]")
        rm = np.ascontiguousarray(np.broadcast_to(np.asarray(radial_mask, dtype=bool), (self.l_max + 1, self.n_r)), dtype=np.uint8)
        _lib.check(self.lib.xfb_plan_set_projection_2d(self.h, v.shape[0], v.ctypes.data_as(C.c_void_p), rm.ctypes.data_as(C.c_void_p),
                                                       float(np.sqrt(number_of_particles)), -1 if so_order_id is None else int(so_order_id)))
        self._proj_ncols = [1] * v.shape[0]

    def unknowns(self, run):
        """fxs_unknowns of the last invariant projection for one run of the batch: tuple over the used orders of
        complex [n_l, 2l+1] arrays (approximate_unknowns, fxs_Projections.py:752-767)."""
        if self.dims == 2:
            t = torch.empty((len(self._proj_ncols),), dtype=torch.complex128, device=self.device)
            _lib.check(self.lib.xfb_get_unknowns_2d(self.h, int(run), _ptr(t), _stream()))
            return t.cpu().numpy()
        out = []
        for l, nc in enumerate(self._proj_ncols):
            t = torch.empty((nc, 2 * l + 1), dtype=torch.complex128, device=self.device)
            _lib.check(self.lib.xfb_get_unknowns(self.h, int(run), l, _ptr(t), _stream()))
            out.append(t)
        return tuple(t.cpu().numpy() for t in out)

    def set_real(self, apply, initial_support, value_threshold=(0, False), limit_imag=2.0, considered=('all',),
                 error_inside_initial_support=True, average_center_shells=1):
        """Options of RealProjection (fxs_Projections.py:72-130) and of the real l2 error (fxs_IO_methods.py:287-300).
        Names without a generate_<name>_projection in the reference are logged and ignored, as assemble_projection does
        (fxs_Projections.py:113-118): the reference's own default `apply` list carries such a name ('assert_real')."""
        d = _lib.RealDesc()
        kept = []
        for name in apply:
            if name in _OPS:
                kept.append(name)
            else:
                log.error('projection {} not known. Ignoring it.'.format(name))
        apply = kept
        if len(apply) > 4:
            raise ValueError("at most 4 real projections")
        self.real_projections = tuple(apply)
        d.n_ops = len(apply)
        for i, name in enumerate(apply):
            d.ops[i] = _OPS[name]
            d.hio_considered[i] = 1 if ('all' in considered or name in considered) else 0
        d.average_center_shells = int(average_center_shells)

        def num(v):
            return isinstance(v, (float, int)) and not isinstance(v, bool)
        d.use_lo, d.use_hi = int(num(value_threshold[0])), int(num(value_threshold[1]))
        d.lo = float(value_threshold[0]) if d.use_lo else 0.0
        d.hi = float(value_threshold[1]) if d.use_hi else 0.0
        d.imag_limit = float(limit_imag)
        d.error_inside_initial_support = int(bool(error_inside_initial_support))
        sup = np.ascontiguousarray(np.asarray(initial_support, dtype=bool), dtype=np.uint8)
        if sup.shape != self.grid_shape:
            raise ValueError(f"initial_support shape {sup.shape} != {self.grid_shape}")
        self.initial_support = sup.astype(bool)
        _lib.check(self.lib.xfb_plan_set_real(self.h, C.byref(d), sup.ctypes.data_as(C.c_void_p)))

    # ------------------------------------------------------------------ operators
    def project_invariants(self, direct):
        """approximate_unknowns + mtip_projection (fxs_Projections.py:752-872) on [nb, N_r, (L+1)^2]."""
        nb = self._nb(direct, (self.n_r, self.n_lm))
        out = torch.empty_like(direct)
        _lib.check(self.lib.xfb_project_invariants(self.h, _ptr(direct), _ptr(out), nb, _stream()))
        return out

    # ------------------------------------------------------------------ degree-2 invariants
    def deg2_invariants(self, direct):
        """harmonic_coeff_to_deg2_invariants_3d (fxs_invariant_tools.py:915-923) for the coefficients of a REAL field:
        direct [nb, N_r, (L+1)^2] complex -> B_l [nb, L+1, N_r, N_r] float64 (real symmetric)."""
        nb = self._nb(direct, (self.n_r, self.n_lm))
        out = torch.empty((nb, self.l_max + 1, self.n_r, self.n_r), dtype=torch.float64, device=direct.device)
        _lib.check(self.lib.xfb_deg2_invariants(self.h, _ptr(direct), _ptr(out), nb, _stream()))
        return out

    def set_deg2_reference(self, reference_invariants, radial_mask, number_of_particles=1.0):
        """Constants of deg2_invariant_l2_diff (_generate_deg2_invariant_diff_3d, fxs_IO_methods.py:412-447):
        reference_invariants [L+1, N_r, N_r] = V_l V_l^H of the final projection matrices (fxs_Projections.py:631-637)."""
        ref = np.asarray(reference_invariants)
        if np.iscomplexobj(ref):
            if np.abs(ref.imag).max() > 0:
                raise _lib.XfbError("deg2 reference invariants must be real (real projection matrices)")
            ref = ref.real
        rm = np.broadcast_to(np.asarray(radial_mask, dtype=bool), (self.l_max + 1, self.n_r))
        ref = np.where(rm[:, :, None] & rm[:, None, :], ref, 0.0)                       # reference_masked[mask] = 0  (:422-424)
        norm = np.ascontiguousarray((ref * ref).sum(axis=(1, 2)), dtype=np.float64)    # :426
        ref = np.ascontiguousarray(ref, dtype=np.float64).copy()
        ref[0] = ref[0] / float(number_of_particles)                                    # :440
        _lib.check(self.lib.xfb_plan_set_deg2_reference(self.h, _dp(ref), _dp(norm)))

    def deg2_invariant_diff(self, direct):
        """Per-order error array [nb, L+1] of deg2_invariant_l2_diff for coefficient arrays [nb, N_r, (L+1)^2]."""
        nb = self._nb(direct, (self.n_r, self.n_lm))
        out = torch.empty((nb, self.l_max + 1), dtype=torch.float64, device=direct.device)
        _lib.check(self.lib.xfb_deg2_invariant_diff(self.h, _ptr(direct), _ptr(out), nb, _stream()))
        return out

    def mtip_enable_deg2_metric(self, on, capacity=1):
        self._deg2_cap = int(capacity)
        _lib.check(self.lib.xfb_mtip_enable_deg2_metric(self.h, int(bool(on)), int(capacity)))

    def mtip_deg2_errors(self, n_done):
        out = torch.empty((self.n_batch, self._deg2_cap, self.l_max + 1), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.xfb_mtip_get_deg2_errors(self.h, _ptr(out), self._deg2_cap, _stream()))
        return out[:, :min(int(n_done), self._deg2_cap)]

    def modify_intensity(self, rho_hat, i_proj):
        nb = self._nb(rho_hat, self.grid_shape)
        self._c128(i_proj, self.grid_shape)
        out = torch.empty_like(rho_hat)
        _lib.check(self.lib.xfb_modify_intensity(self.h, _ptr(rho_hat), _ptr(i_proj), _ptr(out), nb, _stream()))
        return out

    def real_update(self, method, beta, rho_ift, rho_prev, support, rho_rt=None, enforce=None):
        """-> (rho_next, err[nb,2]) ; support: uint8/bool [nb, grid] (1 inside the shrink-wrap support)."""
        nb = self._nb(rho_ift, self.grid_shape)
        sup = support.to(torch.uint8).contiguous()
        out = torch.empty_like(rho_ift)
        err = torch.empty((nb, 2), dtype=torch.float64, device=rho_ift.device)
        enf = None if enforce is None else enforce.to(torch.int32).contiguous()
        _lib.check(self.lib.xfb_real_update(self.h, int(method), float(beta), _ptr(rho_ift), _ptr(rho_rt), _ptr(rho_prev),
                                            _ptr(sup), _ptr(enf), _ptr(out), _ptr(err), nb, _stream()))
        return out, err

    def shrinkwrap(self, rho, sigma, threshold):
        nb = self._nb(rho, self.grid_shape)
        out = torch.empty(rho.shape, dtype=torch.uint8, device=rho.device)
        _lib.check(self.lib.xfb_shrinkwrap(self.h, _ptr(rho), float(sigma), float(threshold), _ptr(out), nb, _stream()))
        return out.bool()

    # ------------------------------------------------------------------ device-resident loop
    def mtip_init(self, rho0):
        nb = self._nb(rho0, self.grid_shape)
        _lib.check(self.lib.xfb_mtip_init(self.h, _ptr(rho0), nb, _stream()))
        self.n_batch = nb

    def mtip_iterate(self, method, ft_stab, betas):
        betas = np.ascontiguousarray(np.atleast_1d(np.asarray(betas, dtype=np.float64)))
        _lib.check(self.lib.xfb_mtip_iterate(self.h, int(method), int(bool(ft_stab)), betas.size, _dp(betas), _stream()))

    def mtip_step_host(self, method, ft_stab, beta, rho_in, rho_out, err_out):
        """One iteration with HOST tensors (pinned): rho_in/rho_out complex128 [nb, grid], err_out float64 [nb, 2]."""
        for t in (rho_in, rho_out, err_out):
            if t.is_cuda or not t.is_contiguous():
                raise TypeError("mtip_step_host expects contiguous CPU (pinned) tensors")
        _lib.check(self.lib.xfb_mtip_step_host(self.h, int(method), int(bool(ft_stab)), float(beta), _ptr(rho_in), _ptr(rho_out),
                                               _ptr(err_out), _stream()))

    def set_host_chunk(self, runs):
        _lib.check(self.lib.xfb_plan_set_host_chunk(self.h, int(runs)))

    def mtip_shrinkwrap(self, sigma, threshold, error_limit):
        _lib.check(self.lib.xfb_mtip_shrinkwrap(self.h, float(sigma), float(threshold), float(error_limit), _stream()))

    # sketch / option tail (reconstruct.py:529-534,606-613,886-904,945-949); the schedule logic lives in reconstruct.run_schedule
    def mtip_set_outer_iteration(self, it):
        _lib.check(self.lib.xfb_mtip_set_outer_iteration(self.h, int(it)))

    def mtip_select_best(self, n_first):
        _lib.check(self.lib.xfb_mtip_select_best(self.h, int(n_first), _stream()))

    def mtip_snapshot_intensity(self):
        _lib.check(self.lib.xfb_mtip_snapshot_intensity(self.h, _stream()))

    def mtip_fix_intensity(self):
        _lib.check(self.lib.xfb_mtip_fix_intensity(self.h, _stream()))

    def mtip_set_non_fxs(self, on):
        _lib.check(self.lib.xfb_mtip_set_non_fxs(self.h, int(bool(on))))

    def mtip_shrinkwrap_center(self, sigma, threshold, error_limit):
        _lib.check(self.lib.xfb_mtip_shrinkwrap_center(self.h, float(sigma), float(threshold), float(error_limit), _stream()))

    def mtip_grid(self, which):
        names = {'last_real': 0, 'last_reciprocal': 1, 'best_real': 2, 'best_reciprocal': 3, 'last_support': 4, 'best_support': 5}
        w = names[which]
        dt = torch.complex128 if w < 4 else torch.uint8
        out = torch.empty((self.n_batch,) + self.grid_shape, dtype=dt, device=self.device)
        _lib.check(self.lib.xfb_mtip_get_grid(self.h, w, _ptr(out), _stream()))
        return out if w < 4 else out.bool()

    def mtip_nonfinite(self):
        """Diagnostics: per run, the number of iterations whose error metric was NaN / inf."""
        buf = (C.c_int32 * self.n_batch)()
        _lib.check(self.lib.xfb_mtip_get_nonfinite(self.h, buf, _stream()))
        return np.frombuffer(buf, dtype=np.int32).copy()

    def mtip_errors(self, capacity=16384):
        n = C.c_int32(0)
        hist = torch.zeros((self.n_batch, capacity), dtype=torch.float64, device=self.device)
        best = torch.zeros((self.n_batch,), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.xfb_mtip_get_errors(self.h, _ptr(hist), capacity, _ptr(best), C.byref(n), _stream()))
        return hist[:, :min(n.value, capacity)], best

    # ------------------------------------------------------------------ profiling
    def profile(self, on=True):
        _lib.check(self.lib.xfb_profile_enable(self.h, int(on)))

    def profile_read(self):
        nmax = 16
        names = C.create_string_buffer(nmax * 32)
        ms = (C.c_double * nmax)()
        ln = (C.c_int64 * nmax)()
        n = C.c_int32(0)
        _lib.check(self.lib.xfb_profile_read(self.h, nmax, names, ms, ln, C.byref(n)))
        out = {}
        for i in range(n.value):
            nm = names.raw[i * 32:(i + 1) * 32].split(b'\0')[0].decode()
            out[nm] = {'ms': ms[i], 'launches': int(ln[i])}
        return out
