"""Harmonic-transform forward/inverse interface of the fxs path, backed by the CUDA library.

Mirrors, member for member,
  * the plugin class ``sh`` the reference injects at ``xframe.lib.math.shtns``
    (xframe/externalLibraries/shtns_plugin.py:11-274, interface library/interfaces.py:13-38), and
  * ``HarmonicTransform`` (xframe/projects/fxs/projectLibrary/harmonic_transforms.py:11-96).

Arrays in and out are numpy (like the reference) or CUDA torch tensors (device fast path: no host copies; the
'direct' ordering then never leaves the GPU).  Index bookkeeping (cplx_l_indices, cplx_m_indices, split indices) is
host-only and available without a GPU; the first transform call creates the device plan and raises ``XfbError``
when no GPU / library is present -- there is no CPU fallback.
"""
import numpy as np

from . import tables
from ._lib import XfbError


def _is_torch(x):
    try:
        import torch
        return isinstance(x, torch.Tensor)
    except ImportError:       # pragma: no cover
        return False


class sh:
    """Drop-in for shtns_plugin.sh: orthonormal complex SHT on a Gauss grid (north -> south), index l(l+1)+m."""

    _CAPACITY_BYTES = 1 << 30      # shells transformed per device call are bounded by this much grid data

    def __init__(self, l_max, mode_flag='complex', output_order='l', anti_aliazing_degree=2, n_phi=False, n_theta=False,
                 device=None):
        l_max = int(l_max)
        self.l_max = l_max
        self.anti_aliazing_degree = anti_aliazing_degree
        self.n_coeff = (l_max + 1) ** 2
        self.mode = mode_flag
        if mode_flag != 'complex':
            # the 3-D fxs path uses mode 'complex' only (reconstruct.py:345-350); 'real' is shtns' analys/synth
            raise XfbError("xframe_b200.sh: only mode_flag='complex' is implemented (the mode the fxs 3-D path uses)")
        self.n_theta, self.n_phi = tables.default_angular_sizes(l_max, n_theta, n_phi)
        self._cos_theta, self._gauss_w = tables.gauss_grid(self.n_theta)
        self._phi = 2 * np.pi * np.arange(self.n_phi) / self.n_phi          # shtns_plugin.py:132
        self._theta = np.arccos(self._cos_theta)                            # shtns_plugin.py:133
        # shtns_plugin.py:105-114
        ls = np.arange(l_max + 1, dtype=int)
        ms = np.concatenate((ls, -ls[:0:-1]))
        self.m, self.l = ms, ls
        self.cplx_m_indices = [ls[np.abs(m):] * (ls[np.abs(m):] + 1) + m for m in ms]
        self.cplx_l_indices = [slice(l ** 2, l ** 2 + 2 * l + 1) for l in range(l_max + 1)]
        self.cplx_m_indices_concat = np.concatenate(self.cplx_m_indices)
        self.cplx_l_split_indices = np.arange(1, l_max + 1) ** 2
        lp1 = np.arange(l_max + 2)
        index = (lp1 * (lp1 + 1) / 2).astype(int)                           # shtns_plugin.py:269-274
        self.cplx_m_split_indices = np.concatenate((index[-1] - index[-2::-1], index[-1] + index[1:-2]))
        self._device = device
        self._plan = None

    phi = property(lambda self: self._phi)
    theta = property(lambda self: self._theta)

    @property
    def grid(self):
        """(theta, phi) mesh, theta-major, like GridFactory.construct_grid('uniform', (thetas, phis)) (:134)."""
        t, p = np.meshgrid(self._theta, self._phi, indexing='ij')
        return np.stack((t, p), axis=-1)

    # -- device plumbing ------------------------------------------------------------------------------
    def _get_plan(self):
        if self._plan is None:
            from .plan import Plan
            shell_bytes = self.n_theta * self.n_phi * 16
            cap = int(max(8, min(8192, self._CAPACITY_BYTES // shell_bytes)))
            # SHT only: the radial size of the plan is irrelevant, 8 radial points keep its Hankel tables tiny
            self._plan = Plan(self.l_max, 8, 1.0, n_theta=self.n_theta, n_phi=self.n_phi, max_batch=(cap + 7) // 8, device=self._device)
            self._cap = ((cap + 7) // 8) * 8
        return self._plan

    def attach_plan(self, plan):
        """Share the tables of an existing Plan (same l_max / angular grid) instead of building a private one."""
        if (plan.l_max, plan.n_theta, plan.n_phi) != (self.l_max, self.n_theta, self.n_phi):
            raise ValueError("plan geometry differs from this transform")
        self._plan, self._cap = plan, plan.max_batch * plan.n_r
        return self

    def _to_dev(self, a):
        import torch
        plan = self._get_plan()
        if _is_torch(a):
            return a.to(device=plan.device, dtype=torch.complex128).contiguous(), True
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.complex128)).to(plan.device), False

    def _analysis_dev(self, data):
        """[..., n_theta, n_phi] -> [..., (L+1)^2] on the device, chunked to the plan capacity."""
        import torch
        plan = self._get_plan()
        lead = data.shape[:-2]
        flat = data.reshape(-1, self.n_theta, self.n_phi)
        if flat.shape[0] <= self._cap:
            out = plan.sht_forward(flat)
        else:
            out = torch.cat([plan.sht_forward(flat[i:i + self._cap].contiguous()) for i in range(0, flat.shape[0], self._cap)])
        return out.reshape(*lead, self.n_coeff)

    def _synthesis_dev(self, coeff):
        import torch
        plan = self._get_plan()
        lead = coeff.shape[:-1]
        flat = coeff.reshape(-1, self.n_coeff)
        if flat.shape[0] <= self._cap:
            out = plan.sht_inverse(flat)
        else:
            out = torch.cat([plan.sht_inverse(flat[i:i + self._cap].contiguous()) for i in range(0, flat.shape[0], self._cap)])
        return out.reshape(*lead, self.n_theta, self.n_phi)

    @staticmethod
    def _out(t, was_torch):
        return t if was_torch else t.cpu().numpy()

    # -- analysis: shtns_plugin.py:151-176,218-229,250-255 ----------------------------------------------
    def _analysis(self, data):
        d, tt = self._to_dev(data)
        if d.shape[-2:] != (self.n_theta, self.n_phi):
            raise ValueError(f"spatial shape {tuple(d.shape[-2:])} != grid {(self.n_theta, self.n_phi)}")
        return self._analysis_dev(d), tt

    def forward_d(self, data):
        """'direct': [N_r, n_theta, n_phi] -> [N_r, (L+1)^2]."""
        c, tt = self._analysis(data)
        return self._out(c, tt)

    def forward_l(self, data):
        """'lm': list over l of [..., 2l+1] (columns m = -l..l)."""
        c, tt = self._analysis(data)
        return [self._out(c[..., idx].contiguous(), tt) for idx in self.cplx_l_indices]

    def forward_m(self, data):
        """'ml': list over m in (0..L, -L..-1) of [..., L-|m|+1] (rows l = |m|..L)."""
        import torch
        c, tt = self._analysis(data)
        return [self._out(c[..., torch.as_tensor(idx, device=c.device)].contiguous(), tt) for idx in self.cplx_m_indices]

    # -- synthesis: shtns_plugin.py:179-194,230-238,257-261 ---------------------------------------------
    def inverse_d(self, data):
        d, tt = self._to_dev(data)
        return self._out(self._synthesis_dev(d), tt)

    def inverse_l(self, data):
        import torch
        parts = [self._to_dev(p) for p in data]
        tt = all(t for _, t in parts)
        full = torch.cat([p for p, _ in parts], dim=-1)          # np.concatenate(data, axis=1) for [N_r, 2l+1] blocks
        return self._out(self._synthesis_dev(full.contiguous()), tt)

    def inverse_m(self, data):
        import torch
        parts = [self._to_dev(p) for p in data]
        tt = all(t for _, t in parts)
        lead = parts[0][0].shape[:-1]
        full = torch.zeros(tuple(lead) + (self.n_coeff,), dtype=torch.complex128, device=parts[0][0].device)
        for (p, _), index in zip(parts, self.cplx_m_indices):
            full[..., torch.as_tensor(index, device=full.device)] = p
        return self._out(self._synthesis_dev(full), tt)

    def m_to_l_ordering(self, m_coeff):                            # shtns_plugin.py:240-247
        l_coeff = np.zeros_like(m_coeff)
        pos = 0
        for index in self.cplx_m_indices:
            l_coeff[index] = m_coeff[pos:pos + len(index)]
            pos += len(index)
        return l_coeff

    def test(self, data):                                          # shtns_plugin.py:263-267
        d, tt = self._to_dev(data + 0.j if not _is_torch(data) else data)
        return self._out(self._synthesis_dev(self._analysis_dev(d)), tt)

    # shtns_plugin.py:85-103
    def max_order_from_n_angular_steps(self, n_phi):
        n_phi = 2 ** int(np.log2(n_phi))
        return n_phi // (self.anti_aliazing_degree + 1)

    def n_angular_step_from_max_order(self, max_order):
        n_phi = 2 ** (int(np.log2((self.anti_aliazing_degree + 1) * max_order)) + 1)
        return {'n_phi': n_phi, 'n_theta': n_phi // 2}


def get_spherical_harmonic_transform_obj(l_max, mode='complex', anti_aliazing_degree=2, n_phi=False, n_theta=False, device=None):
    """mathLibrary.get_spherical_harmonic_transform_obj (mathLibrary.py:28-31 + the 'shtns' dependency slot)."""
    return sh(l_max, mode_flag=mode, anti_aliazing_degree=anti_aliazing_degree, n_phi=n_phi, n_theta=n_theta, device=device)


class HarmonicTransform:
    """harmonic_transforms.py:11-96.  ``opt`` keys: dimensions, max_order, n_phi, n_theta, anti_aliazing_degree, indices."""

    def __init__(self, data_type, opt):
        self.data_type = data_type
        self.opt = opt
        self.dim = opt['dimensions']
        ht, iht, grid_param, trf_by_indices = self.chose_transforms()
        self.transforms_by_indices = trf_by_indices
        self.forward = ht
        self.inverse = iht
        self.grid_param = grid_param
        self.max_order = opt['max_order']

    @classmethod
    def from_data_array(cls, data_type, array):                    # harmonic_transforms.py:23-31
        shape, dimension = array.shape, array.ndim
        if dimension == 2:
            opt = {'dimensions': dimension, 'max_order': False, 'n_angular_points': shape[1]}
        else:
            opt = {'dimensions': dimension, 'max_order': shape[1] - 1, 'n_phi': shape[2], 'n_theta': shape[1]}
        return cls(data_type, opt)

    def chose_transforms(self):
        opt = self.opt
        if self.dim == 2:
            from .circular import CircularHarmonicTransform
            max_order = opt.get('max_order', False)
            size = opt['n_angular_points'] if isinstance(max_order, bool) else max_order * 2 + 1   # :44-47
            ch = CircularHarmonicTransform(size, self.data_type, device=opt.get('device'))
            self._ch = ch
            trf = {'m': {'forward': ch.forward, 'inverse': ch.inverse}}
            return ch.forward, ch.inverse, {'phis': np.arange(size) / size * 2 * np.pi}, trf
        if self.dim != 3:
            raise XfbError(f"dimensions={self.dim} not supported")
        l_max = int(opt['max_order'])
        aa = opt.get('anti_aliazing_degree', False)
        kw = {} if isinstance(aa, bool) else {'anti_aliazing_degree': aa}
        obj = get_spherical_harmonic_transform_obj(l_max, mode=self.data_type, n_phi=opt.get('n_phi', 0), n_theta=opt.get('n_theta', 0),
                                                   device=opt.get('device'), **kw)
        self.m_indices, self.l_indices = obj.cplx_m_indices, obj.cplx_l_indices
        self.m, self.l = obj.m, obj.l
        self.m_split_indices, self.l_split_indices = obj.cplx_m_split_indices, obj.cplx_l_split_indices
        self.n_coeff = obj.n_coeff
        self.test = obj.test
        self._sh = obj
        trf = {'lm': {'forward': obj.forward_l, 'inverse': obj.inverse_l},
               'ml': {'forward': obj.forward_m, 'inverse': obj.inverse_m},
               'direct': {'forward': obj.forward_d, 'inverse': obj.inverse_d}}
        indices = opt.get('indices', 'lm')
        return trf[indices]['forward'], trf[indices]['inverse'], {'phis': obj.phi, 'thetas': obj.theta}, trf
