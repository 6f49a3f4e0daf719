"""Radial (Hankel) transform factory of the fxs path, backed by the CUDA library.

Mirrors xframe/projects/fxs/projectLibrary/hankel_transforms.py:
  generate_weightDict (:22-35), assemble_weights (:36-49), generate_ht (:540-559) and the 'GPU flavour'
  generate_spherical_ht_gpu (:660-766), whose zht / izht map complex128 [N_r, (L+1)^2] -> same.

The reference folds the prefactor (-+i)^l (dr)^3 sqrt(2/pi) into complex weight arrays [p, k, l] and contracts them
with a numpy broadcast-sum or an OpenCL kernel.  Here the REAL weights w[l, p, k] go to the device once and the
prefactor is applied in the GEMM epilogue (csrc/gemm.cuh:hankel_kernel); `assemble_weights` is still provided with
the reference's output layout because other xFrame code reads it.
"""
import numpy as np

from . import tables
from ._lib import XfbError

ht_modes = ['trapz', 'Zernike', 'midpoint', 'gauss']          # hankel_transforms.py:17


def _q_max_from_r_max(r_max, n_r, rc):
    """polar_spherical_dft_reciprocity_relation_radial_cutoffs (mathLibrary.py:1169-1176)."""
    return rc * n_r / r_max


def generate_weightDict(max_order, n_radial_points, reciprocity_coefficient=np.pi, dimensions=3, n_cpus=False, mode=ht_modes[0], **kwargs):
    """{'weights': [orders, summed radial index, new radial index], 'posHarmOrders', 'mode'} (:22-35, :377-397, :299-320)."""
    if mode not in ('trapz', 'midpoint', 'gauss'):
        raise XfbError(f"Hankel weights mode '{mode}' is not implemented by xframe_b200 (trapz, midpoint, gauss)")
    if dimensions == 3:
        w = tables.hankel_weights(int(max_order), int(n_radial_points), float(reciprocity_coefficient), mode)
    elif dimensions == 2:
        w = tables.polar_hankel_weights(int(max_order), int(n_radial_points), float(reciprocity_coefficient), mode)
    else:
        raise XfbError(f"dimensions={dimensions} not supported")
    return {'weights': w, 'posHarmOrders': np.arange(int(max_order) + 1), 'mode': mode}


def assemble_weights(weights, pos_orders, r_max, reciprocity_coefficient=np.pi, dimensions=3, mode=ht_modes[0]):
    """Complex forward / inverse weight arrays [p, k, order] exactly as the reference assembles them
    (assemble_weights_trapz :349-375, assemble_weights_mid :426-452, assemble_weights_gauss :505-535)."""
    weights = np.asarray(weights)
    n_r = weights.shape[-1]
    orders = np.arange(weights.shape[0]) if mode == ht_modes[0] else np.asarray(pos_orders)
    q_max = _q_max_from_r_max(r_max, n_r, reciprocity_coefficient)
    div = 2 if mode == 'gauss' else n_r             # gauss: the [-1,1] -> [0,r_max] map instead of the uniform step
    if dimensions == 2:
        all_orders = np.concatenate((orders, -orders[:0:-1]))
        fpre = (-1.j) ** (all_orders[None, None, :]) * (r_max / div) ** 2
        ipre = (1.j) ** (all_orders[None, None, :]) * (q_max / div) ** 2
        weights = np.concatenate((weights, (-1.0) ** orders[:0:-1, None, None] * weights[:0:-1]), axis=0)
    elif dimensions == 3:
        fpre = (-1.j) ** (orders[None, None, :]) * (r_max / div) ** 3 * np.sqrt(2 / np.pi)
        ipre = (1.j) ** (orders[None, None, :]) * (q_max / div) ** 3 * np.sqrt(2 / np.pi)
    else:
        raise XfbError(f"dimensions={dimensions} not supported")
    w = np.moveaxis(weights, 0, 2)
    return {'forward': w * fpre, 'inverse': w * ipre, 'mode': mode}


class _HankelPair:
    """zht / izht closures sharing one device plan (built on first use)."""

    def __init__(self, weights, used_orders, r_max, rc, dimensions, mode, device=None, max_batch=1, plan=None):
        self.weights = np.ascontiguousarray(np.asarray(weights).real, dtype=np.float64)
        self.orders = np.asarray(used_orders)
        self.l_max = int(self.orders.max())
        if self.weights.shape[0] < self.l_max + 1:
            raise ValueError("weights do not cover the used orders")
        if not np.array_equal(self.orders, np.arange(self.l_max + 1)):
            raise XfbError("xframe_b200 Hankel transform needs used_orders = arange(l_max+1)")
        self.n_r = self.weights.shape[-1]
        self.r_max, self.rc, self.dim, self.mode = float(r_max), float(rc), int(dimensions), mode
        self.device, self.max_batch, self._plan = device, int(max_batch), plan

    def scales(self):
        if self.dim == 3:
            return tables.hankel_scales(self.r_max, self.n_r, self.rc, self.mode)
        return tables.polar_hankel_scales(self.r_max, self.n_r, self.rc, self.mode)

    def plan(self):
        if self._plan is None:
            if self.dim == 3:
                from .plan import Plan
                self._plan = Plan(self.l_max, self.n_r, _q_max_from_r_max(self.r_max, self.n_r, self.rc), reciprocity_coefficient=self.rc,
                                  ft_type=self.mode, max_batch=self.max_batch, device=self.device,
                                  hankel_weights=self.weights[:self.l_max + 1], hankel_scales=self.scales())
            else:
                from .plan import Plan
                self._plan = Plan(self.l_max, self.n_r, _q_max_from_r_max(self.r_max, self.n_r, self.rc), reciprocity_coefficient=self.rc,
                                  ft_type=self.mode, max_batch=self.max_batch, device=self.device, dimensions=2,
                                  hankel_weights=self.weights[:self.l_max + 1], hankel_scales=self.scales())
        return self._plan

    def apply(self, coeff, inverse):
        import torch
        plan = self.plan()
        was_torch = isinstance(coeff, torch.Tensor)
        c = coeff if was_torch else torch.from_numpy(np.ascontiguousarray(coeff, dtype=np.complex128))
        c = c.to(device=plan.device, dtype=torch.complex128).contiguous()
        squeeze = c.dim() == 2
        if squeeze:
            c = c[None]
        nb = c.shape[0]
        outs = [plan.hankel(c[i:i + plan.max_batch].contiguous(), inverse=inverse) for i in range(0, nb, plan.max_batch)]
        out = outs[0] if len(outs) == 1 else torch.cat(outs)
        if squeeze:
            out = out[0]
        return out if was_torch else out.cpu().numpy()


def generate_ht(weights, used_orders, r_max, reciprocity_coefficient=np.pi, dimensions=3, use_gpu=True, mode=ht_modes[0], device=None,
                max_batch=1, plan=None):
    """(zht, izht) on the 'direct' coefficient layout [N_r, (L+1)^2] (3-D) or [N_r, 2M+1] (2-D), like the reference's
    GPU flavours (hankel_transforms.py:540-559,660-870).  `use_gpu=False` is refused: there is no CPU path here."""
    if not use_gpu:
        raise XfbError("xframe_b200.generate_ht: use_gpu=False requested, but this package has no CPU path")
    if mode not in ('trapz', 'midpoint', 'gauss'):
        raise XfbError(f"Hankel transform mode '{mode}' is not implemented by xframe_b200 (trapz, midpoint, gauss)")
    pair = _HankelPair(weights, used_orders, r_max, reciprocity_coefficient, dimensions, mode, device, max_batch, plan)

    def zht(harmonic_coeff):
        return pair.apply(harmonic_coeff, False)

    def izht(reciprocal_coeff):
        return pair.apply(reciprocal_coeff, True)

    zht.pair = izht.pair = pair
    return zht, izht
