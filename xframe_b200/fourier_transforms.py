"""Fourier-transform composer of the fxs path: ft = SHT -> Hankel -> SHT^-1 (fourier_transforms.py:17-86).

`generate_ft` keeps the reference signature.  All three stages run on the device through ONE plan
(`xfb_ft` of include/xfb200.h); the harmonic-transform object and the Hankel pair handed back share that plan.
"""
import numpy as np

from ._lib import XfbError
from .hankel_transforms import generate_ht, generate_weightDict


def load_fourier_transform_weights(dimensions, ft_opt, grid_opt, database=None):
    """fourier_transforms.py:17-39 without the on-disk cache (weights take milliseconds to build here)."""
    get = (lambda o, k: o[k] if isinstance(o, dict) else getattr(o, k))
    rc = get(ft_opt, 'reciprocity_coefficient')
    return generate_weightDict(get(grid_opt, 'max_order'), get(grid_opt, 'n_radial_points'), reciprocity_coefficient=rc,
                               dimensions=dimensions, mode=get(ft_opt, 'type'))


def select_harmonic_transforms(harm_trf, dimensions, use_gpu=True):          # fourier_transforms.py:43-51
    if dimensions == 3:
        trfs = harm_trf.transforms_by_indices['direct' if use_gpu else 'ml']
    elif dimensions == 2:
        trfs = harm_trf.transforms_by_indices['m']
    else:
        raise XfbError(f"dimensions={dimensions} not supported")
    return trfs['forward'], trfs['inverse']


def generate_ft(r_max, weights, harm_trf, dimensions, pos_orders=False, use_gpu=True, reciprocity_coefficient=np.pi, mode='trapz',
                max_batch=1):
    """(ft, ift): [N_r, n_theta, n_phi] -> same (3-D) or [N_r, n_phi] -> same (2-D); numpy or CUDA tensors, optionally with
    a leading batch axis.  fourier_transforms.py:53-85."""
    if not use_gpu:
        raise XfbError("xframe_b200.generate_ft: use_gpu=False requested, but this package has no CPU path")
    orders = weights['posHarmOrders'] if isinstance(pos_orders, bool) else pos_orders
    hankel, ihankel = generate_ht(weights['weights'], orders, r_max, reciprocity_coefficient=reciprocity_coefficient,
                                  dimensions=dimensions, use_gpu=True, mode=mode, device=harm_trf.opt.get('device'), max_batch=max_batch)
    ht, iht = select_harmonic_transforms(harm_trf, dimensions, True)
    pair = hankel.pair

    if dimensions not in (2, 3):
        raise XfbError(f"dimensions={dimensions} not supported")
    ang = harm_trf._sh if dimensions == 3 else harm_trf._ch
    n_r = pair.n_r

    def fused(data, inverse):
        import torch
        if pair._plan is None:
            from .plan import Plan
            kw = dict(n_theta=ang.n_theta, n_phi=ang.n_phi) if dimensions == 3 else {}
            pair._plan = Plan(pair.l_max, n_r, reciprocity_coefficient * n_r / r_max, reciprocity_coefficient=reciprocity_coefficient,
                              ft_type=mode, max_batch=max_batch, device=harm_trf.opt.get('device'), dimensions=dimensions,
                              hankel_weights=pair.weights[:pair.l_max + 1], hankel_scales=pair.scales(), **kw)
            if dimensions == 3 and ang._plan is None:
                ang.attach_plan(pair._plan)
        plan = pair._plan
        was_torch = isinstance(data, torch.Tensor)
        d = data if was_torch else torch.from_numpy(np.ascontiguousarray(data, dtype=np.complex128))
        d = d.to(device=plan.device, dtype=torch.complex128).contiguous()
        squeeze = d.dim() == dimensions
        if squeeze:
            d = d[None]
        outs = [plan.ft(d[i:i + plan.max_batch].contiguous(), inverse=inverse) for i in range(0, d.shape[0], plan.max_batch)]
        out = outs[0] if len(outs) == 1 else torch.cat(outs)
        if squeeze:
            out = out[0]
        return out if was_torch else out.cpu().numpy()

    def ft(data):
        return fused(data, False)

    def ift(data):
        return fused(data, True)
    ft.hankel, ift.hankel = hankel, ihankel
    return ft, ift
