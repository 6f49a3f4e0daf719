"""Circular-harmonic transforms of the 2-D fxs path, backed by the CUDA library.

Mirrors circularHarmonicTransform_{complex,real}_{forward,inverse} (xframe/library/mathLibrary.py:469-496) as selected
by HarmonicTransform for `dimensions: 2` (harmonic_transforms.py:36-60):

    complex forward  fft(x, axis=1) / n_phi             [N_r, n_phi] -> [N_r, n_phi]   orders (0..M, -M..-1)
    complex inverse  ifft(c * n_phi, axis=1)
    real    forward  rfft(Re x) / n_phi                 [N_r, n_phi] -> [N_r, M+1]
    real    inverse  irfft(c * size, size)              [N_r, M+1]   -> [N_r, size] (real)

numpy in / numpy out like the reference, or CUDA tensors (no host copies).  n_phi = 2M+1 is odd; the device routine is a
direct DFT (csrc/polar.cuh).  There is no CPU fallback.
"""
import numpy as np

from ._lib import XfbError


class CircularHarmonicTransform:
    _CAPACITY_SHELLS = 1 << 16

    def __init__(self, size, data_type='complex', device=None, plan=None):
        size = int(size)
        if size % 2 != 1 or size < 3:
            raise XfbError(f"circular harmonic transform needs an odd number of angular points 2M+1 (got {size})")
        if data_type not in ('complex', 'real'):
            raise AssertionError(f'harmonic transform mode "{data_type}" is not known. Known modes are "real" and "complex".')
        self.size, self.max_order, self.data_type, self._device = size, (size - 1) // 2, data_type, device
        self.phis = np.arange(size) / size * 2 * np.pi
        self._plan = plan
        self.forward = self._forward_complex if data_type == 'complex' else self._forward_real
        self.inverse = self._inverse_complex if data_type == 'complex' else self._inverse_real

    def _get_plan(self):
        if self._plan is None:
            from .plan import Plan
            self._plan = Plan(self.max_order, 8, 1.0, max_batch=self._CAPACITY_SHELLS // 8, device=self._device, dimensions=2)
        return self._plan

    def _dev(self, a, real_part=False):
        import torch
        plan = self._get_plan()
        was_torch = isinstance(a, torch.Tensor)
        t = a if was_torch else torch.from_numpy(np.asarray(a))
        t = t.to(plan.device)
        if real_part and t.is_complex():
            t = t.real
        return t.to(torch.complex128).contiguous(), was_torch

    def _chunks(self, fn, flat):
        import torch
        plan = self._get_plan()
        cap = plan.max_batch * plan.n_r
        if flat.shape[0] <= cap:
            return fn(flat)
        return torch.cat([fn(flat[i:i + cap].contiguous()) for i in range(0, flat.shape[0], cap)])

    def _run(self, t, inverse):
        plan = self._get_plan()
        if t.shape[-1] != self.size:
            raise ValueError(f"last axis {t.shape[-1]} != number of angular points {self.size}")
        lead = t.shape[:-1]
        flat = t.reshape(-1, self.size)
        out = self._chunks(plan.sht_inverse if inverse else plan.sht_forward, flat)
        return out.reshape(*lead, self.size)

    @staticmethod
    def _out(t, was_torch):
        return t if was_torch else t.cpu().numpy()

    def _forward_complex(self, data):
        t, tt = self._dev(data)
        return self._out(self._run(t, False), tt)

    def _inverse_complex(self, data):
        t, tt = self._dev(data)
        return self._out(self._run(t, True), tt)

    def _forward_real(self, data):
        t, tt = self._dev(data, real_part=True)                    # np.copy(data_array.real)  (mathLibrary.py:486)
        return self._out(self._run(t, False)[..., :self.max_order + 1].contiguous(), tt)

    def _inverse_real(self, data):
        import torch
        c, tt = self._dev(data)
        if c.shape[-1] != self.max_order + 1:
            raise ValueError(f"last axis {c.shape[-1]} != M+1 = {self.max_order + 1}")
        full = torch.cat((c, c[..., 1:].conj().flip(-1)), dim=-1).contiguous()      # irfft: Hermitian completion, Im c_0 ignored
        full[..., 0] = full[..., 0].real.to(torch.complex128)
        return self._out(self._run(full, True).real.contiguous(), tt)
