"""Host scalars of the phasing schedule: beta(step), sigma(sw_step), threshold(sw_step).

Same behaviour as ExponentialRamp / LinearRamp of the reference (library/mathLibrary.py:1033-1129),
including its handling of `False` placeholders in the settings lists.
"""
import math

import numpy as np


def _is_number(v):
    return np.issubdtype(np.array(v).dtype, np.number)


class ExponentialRamp:
    """value(x) = A exp(kappa x) + B, clipped at `stop`; kappa's sign follows the ramp direction."""

    def __init__(self, start, stop, exponent, stop_argument=1):
        self.start, self.stop = start, stop
        kappa = -abs(exponent) if stop < start else abs(exponent)
        self.kappa = kappa
        self.A = (start - stop) / (1 - math.exp(kappa * stop_argument))
        self.B = start - self.A

    def eval(self, x):
        v = self.A * math.exp(x * self.kappa) + self.B
        return max(v, self.stop) if self.start > self.stop else min(v, self.stop)

    __call__ = eval


class LinearRamp:
    def __init__(self, start, stop=False, slope=False, default_start=False, default_stop=False):
        self.start = tuple(start) if isinstance(start, (list, tuple)) else (start, 0)
        self.undefined = False
        if not _is_number(self.start[0]):
            if default_start == False:  # noqa: E712 -- the reference's test (0 counts as unset)
                self.undefined = True
            else:
                self.start = (default_start, 0)
        self.stop = None
        if isinstance(stop, (list, tuple)):
            sv, sa = stop[0], stop[1]
            if not _is_number(sv) and _is_number(default_stop):
                sv = default_stop
            if _is_number(sv) and _is_number(sa) and sa >= self.start[1]:
                self.stop = (sv, sa)
        self.slope = None if isinstance(slope, bool) else slope
        self.A, self.B, self.C = 0.0, 0.0, None
        if not self.undefined:
            s0, a0 = self.start
            if self.stop is None and self.slope is None:
                self.A, self.B = 0.0, s0
                return
            if self.stop is not None:
                self.C = self.stop[0]
                self.A = 0.0 if (self.stop[1] - a0) == 0 else (self.stop[0] - s0) / (self.stop[1] - a0)
                if self.slope is not None:
                    self.A = self.slope
            elif self.slope == 0:
                self.C, self.A = float('nan'), 0.0
            else:
                self.C, self.A = math.copysign(float('inf'), self.slope), self.slope
            self.B = s0 - self.A * a0

    def eval(self, x):
        if self.undefined:
            return float('nan')
        v = self.A * x + self.B
        if self.A < 0:
            v = max(v, self.C)
        elif self.A > 0:
            v = min(v, self.C)
        return v

    __call__ = eval
