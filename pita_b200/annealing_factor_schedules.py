"""Annealing-factor schedules gamma(t) with the reference's names (models/components/annealing_factor_schedules.py)."""
from __future__ import annotations

import torch


def _t(t):
    return t if isinstance(t, torch.Tensor) else torch.tensor(t)


class BaseAnnealingFactorSchedule:
    def gamma(self, t):
        raise NotImplementedError

    def dgamma_dt(self, t):
        raise NotImplementedError


class ConstantAnnealingFactorSchedule(BaseAnnealingFactorSchedule):
    def __init__(self, annealing_factor):
        self.annealing_factor = annealing_factor

    def gamma(self, t):
        return torch.ones_like(_t(t)) * self.annealing_factor

    def dgamma_dt(self, t):
        return torch.zeros_like(_t(t))


class LinearAnnealingFactorSchedule(BaseAnnealingFactorSchedule):
    """Linear between (t_start, start) and (t_end, final), flat outside (reference :35-71)."""

    def __init__(self, annealing_factor, annealing_factor_start, t_start=1.0, t_end=0.0):
        self.annealing_factor, self.annealing_factor_start = annealing_factor, annealing_factor_start
        self.t_start, self.t_end = t_start, t_end

    @property
    def _slope(self):
        return (self.annealing_factor - self.annealing_factor_start) / (self.t_end - self.t_start)

    def gamma(self, t):
        t = _t(t)
        ramp = self._slope * (t - self.t_start) + self.annealing_factor_start
        inner = torch.where(t < self.t_end, torch.full_like(ramp, self.annealing_factor), ramp)
        return torch.where(t > self.t_start, torch.full_like(ramp, self.annealing_factor_start), inner)

    def dgamma_dt(self, t):
        t = _t(t)
        flat = (t > self.t_start) | (t < self.t_end)
        return torch.where(flat, torch.zeros_like(t), torch.full_like(t, self._slope))


class SigmoidAnnealingFactorSchedule(BaseAnnealingFactorSchedule):
    """Smooth sigmoid ramp (reference :74-109)."""

    def __init__(self, annealing_factor, annealing_factor_start, t_start=1.0, t_end=0.0, sharpness=10.0):
        self.annealing_factor, self.annealing_factor_start = annealing_factor, annealing_factor_start
        self.t_start, self.t_end, self.sharpness = t_start, t_end, sharpness
        self.center, self.width = (t_start + t_end) / 2, t_start - t_end

    def _smooth(self, t):
        return 1 / (1 + torch.exp(-self.sharpness * (self.center - t) / self.width))

    def gamma(self, t):
        s = self._smooth(_t(t))
        return self.annealing_factor_start + (self.annealing_factor - self.annealing_factor_start) * s

    def dgamma_dt(self, t):
        s = self._smooth(_t(t))
        return (self.annealing_factor - self.annealing_factor_start) * (self.sharpness / self.width) * s * (1 - s)
