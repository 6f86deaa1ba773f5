"""Noise schedules with the reference's class names and methods (models/components/noise_schedules.py).
Host-side scalar algebra only; the sampling loop folds h(t), g(t), dh/dt into kernel arguments."""
from __future__ import annotations

import math

import torch


def _pow(v, p):
    return v ** p


class BaseNoiseSchedule:
    def g(self, t):
        raise NotImplementedError

    def h(self, t):
        raise NotImplementedError


class ElucidatingNoiseSchedule(BaseNoiseSchedule):
    """h(t) = (smax^(1/rho) + (1-t)(smin^(1/rho) - smax^(1/rho)))^(2 rho); g = sqrt(dh/dt)  (reference :98-125)."""

    def __init__(self, sigma_min, sigma_max, rho, P_mean=-1.2, P_std=1.2):
        self.sigma_min, self.sigma_max, self.rho = sigma_min, sigma_max, rho
        self.term1 = sigma_max ** (1 / rho)
        self.term2 = sigma_min ** (1 / rho) - sigma_max ** (1 / rho)
        self.P_mean, self.P_std = P_mean, P_std

    def _base(self, t):
        return self.term1 + (1 - t) * self.term2

    def h(self, t):
        return _pow(self._base(t), 2 * self.rho)

    def dh_dt(self, t):
        return -2 * self.rho * self.term2 * _pow(self._base(t), 2 * self.rho - 1)

    def g(self, t):
        return _pow(-2 * self.rho * _pow(self._base(t), 2 * self.rho - 1) * self.term2, 0.5)

    def t(self, ht):
        return 1 - ((_pow(ht, 1 / (2 * self.rho)) - self.term1) / self.term2)

    def sample_ln_sigma(self, num_samples, device):
        return torch.randn(num_samples, device=device) * self.P_std + self.P_mean


class GeometricNoiseSchedule(BaseNoiseSchedule):
    """reference :63-95."""

    def __init__(self, sigma_min, sigma_max):
        self.sigma_min, self.sigma_max = sigma_min, sigma_max
        self.sigma_diff = sigma_max / sigma_min

    def g(self, t):
        return self.sigma_min * (self.sigma_diff ** t) * ((2 * math.log(self.sigma_diff)) ** 0.5)

    def h(self, t):
        return (self.sigma_min * (((self.sigma_diff ** (2 * t)) - 1) ** 0.5)) ** 2

    def dh_dt(self, t):
        return self.g(t) ** 2


class LinearNoiseSchedule(BaseNoiseSchedule):
    """reference :19-27."""

    def __init__(self, beta):
        self.beta = beta

    def g(self, t):
        return torch.full_like(t, self.beta ** 0.5) if torch.is_tensor(t) else self.beta ** 0.5

    def h(self, t):
        return self.beta * t

    def dh_dt(self, t):
        return self.beta + 0 * t
