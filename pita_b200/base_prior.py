"""`Prior` / `MeanFreePrior` with the reference's interface (energies/base_prior.py)."""
from __future__ import annotations

import math

import torch

from . import ops


class MeanFreePrior(torch.distributions.Distribution):
    arg_constraints = {}

    def __init__(self, n_particles, spatial_dim, scale, device="cuda"):
        super().__init__(validate_args=False)
        self.n_particles, self.spatial_dim, self.dim = n_particles, spatial_dim, n_particles * spatial_dim
        self.scale, self.device = scale, device

    def log_prob(self, x):
        v = x.reshape(-1, self.n_particles, self.spatial_dim)
        dof = (self.n_particles - 1) * self.spatial_dim
        r2 = v.pow(2).sum(dim=(-1, -2)) / self.scale ** 2
        return -0.5 * r2 - 0.5 * dof * math.log(2 * torch.pi * self.scale ** 2)

    def sample(self, n_samples):
        if not isinstance(n_samples, int):
            n_samples = int(torch.Size(n_samples).numel())
        z = torch.randn(n_samples, self.dim, device=self.device) * self.scale  # reference :80
        return ops.remove_mean(z, self.n_particles)


class Prior:
    def __init__(self, scale, n_particles=None, spatial_dim=None, dim=None, device="cuda", should_mean_free=True):
        assert n_particles is not None and spatial_dim is not None
        self.dim, self.n_particles, self.spatial_dim, self.scale = n_particles * spatial_dim, n_particles, spatial_dim, scale
        if should_mean_free:
            self.dist = MeanFreePrior(n_particles, spatial_dim, scale, device)
        else:
            self.dist = torch.distributions.MultivariateNormal(torch.zeros(self.dim, device=device),
                                                               torch.eye(self.dim, device=device) * (scale ** 2))

    def log_prob(self, x):
        return self.dist.log_prob(x)

    def sample(self, n_samples):
        return self.dist.sample((n_samples,)) if not isinstance(self.dist, MeanFreePrior) else self.dist.sample(n_samples)
