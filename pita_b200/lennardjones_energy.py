"""Lennard-Jones target with the reference's call contract (energies/lennardjones_energy.py:158-227):
`energy(samples[B, 3n], return_force=False) -> logp[B]` or `(logp[B], force[B, 3n])`, logp = -E/T.
Energy and force come from ONE fused CUDA kernel (`pita_lj_energy_force`, analytic force — no autograd).
Dataset loading / plotting of the reference's base classes is out of scope."""
from __future__ import annotations

import torch

from . import ops


class LennardJonesPotential:
    """reference :57-155 (`_energy`, `_log_prob`); spline smoothing (:131-133) is not built."""

    def __init__(self, dim, n_particles, eps=1.0, rm=1.0, oscillator=True, oscillator_scale=1.0, two_event_dims=True,
                 energy_factor=1.0, range_min=0.65, range_max=2.0, interpolation=1000, temperature=1.0):
        if eps != 1.0 or rm != 1.0:
            raise NotImplementedError("pita_b200 LJ kernel is built for eps = rm = 1 (the reference's only use)")
        self._n_particles, self._n_dims = n_particles, dim // n_particles
        self._energy_factor = energy_factor
        self._osc = oscillator_scale if oscillator else 0.0
        self._temperature = temperature

    def _log_prob(self, x, smooth=False):
        if smooth:
            raise NotImplementedError("smooth=True (cubic-spline core) is not built")
        lp, _ = ops.lj_energy_force(x.reshape(-1, self._n_particles * self._n_dims), self._n_particles, self._temperature,
                                    self._energy_factor, self._osc, need_force=False)
        return lp[:, None]

    def _energy(self, x, smooth=False):
        return -self._log_prob(x, smooth) * self._temperature


class LennardJonesEnergy:
    def __init__(self, dimensionality, n_particles, spatial_dim=3, data_path=None, device="cuda",
                 plot_samples_epoch_period=5, plotting_buffer_sample_size=512, energy_factor=1.0, is_molecule=True,
                 smooth=False, temperature=1.0, should_normalize=False, data_normalization_factor=1.0, *args, **kwargs):
        if n_particles != 13 and n_particles != 55:
            raise NotImplementedError  # reference :177-178
        if spatial_dim != 3 or dimensionality != 3 * n_particles:
            raise NotImplementedError("3-D configurations only")
        if smooth:
            raise NotImplementedError("smooth=True (cubic-spline core) is not built")
        self.name = "LJ13_efm" if n_particles == 13 else "LJ55"
        self.n_particles, self.n_spatial_dim, self.dimensionality = n_particles, spatial_dim, dimensionality
        self.is_molecule, self.temperature, self.device = is_molecule, temperature, device
        self.energy_factor = energy_factor
        self.should_normalize, self.data_normalization_factor = should_normalize, data_normalization_factor
        self.smooth = smooth

    def unnormalize(self, x):
        return x * self.data_normalization_factor

    def __call__(self, samples: torch.Tensor, return_force=False):
        if self.should_normalize:
            samples = self.unnormalize(samples)
        logp, force = ops.lj_energy_force(samples, self.n_particles, self.temperature, self.energy_factor, 1.0,
                                          need_force=return_force)
        if return_force:
            # reference :214-223: `samples` is rebound to the UNNORMALISED coordinates before autograd.grad, so the force is the
            # gradient w.r.t. those — no data_normalization_factor applied
            return logp, force
        return logp
