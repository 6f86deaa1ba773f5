"""`EnergyNet` with the reference's interface (models/components/energy_net.py).

One CUDA kernel (`pita_egnn_energy`) yields E_theta, grad_x E_theta and dE_theta/dh from a primal forward
plus a hand-derived reverse pass; nothing here uses autograd.  `pin=True` mixes in the target energy exactly as
the reference does (:41-48).  NB the reference's target returns `logprobs.detach()` (lennardjones_energy.py:227), so
the pinned energy's x-gradient carries NO target force — only the (1 - (1-t)^3) share of grad E_theta; pinned to the
reference by tests/golden/fk_n13_pin.npz."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import ops


class EnergyNet(nn.Module):
    def __init__(self, score_net: nn.Module, precondition_beta: Optional[bool] = False):
        super().__init__()
        self.net = score_net
        self.precondition_beta = precondition_beta

    def _terms(self, ht, xt, beta, need_grad, need_dh):
        net = self.net
        B = xt.shape[0]
        e, g, dh = ops.egnn_energy(net.packed_weights(xt.device), net.hidden_nf, net.n_layers, net._n_particles, ht, xt, beta,
                                   need_grad=need_grad, need_dh=need_dh)
        if self.precondition_beta:  # :38-39
            b = ops._expand(beta, B, xt.device)
            e = e * b
            g = g * b[:, None] if g is not None else None
            dh = dh * b if dh is not None else None
        return e, g, dh

    def energy_grad_dh(self, ht, xt, beta, pin=False, energy_function=None, t=None, dh_dt=None):
        """(E, grad_x E, dE/dt) in one pass.  dE/dt = dE/dh * dh_dt (+ the pin mixing terms)."""
        e, g, dh = self._terms(ht, xt, beta, True, True)
        de_dt = dh * dh_dt if dh_dt is not None else dh
        if pin:
            assert t is not None and energy_function is not None
            u0 = torch.clamp(-energy_function(xt), max=1e3, min=-1e3)  # detached in the reference: no d/dx through the target
            w = (1 - t) ** 3
            de_dt = -3 * (1 - t) ** 2 * u0 + 3 * (1 - t) ** 2 * e + (1 - w) * de_dt
            g = (1 - w)[:, None] * g
            e = w * u0 + (1 - w) * e
        return e, g, de_dt

    def laplacian(self, ht, xt, beta, pin=False, t=None):
        """tr(Hess_x forward_energy) — what the reference obtains from compute_laplacian_exact(partial(forward_energy, ...))
        (sdes.py:204-216).  The pinned target is detached in the reference, so only the (1 - (1-t)^3) E part contributes."""
        net = self.net
        lap = ops.egnn_energy_laplacian(net.packed_weights(xt.device), net.hidden_nf, net.n_layers, net._n_particles, ht, xt, beta)
        if self.precondition_beta:
            lap = lap * ops._expand(beta, xt.shape[0], xt.device)
        if pin:
            assert t is not None
            lap = (1 - (1 - t) ** 3) * lap
        return lap

    def forward_energy(self, ht, xt, beta, pin=False, energy_function=None, t=None):
        e, _, _ = self._terms(ht, xt, beta, False, False)
        if pin:
            assert t is not None and energy_function is not None
            u0 = torch.clamp(-energy_function(xt), max=1e3, min=-1e3)
            return (1 - t) ** 3 * u0 + (1 - (1 - t) ** 3) * e
        return e

    def forward(self, ht, xt, beta, pin=False, t=None, energy_function=None):
        if pin:
            return self.energy_grad_dh(ht, xt, beta, pin=True, energy_function=energy_function, t=t)[1]
        return self._terms(ht, xt, beta, True, False)[1]

    def denoiser(self, h_t, x_t, beta):
        return x_t - ops._expand(h_t, x_t.shape[0], x_t.device)[:, None] * self.forward(h_t, x_t, beta)

    def denoiser_and_energy(self, ht, xt, beta):
        """(:68-79) returns (x - h grad U, dU/dh, U)."""
        e, g, dh = self._terms(ht, xt, beta, True, True)
        return xt - ops._expand(ht, xt.shape[0], xt.device)[:, None] * g, dh, e

    def reinitialize(self, score_net: nn.Module):
        self.net = score_net
