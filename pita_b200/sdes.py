"""`SDETerms` / `VEReverseSDE` with the reference's interface (models/components/sdes.py).

`f` keeps the reference's signature and return type for callers that want the individual terms; the
sampling loop (sde_integration.py) does not go through it but launches the same kernels with the drift
assembly fused into the Euler-Maruyama step.  Network terms come from two kernel launches:
`pita_egnn_energy` (U_t, grad U_t, dU_t/dh) and `pita_egnn_score_div` (s_t, div s_t) — no autograd, no
vmap(jacrev)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import ops


@dataclass
class SDETerms:
    drift_X: torch.Tensor
    drift_A: torch.Tensor
    divergence_score: Optional[torch.Tensor] = None
    cross_term: Optional[torch.Tensor] = None
    dUt_dt: Optional[torch.Tensor] = None
    diffusion: Optional[torch.Tensor] = None

    @staticmethod
    def cpu(data):
        mv = lambda v: None if v is None else v.cpu()  # noqa: E731
        return SDETerms(mv(data.drift_X), mv(data.drift_A), mv(data.divergence_score), mv(data.cross_term),
                        mv(data.dUt_dt), mv(data.diffusion))

    @staticmethod
    def concatenate(data_list):
        if not data_list:
            raise ValueError("The data_list is empty.")
        cat = lambda name: (None if getattr(data_list[0], name) is None  # noqa: E731
                            else torch.cat([getattr(d, name) for d in data_list], dim=0))
        return SDETerms(cat("drift_X"), cat("drift_A"), cat("divergence_score"), cat("cross_term"), cat("dUt_dt"),
                        cat("diffusion"))


def _per_particle(t, B, device):
    if not torch.is_tensor(t):
        t = torch.tensor(float(t))
    t = t.detach().to(device=device, dtype=torch.float32)
    return t.reshape(1).expand(B).contiguous() if t.dim() == 0 else t


class VEReverseSDE:
    def __init__(self, noise_schedule, energy_net=None, score_net=None, cdf=None, pin_energy=False,
                 debias_inference=True):
        self.energy_net = energy_net
        self.score_net = score_net
        self.noise_schedule = noise_schedule
        self.pin_energy = pin_energy
        self.debias_inference = debias_inference
        # `cdf` (a compiled vmap(jacrev) divergence in the reference, sdes.py:111-113) is accepted for
        # signature compatibility and ignored: the divergence comes from the score kernel itself.
        self.compiled_divergence_fn = cdf
        self.is_compiled = True
        self.trainer = None

    def g(self, t):
        return self.noise_schedule.g(t)

    def dh_dt(self, t):
        ns = self.noise_schedule
        return ns.dh_dt(t) if hasattr(ns, "dh_dt") else ns.g(t) ** 2

    def f_not_debiased(self, t, x, beta, gamma_energy):
        assert self.score_net is not None
        ht = self.noise_schedule.h(t)
        s = self.score_net(ht, x, beta)
        drift_X = gamma_energy.reshape(-1, 1) * (s * self.g(t).pow(2).unsqueeze(-1))  # sdes.py:120-122
        return SDETerms(drift_X=drift_X, drift_A=torch.zeros(x.shape[0], device=x.device))

    def f(self, t, x, beta, gamma_energy_schedule, gamma_score, energy_function, resampling_interval=-1):
        B, dev = x.shape[0], x.device
        t = _per_particle(t, B, dev)
        gamma = gamma_energy_schedule.gamma(t).to(dev)
        if not self.debias_inference:
            return self.f_not_debiased(t, x, beta, gamma)
        assert self.energy_net is not None
        ht = self.noise_schedule.h(t)
        g2 = self.g(t).pow(2)
        U, nabla_U, dU_dt = self.energy_net.energy_grad_dh(ht, x, beta, pin=self.pin_energy, energy_function=energy_function,
                                                           t=t, dh_dt=self.dh_dt(t))
        if self.score_net is not None:
            s_t, div_s = self.score_net.score_and_divergence(ht, x, beta)
            bt = s_t * g2[:, None] / 2
            div_bt = div_s * g2 / 2
        else:  # no score net (sdes.py:169-170, 204-216): b = -grad U g^2/2, div b = -laplacian(U) g^2/2
            bt = -nabla_U * g2[:, None] / 2
            div_bt = -self.energy_net.laplacian(ht, x, beta, pin=self.pin_energy, t=t) * g2 / 2
        drift_X = gamma[:, None] * -nabla_U * g2[:, None] / 2 + gamma[:, None] * bt  # sdes.py:172-174 (gamma_score := gamma, :143)
        inner = (-nabla_U * bt).sum(-1)
        raw = gamma * gamma * inner + gamma * div_bt + gamma * dU_dt + gamma_energy_schedule.dgamma_dt(t).to(dev) * U
        # clamp at this call's own 0.9-quantile (sdes.py:230) — one chunk == one call
        drift_A, _ = ops.fk_quantile_accumulate(raw, None, B, 0.9, 0.0, False)
        return SDETerms(drift_X=drift_X, drift_A=drift_A, divergence_score=div_bt, cross_term=inner, dUt_dt=dU_dt)

    def diffusion(self, t, x, diffusion_scale):
        t = _per_particle(t, x.shape[0], x.device)
        return diffusion_scale * self.g(t)[:, None] * torch.randn_like(x)
