"""Checks the hand-derived reverse / tangent algebra (oracle/egnn_analytic.py — the algebra the CUDA
kernels implement) against the autograd oracle, in fp64 on the CPU."""
import numpy as np
import pytest
import torch

import egnn_analytic as A
import pita_oracle as O


def _rel(a, b):
    a, b = a.detach().double().numpy(), b.detach().double().numpy()
    return float((np.abs(a - b) / np.maximum(np.abs(b), 1.0)).max())


@pytest.mark.parametrize("n,B", [(13, 4), (55, 2), (5, 3)])
def test_energy_reverse_pass(n, B):
    sd = O.random_egnn_state(seed=3 + n, dtype=torch.float64, coord_gain=0.3)
    sched = O.EDMSchedule(0.05)
    x = O.centre(O.md_shaped_coords(B, n, seed=n, dtype=torch.float64) * 1.3, n)
    t = torch.linspace(0.2, 0.9, B, dtype=torch.float64)
    beta = 0.8
    tr = t.clone().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    E = O.model_energy(sd, sched.h(tr), xr, beta, n)
    gx, gt = torch.autograd.grad(E.sum(), (xr, tr))
    E2, gE2, dE_dh = A.energy_terms(sd, sched.h(t), x, torch.tensor(beta, dtype=torch.float64), n)
    assert _rel(E2, E) < 1e-11
    assert _rel(gE2, gx) < 1e-10
    assert _rel(dE_dh * sched.dh_dt(t), gt) < 1e-10


@pytest.mark.parametrize("n,B", [(13, 3), (55, 1), (4, 3)])
def test_score_divergence_tangent_pass(n, B):
    sd = O.random_egnn_state(seed=5 + n, dtype=torch.float64, coord_gain=0.3)
    sched = O.EDMSchedule(0.05)
    x = O.centre(O.md_shaped_coords(B, n, seed=n + 1, dtype=torch.float64) * 1.2, n) + 0.05  # NOT mean free
    t = torch.linspace(0.3, 0.8, B, dtype=torch.float64)
    ht = sched.h(t)
    beta = 1.25
    s_ref = O.model_score(sd, ht, x, beta, n)
    div_ref = O.exact_divergence(lambda h1, x1: O.model_score(sd, h1, x1, beta, n), ht, x)
    s, div = A.score_and_divergence(sd, ht, x, torch.tensor(beta, dtype=torch.float64), n)
    assert _rel(s, s_ref) < 1e-11
    assert _rel(div, div_ref) < 1e-10
