"""the no-score-net branch of VEReverseSDE.f (b = -grad U g^2/2, div b = -laplacian(U) g^2/2 through
compute_laplacian_exact, reference sdes.py:150-153, 204-216; utils.py:68-77) on the GPU: csrc/egnn_lap.cu through the C-ABI
against the UNMODIFIED reference's outputs (tests/golden/fk_n13_laplacian.npz) and the fp64 oracle."""
import numpy as np
import pytest
import torch

import pita_oracle as O
from helpers import assert_close, golden, make_net, state_from_golden

pytestmark = pytest.mark.gpu


def _sde(n, sdE, pin=False):
    from pita_b200.energy_net import EnergyNet
    from pita_b200.noise_schedules import ElucidatingNoiseSchedule
    from pita_b200.sdes import VEReverseSDE
    return VEReverseSDE(ElucidatingNoiseSchedule(0.05, 80.0, 7.0), energy_net=EnergyNet(make_net(n, sdE)), score_net=None,
                        pin_energy=pin, debias_inference=True)


def test_laplacian_branch_vs_reference_golden():
    from pita_b200.annealing_factor_schedules import ConstantAnnealingFactorSchedule
    g = golden("fk_n13_laplacian.npz")
    n = int(g["n"])
    sde = _sde(n, state_from_golden(g, "E."))
    x = torch.from_numpy(g["x"]).float().cuda()
    terms = sde.f(torch.tensor(float(g["t"]), device="cuda"), x, torch.tensor(float(g["beta"]), device="cuda"),
                  ConstantAnnealingFactorSchedule(float(g["gamma"])), 1.0, None, resampling_interval=1)
    assert_close(terms.divergence_score, g["div_b"], "div_b = -laplacian(U) g^2/2")
    assert_close(terms.cross_term, g["cross"], "cross term")
    assert_close(terms.drift_X, g["drift_X"], "drift_X")
    assert_close(terms.drift_A, g["drift_A"], "drift_A", rtol=2e-4)


@pytest.mark.parametrize("n,B", [(13, 37), (55, 3)])
def test_energy_laplacian_vs_oracle(n, B):
    """tr(Hess_x E) against vmap(hessian) of the fp64 oracle, per-particle h(t) and beta, ragged batch."""
    from pita_b200 import ops
    sd = O.random_egnn_state(seed=90 + n, dtype=torch.float64, coord_gain=0.3)
    gen = torch.Generator().manual_seed(n)
    x = O.centre(O.md_shaped_coords(B, n, seed=3 + n, dtype=torch.float64) * (1 + 0.05 * torch.randn(B, 1, generator=gen, dtype=torch.float64)), n)
    ht = O.EDMSchedule(0.05).h(torch.linspace(0.25, 0.7, B, dtype=torch.float64))
    beta = torch.linspace(0.6, 1.3, B, dtype=torch.float64)
    ref = []
    for b in range(B):
        ref.append(O.exact_laplacian(lambda h1, x1, _b=b: O.model_energy(sd, h1, x1, beta[_b:_b + 1], n), ht[b:b + 1], x[b:b + 1]))
    ref = torch.cat(ref)
    net = make_net(n, sd)
    lap = ops.egnn_energy_laplacian(net.packed_weights("cuda"), 32, 3, n, ht.float().cuda(), x.float().cuda(), beta.float().cuda())
    assert_close(lap, ref, "laplacian of the model energy (n=%d)" % n)
    assert float(ref.abs().max()) > 1.0
    lap2 = ops.egnn_energy_laplacian(net.packed_weights("cuda"), 32, 3, n, ht.float().cuda(), x.float().cuda(), beta.float().cuda())
    assert torch.equal(lap, lap2)  # deterministic


def test_loop_without_score_net_vs_oracle():
    """integrate_sde with score_net=None (fused step fed -grad U / -laplacian U) against the oracle's Laplacian branch."""
    from pita_b200.annealing_factor_schedules import ConstantAnnealingFactorSchedule
    from pita_b200.lennardjones_energy import LennardJonesEnergy
    from pita_b200.sde_integration import WeightedSDEIntegrator
    n, N, S, chunk, time_range, gam, beta = 13, 24, 4, 12, 0.15, 4.0 / 3.0, 0.9
    sdE = O.random_egnn_state(seed=61, dtype=torch.float64, coord_gain=0.3)
    sched = O.EDMSchedule(0.05)
    gen = torch.Generator().manual_seed(N)
    scale = float((sched.h(torch.tensor(time_range, dtype=torch.float64)) / gam) ** 0.5)
    x1 = O.centre(O.md_shaped_coords(N, n, seed=N, dtype=torch.float64) + scale * torch.randn(N, 3 * n, generator=gen, dtype=torch.float64), n)
    noise = {s: torch.randn(N, 3 * n, generator=gen, dtype=torch.float64) for s in range(S)}
    u0 = {s: float(torch.rand(1, generator=gen, dtype=torch.float64)) for s in range(S + 1)}
    cfg = O.LoopConfig(n=n, steps=S, chunk=chunk, beta=beta, resampling_interval=1, time_range=time_range)
    cursor = {}

    def noise_fn(step, xc):
        lo = cursor.get(step, 0)
        cursor[step] = lo + xc.shape[0]
        return noise[step][lo:lo + xc.shape[0]]

    x_ref, logw_ref, uniq_ref = O.integrate(sdE, None, sched, O.ConstGamma(gam), cfg, x1, noise_fn, lambda s: u0[s])
    integ = WeightedSDEIntegrator(sde=_sde(n, sdE), num_integration_steps=S, lightning_module=None, batch_size=chunk,
                                  num_negative_time_steps=0, post_mcmc_steps=0, start_resampling_step=0, end_resampling_step=10 ** 9,
                                  resampling_interval=1, time_range=time_range)
    integ.noise_fn = lambda step, x: noise[step].float().cuda()
    integ.u0_fn = lambda step: u0[step]
    x, logw, uniq, _, _ = integ.integrate_sde(x1.float().cuda(), LennardJonesEnergy(dimensionality=3 * n, n_particles=n),
                                              ConstantAnnealingFactorSchedule(gam), inverse_temperature=beta)
    assert list(uniq) == list(uniq_ref)
    assert_close(x, x_ref, "x_final (no score net)", rtol=1e-3)
