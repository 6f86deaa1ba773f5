"""Post-processing on the target energy: negative-time descent, MALA proposal and accept/reject
(reference models/components/sde_integration.py:28-45, 353-470) vs (a) the UNMODIFIED reference's outputs on recorded draws
(tests/golden/post_n13.npz, oracle/make_golden.py::golden_post) and (b) a float64 restatement on larger batches."""
import numpy as np
import pytest
import torch

import pita_oracle as O
from helpers import assert_close, golden

pytestmark = pytest.mark.gpu


def _coords(B, n, seed):
    return O.centre(O.md_shaped_coords(B, n, seed=seed, dtype=torch.float64), n)


@pytest.mark.parametrize("n,B", [(13, 777), (55, 130)])
@pytest.mark.parametrize("langevin", [False, True])
def test_descent_step_vs_formula(n, B, langevin):
    """x <- remove_mean(x + force*dt [+ xi*sqrt(2 dt)])   (reference :353-360)"""
    from pita_b200 import ops
    x = _coords(B, n, 3)
    _, force = O.lj_logprob_force(x, n)
    gen = torch.Generator().manual_seed(5)
    xi = torch.randn(B, 3 * n, generator=gen, dtype=torch.float64)
    dt = 1e-4
    ref = x + force * dt
    if langevin:
        ref = ref + xi * np.sqrt(2 * dt)
    c = lambda v: v.float().cuda()  # noqa: E731
    got = ops.descent_step(c(x), c(force), c(xi) if langevin else None, n, dt, True)
    assert_close(got, O.centre(ref, n), "descent (mean-free)", rtol=1e-5)
    got = ops.descent_step(c(x), c(force), c(xi) if langevin else None, n, dt, False)
    assert_close(got, ref, "descent", rtol=1e-5)


@pytest.mark.parametrize("n,B", [(13, 1000), (55, 97)])
@pytest.mark.parametrize("mean_free", [False, True])
def test_mala_step_vs_formula(n, B, mean_free):
    """One MALA step (proposal :28-45, accept/reject :372-397) on injected noise and uniforms."""
    from pita_b200 import ops
    dt = 2e-3
    x = _coords(B, n, 11)
    logp, force = O.lj_logprob_force(x, n)
    gen = torch.Generator().manual_seed(6)
    xi = torch.randn(B, 3 * n, generator=gen, dtype=torch.float64)
    u = torch.rand(B, generator=gen, dtype=torch.float64).clamp_min(1e-12)
    c = lambda v: v.float().cuda()  # noqa: E731

    fwd_mean = x + 0.5 * dt * force
    x_prop = fwd_mean + np.sqrt(np.float32(dt)) * xi
    log_q_fwd = -((x_prop - fwd_mean) ** 2).sum(1) / (2 * dt)
    xp, lqf = ops.mala_propose(c(x), c(force), c(xi), n, dt)
    assert_close(xp, x_prop, "x_prop", rtol=1e-5)
    assert_close(lqf, log_q_fwd, "log q fwd", rtol=1e-4)

    logp_prop, force_prop = O.lj_logprob_force(x_prop, n)
    log_q_bwd = -((x - (x_prop + 0.5 * dt * force_prop)) ** 2).sum(1) / (2 * dt)
    ratio = (logp_prop - logp) + (log_q_bwd - log_q_fwd)
    margin = (torch.log(u) - ratio).abs()
    acc_ref = torch.log(u) < ratio
    x_ref = torch.where(acc_ref[:, None], x_prop, x)
    lp_ref = torch.where(acc_ref, logp_prop, logp)
    if mean_free:
        x_ref = O.centre(x_ref, n)

    xc, lpc = c(x).clone(), c(logp).clone()
    acc = ops.mala_accept(xc, lpc, c(x_prop), c(logp_prop), c(force_prop), c(log_q_fwd), c(u), n, dt, mean_free)
    sure = margin > 1e-3 * ratio.abs().clamp_min(1.0)  # decisions not within fp32 rounding of the threshold
    assert sure.float().mean() > 0.95
    assert torch.equal(acc.cpu().bool()[sure], acc_ref[sure])
    assert 0.0 < acc.mean().item() < 1.0, "test must exercise both branches"
    same = acc.cpu().bool() == acc_ref
    assert_close(xc.cpu()[same], x_ref[same], "x after accept", rtol=1e-5)
    assert_close(lpc.cpu()[same], lp_ref[same], "logp after accept", rtol=1e-5)


@pytest.mark.parametrize("adaptive", [False, True])
def test_integrator_post_processing(adaptive):
    """negative_time_descent raises log p; MALA keeps non-finite rows apart (valid rows first, :400) and reports rates."""
    from pita_b200.lennardjones_energy import LennardJonesEnergy
    from pita_b200.sde_integration import WeightedSDEIntegrator
    n, B = 13, 512
    tgt = LennardJonesEnergy(dimensionality=3 * n, n_particles=n)
    integ = WeightedSDEIntegrator(sde=None, num_integration_steps=1, start_resampling_step=0, end_resampling_step=1,
                                  num_negative_time_steps=20, post_mcmc_steps=8, adaptive_mcmc=adaptive, dt_negative_time=1e-4)
    x = _coords(B, n, 21).float().cuda()
    lp0 = tgt(x)
    xd = integ.negative_time_descent(x, tgt)
    assert (tgt(xd) >= lp0 - 1e-3).all() and tgt(xd).mean() > lp0.mean()
    assert xd.reshape(B, n, 3).mean(1).abs().max() < 1e-5
    xbad = xd.clone()
    xbad[5] = float("nan")
    fn = integ.metropolis_hastings_mala_adaptive if adaptive else integ.metropolis_hastings_mala
    args = dict(dt_init=1e-4) if adaptive else {}
    xm, rates = fn(xbad, tgt, return_acceptance_rate=True, **args)
    assert xm.shape == xbad.shape and len(rates) == 8 and all(0.0 <= r <= 1.0 for r in rates)
    assert torch.isnan(xm[-1]).all() and torch.isfinite(xm[:-1]).all()
    assert max(rates) > 0.3  # dt=1e-4 on relaxed LJ-13 configurations accepts most proposals


def _integ(**kw):
    from pita_b200.sde_integration import WeightedSDEIntegrator
    args = dict(sde=None, num_integration_steps=10, start_resampling_step=0, end_resampling_step=10, num_negative_time_steps=6,
                post_mcmc_steps=6, dt_negative_time=2e-4)
    args.update(kw)
    return WeightedSDEIntegrator(**args)


def _target(n=13):
    from pita_b200.lennardjones_energy import LennardJonesEnergy
    return LennardJonesEnergy(dimensionality=3 * n, n_particles=n)


def test_mala_proposal_vs_reference_golden():
    """mala_proposal (reference :28-45) on the reference's own noise draw."""
    from pita_b200 import sde_integration as SI
    g = golden("post_n13.npz")
    x = torch.from_numpy(g["prop.x"]).float().cuda()
    noise = torch.from_numpy(g["prop.noise"]).float().cuda()
    orig = torch.randn_like
    torch.randn_like = lambda t, *a, **k: noise  # the drop-in draws its proposal noise with torch.randn_like, like the reference
    try:
        xp, lqf, lqb = SI.mala_proposal(x, _target(), float(g["prop.dt"]))
    finally:
        torch.randn_like = orig
    assert_close(xp, g["prop.x_prop"], "x_prop", rtol=1e-5)
    assert_close(lqf, g["prop.log_q_fwd"], "log q(x'|x)")
    assert_close(lqb, g["prop.log_q_bwd"], "log q(x|x')")


@pytest.mark.parametrize("adaptive", [False, True])
def test_mala_loops_vs_reference_golden(adaptive):
    """metropolis_hastings_mala / _adaptive (reference :362-470): six steps on the reference's recorded proposal noise and
    uniforms, one non-finite particle.  Acceptance rates (hence every accept/reject decision count and, in the adaptive
    loop, the whole step-size trajectory) must match exactly; particles to 1e-4; the invalid row comes back last."""
    g = golden("post_n13.npz")
    tag = "mala_adaptive" if adaptive else "mala"
    integ = _integ(adaptive_mcmc=adaptive)
    noise = torch.from_numpy(g[tag + ".noise"]).float().cuda()
    uni = torch.from_numpy(g[tag + ".uniform"]).float().cuda()
    integ.mcmc_noise_fn = lambda k, x: noise[k]
    integ.mcmc_uniform_fn = lambda k, lp: uni[k]
    x0 = torch.from_numpy(g["x0"]).float().cuda()
    if adaptive:
        x, rates = integ.metropolis_hastings_mala_adaptive(x0, _target(), dt_init=2e-4, return_acceptance_rate=True)
    else:
        x, rates = integ.metropolis_hastings_mala(x0, _target(), return_acceptance_rate=True)
    np.testing.assert_allclose(np.array(rates), g[tag + ".rates"], rtol=0, atol=1e-7)
    ref = g[tag + ".x"]
    assert not np.isfinite(ref[-1]).all() and not torch.isfinite(x[-1]).all(), "the non-finite particle is moved to the end"
    assert_close(x[:-1], ref[:-1], "particles after MALA")


def test_mala_without_rates_is_the_references_noop():
    """Reference quirk pinned (:386, :401): return_acceptance_rate=False makes the non-adaptive loop a no-op."""
    g = golden("post_n13.npz")
    x0 = torch.from_numpy(g["x0"]).float().cuda()
    x, rates = _integ().metropolis_hastings_mala(x0, _target(), return_acceptance_rate=False)
    assert rates is None
    ref, bad = g["mala_norate.x"], int(g["bad_row"])
    keep = [r for r in range(ref.shape[0]) if r != bad]
    assert_close(x[keep], ref[keep], "particles (unchanged, original order)", rtol=1e-6)
    assert not torch.isfinite(x[bad]).all()


@pytest.mark.parametrize("langevin", [False, True])
def test_negative_time_descent_vs_reference_golden(langevin):
    g = golden("post_n13.npz")
    integ = _integ(do_langevin=langevin)
    if langevin:
        noise = torch.from_numpy(g["langevin.noise"]).float().cuda()
        integ.descent_noise_fn = lambda k, x: noise[k]
    x = integ.negative_time_descent(torch.from_numpy(g["prop.x"]).float().cuda(), _target())
    assert_close(x, g[("langevin" if langevin else "descent") + ".x"], "negative_time_descent")
