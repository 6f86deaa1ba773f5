"""Fused SDE/FK step, chunk-quantile clamp and the whole annealed loop vs the oracle (through the C-ABI)."""
import numpy as np
import pytest
import torch

import pita_oracle as O
from helpers import assert_close, golden, make_net, state_from_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,B", [(13, 1000), (55, 333)])
@pytest.mark.parametrize("debias", [True, False])
def test_fused_step_vs_formula(n, B, debias):
    from pita_b200 import ops
    D = 3 * n
    gen = torch.Generator().manual_seed(B)
    x, gu, sc, nz = (torch.randn(B, D, generator=gen, dtype=torch.float64) for _ in range(4))
    div, dEdh, en = (torch.randn(B, generator=gen, dtype=torch.float64) for _ in range(3))
    p = dict(g2=3.7, gamma=1.3333, dgamma_dt=0.21, dh_dt=5.5, dt=0.01, sqrt_dt=0.1, noise_scale=1.9)
    if debias:
        bt = sc * p["g2"] / 2
        dX = p["gamma"] * (-gu) * p["g2"] / 2 + p["gamma"] * bt
        raw = p["gamma"] ** 2 * (-gu * bt).sum(-1) + p["gamma"] * div * p["g2"] / 2 + p["gamma"] * dEdh * p["dh_dt"] + p["dgamma_dt"] * en
    else:
        dX = p["gamma"] * sc * p["g2"]
        raw = None
    ref = O.centre(x + dX * p["dt"] + p["noise_scale"] * nz * p["sqrt_dt"], n)
    c = lambda v: v.float().cuda()  # noqa: E731
    xo, a_raw = ops.sde_fk_step(c(x), c(gu), c(sc), c(nz), c(div), c(dEdh), c(en), n, debias=debias, want_a_raw=debias, **p)
    assert_close(xo, ref, "x_next", rtol=1e-5)
    if debias:
        assert_close(a_raw, raw, "a_raw", rtol=1e-5)
    # frozen step: x_out = remove_mean(x)
    xf, _ = ops.sde_fk_step(c(x), None, None, None, None, None, None, n, debias=False, freeze_x=True, want_a_raw=False, **p)
    assert_close(xf, O.centre(x, n), "frozen x", rtol=1e-6)
    # in-kernel Philox noise: unit variance, zero mean, reproducible, different per offset
    z = torch.zeros(B, D, device="cuda")
    q = dict(p, noise_scale=1.0, sqrt_dt=1.0)
    n1, _ = ops.sde_fk_step(z, z, z, None, None, None, None, n, debias=False, remove_mean=False, seed=5, offset=1, want_a_raw=False, **q)
    n2, _ = ops.sde_fk_step(z, z, z, None, None, None, None, n, debias=False, remove_mean=False, seed=5, offset=1, want_a_raw=False, **q)
    n3, _ = ops.sde_fk_step(z, z, z, None, None, None, None, n, debias=False, remove_mean=False, seed=5, offset=2, want_a_raw=False, **q)
    assert torch.equal(n1, n2) and not torch.equal(n1, n3)
    assert abs(n1.mean().item()) < 0.02 and abs(n1.std().item() - 1.0) < 0.02


@pytest.mark.parametrize("B,chunk", [(512, 512), (1000, 128), (5000, 512), (37, 512), (8192, 8192), (1, 16)])
def test_quantile_clamp_vs_torch(B, chunk):
    from pita_b200 import ops
    gen = torch.Generator().manual_seed(B + chunk)
    raw = torch.randn(B, generator=gen) * 10
    a = torch.randn(B, generator=gen)
    a_out, drift = ops.fk_quantile_accumulate(raw.cuda(), a.cuda(), chunk, 0.9, 0.01, False, want_drift=True)
    ref = torch.cat([O.quantile_clamp(raw[lo:lo + chunk]) for lo in range(0, B, chunk)])
    assert torch.equal(drift.cpu(), ref), (drift.cpu() - ref).abs().max()
    assert_close(a_out, a.double() + ref.double() * 0.01, "a_next", rtol=1e-6)
    z, _ = ops.fk_quantile_accumulate(raw.cuda(), a.cuda(), chunk, 0.9, 0.01, True)
    assert z.abs().max().item() == 0.0
    only, _ = ops.fk_quantile_accumulate(raw.cuda(), None, chunk, 0.9, 0.0, False)
    assert torch.equal(only.cpu(), ref)


def _build_integrator(n, sdE, sdS, S, chunk, **kw):
    from pita_b200.energy_net import EnergyNet
    from pita_b200.noise_schedules import ElucidatingNoiseSchedule
    from pita_b200.score_net import ScoreNet
    from pita_b200.sde_integration import WeightedSDEIntegrator
    from pita_b200.sdes import VEReverseSDE
    sde = VEReverseSDE(ElucidatingNoiseSchedule(0.05, 80.0, 7.0), energy_net=EnergyNet(make_net(n, sdE)),
                       score_net=ScoreNet(make_net(n, sdS)), debias_inference=kw.pop("debias", True))
    return WeightedSDEIntegrator(sde=sde, num_integration_steps=S, lightning_module=None, batch_size=chunk,
                                 num_negative_time_steps=0, post_mcmc_steps=0, **kw)


def _replay_reference_stream(g, sdE, sdS, gamma_sched=None, resample_at_end=True):
    """Replays the reference's CPU random stream through the oracle (which reproduces the reference trajectory,
    tests/test_oracle_golden.py) and returns the per-step noise / offsets it consumed."""
    n, N, S, chunk = int(g["n"]), int(g["N"]), int(g["S"]), int(g["chunk"])
    torch.manual_seed(int(g["seed"]))
    x1 = O.mean_free_prior(N, n, float(g["prior_scale"]), dtype=torch.float64)
    noises, u0s = {}, {}
    cfg = O.LoopConfig(n=n, steps=S, chunk=chunk, beta=float(g["beta"]), resampling_interval=int(g["interval"]),
                       start_resampling_step=int(g["start"]), end_resampling_step=int(g["end"]), resample_at_end=resample_at_end)

    def noise_fn(step, xc):
        z = torch.randn_like(xc)
        noises.setdefault(step, []).append(z)
        return z

    def u0_fn(step):
        u0s[step] = float(torch.rand(size=(1,), dtype=torch.float64))
        return u0s[step]

    gs = gamma_sched if gamma_sched is not None else O.ConstGamma(float(g["gamma"]))
    x_ref, logw_ref, uniq_ref = O.integrate(sdE, sdS, O.EDMSchedule(0.05), gs, cfg, x1, noise_fn, u0_fn)
    assert list(uniq_ref) == list(g["num_unique"])
    return x1, {k: torch.cat(v) for k, v in noises.items()}, u0s


def test_loop_vs_reference_golden():
    """integrate_sde on the reference's own 40-step trajectory fixture (fp64 reference, fp32 kernels), with the
    reference's noise / offsets injected.  (1) teacher-forced: every step starts from the reference's recorded
    state and must reproduce the next one to 1e-4 with identical ancestors; (2) free-running: identical
    ancestor counts at every step and the final particles / end log-weights within a looser bound (errors
    compound through 40 steps and 18 resamplings)."""
    from pita_b200.annealing_factor_schedules import ConstantAnnealingFactorSchedule
    from pita_b200.lennardjones_energy import LennardJonesEnergy
    g = golden("loop_n13.npz")
    n, N, S, chunk = int(g["n"]), int(g["N"]), int(g["S"]), int(g["chunk"])
    sdE, sdS = state_from_golden(g, "E."), state_from_golden(g, "S.")
    x1, noises, u0s = _replay_reference_stream(g, sdE, sdS)
    kw = dict(start_resampling_step=int(g["start"]), end_resampling_step=int(g["end"]), resampling_interval=int(g["interval"]),
              resample_at_end=True)
    gam = ConstantAnnealingFactorSchedule(float(g["gamma"]))
    tgt = LennardJonesEnergy(dimensionality=3 * n, n_particles=n, temperature=1.0)
    # (1) teacher forced
    integ = _build_integrator(n, sdE, sdS, S, chunk, **kw)
    integ.noise_fn = lambda step, x: noises[step].float().cuda()
    integ.u0_fn = lambda step: u0s[step]
    integ.prepare(N, 3 * n, "cuda")
    times = torch.linspace(1.0, 0.0, S + 1)[:-1]
    dt = 1.0 / S
    xs = np.concatenate([g["x1"][None].astype(np.float32), g["x_steps"]])
    a_prev = np.concatenate([np.zeros((1, N)), g["a_steps"]])
    for step in range(S):
        x_in = torch.from_numpy(xs[step]).cuda()
        a_in = torch.from_numpy(a_prev[step]).float().cuda()
        x_out, a_out, _, _ = integ._fk_step(float(times[step]), step, x_in, a_in, dt, float(np.float32(np.sqrt(dt))),
                                            float(g["beta"]), n, gam, tgt, int(g["interval"]))
        assert_close(x_out, xs[step + 1], f"x after step {step} (teacher forced)")
        assert_close(a_out, g["a_steps"][step], f"a after step {step} (teacher forced)")
    # (2) free running
    integ = _build_integrator(n, sdE, sdS, S, chunk, **kw)
    integ.noise_fn = lambda step, x: noises[step].float().cuda()
    integ.u0_fn = lambda step: u0s[step]
    x, logw, uniq, terms, acc = integ.integrate_sde(x1.float().cuda(), tgt, gam, inverse_temperature=float(g["beta"]))
    assert list(uniq) == list(g["num_unique"]), (uniq, list(g["num_unique"]))
    assert logw.shape == g["logweights"].shape
    assert_close(logw, g["logweights"], "logweights vs reference", rtol=2e-3)
    assert_close(x, g["x_final"], "x_final vs reference", rtol=2e-3)


@pytest.mark.parametrize("n,N,S,chunk,time_range", [(13, 64, 12, 32, 0.3), (55, 6, 3, 3, 0.06), (13, 50, 8, 16, 0.2)])
def test_loop_vs_oracle(n, N, S, chunk, time_range):
    """Free-running loop vs the fp64 oracle with injected noise/offsets, resampling every step, ragged last chunk
    (N % chunk != 0 in the third case).  Short time ranges keep the explicit Euler map contractive."""
    from pita_b200.annealing_factor_schedules import ConstantAnnealingFactorSchedule
    from pita_b200.lennardjones_energy import LennardJonesEnergy
    sdE = O.random_egnn_state(seed=31 + n, dtype=torch.float64, coord_gain=0.3)
    sdS = O.random_egnn_state(seed=32 + n, dtype=torch.float64, coord_gain=0.3)
    gam = 4.0 / 3.0
    sched = O.EDMSchedule(0.05)
    gen = torch.Generator().manual_seed(N)
    scale = float((sched.h(torch.tensor(time_range, dtype=torch.float64)) / gam) ** 0.5)
    x1 = O.centre(O.md_shaped_coords(N, n, seed=N, dtype=torch.float64) + scale * torch.randn(N, 3 * n, generator=gen, dtype=torch.float64), n)
    noise = {s: torch.randn(N, 3 * n, generator=gen, dtype=torch.float64) for s in range(S)}
    u0 = {s: float(torch.rand(1, generator=gen, dtype=torch.float64)) for s in range(S + 1)}
    cfg = O.LoopConfig(n=n, steps=S, chunk=chunk, beta=0.9, resampling_interval=1, time_range=time_range)
    cursor = {}

    def noise_fn(step, xc):
        lo = cursor.get(step, 0)
        cursor[step] = lo + xc.shape[0]
        return noise[step][lo:lo + xc.shape[0]]

    x_ref, logw_ref, uniq_ref = O.integrate(sdE, sdS, sched, O.ConstGamma(gam), cfg, x1, noise_fn, lambda s: u0[s])
    integ = _build_integrator(n, sdE, sdS, S, chunk, start_resampling_step=0, end_resampling_step=10 ** 9, resampling_interval=1,
                              time_range=time_range)
    integ.noise_fn = lambda step, x: noise[step].float().cuda()
    integ.u0_fn = lambda step: u0[step]
    tgt = LennardJonesEnergy(dimensionality=3 * n, n_particles=n)
    x, logw, uniq, _, _ = integ.integrate_sde(x1.float().cuda(), tgt, ConstantAnnealingFactorSchedule(gam), inverse_temperature=0.9)
    assert list(uniq) == list(uniq_ref)
    assert_close(x, x_ref, "x_final", rtol=1e-3)


def test_sample_histograms_match_oracle():
    """SURVEY §8d gate 3 (sample-level histograms): LJ-13, 256 particles, 6 steps with resampling every step, common random
    numbers.  The pooled inter-atomic distance histogram and the (log) target-energy histogram of the final samples must
    coincide with the fp64 oracle's: 1-D W2 <= 2 % of the oracle histogram's standard deviation (two INDEPENDENT oracle runs
    of this size differ by 16-46 %, so the bound is a coupling check, not a statistical one).  The in-kernel Philox noise
    path is run next to it and only has to land in that statistical band."""
    from pita_b200.annealing_factor_schedules import ConstantAnnealingFactorSchedule
    from pita_b200.lennardjones_energy import LennardJonesEnergy
    n, N, S, chunk, time_range = 13, 256, 6, 128, 0.2
    sdE = O.random_egnn_state(seed=31 + n, dtype=torch.float64, coord_gain=0.3)
    sdS = O.random_egnn_state(seed=32 + n, dtype=torch.float64, coord_gain=0.3)
    gam, sched = 4.0 / 3.0, O.EDMSchedule(0.05)
    gen = torch.Generator().manual_seed(1)
    scale = float((sched.h(torch.tensor(time_range, dtype=torch.float64)) / gam) ** 0.5)
    x1 = O.centre(O.md_shaped_coords(N, n, seed=1, dtype=torch.float64) + scale * torch.randn(N, 3 * n, generator=gen, dtype=torch.float64), n)
    noise = {s: torch.randn(N, 3 * n, generator=gen, dtype=torch.float64) for s in range(S)}
    u0 = {s: float(torch.rand(1, generator=gen, dtype=torch.float64)) for s in range(S + 1)}
    cfg = O.LoopConfig(n=n, steps=S, chunk=chunk, beta=0.9, resampling_interval=1, time_range=time_range)
    cursor = {}

    def noise_fn(step, xc):
        lo = cursor.get(step, 0)
        cursor[step] = lo + xc.shape[0]
        return noise[step][lo:lo + xc.shape[0]]

    x_ref, _, uniq_ref = O.integrate(sdE, sdS, sched, O.ConstGamma(gam), cfg, x1, noise_fn, lambda s: u0[s])

    def observables(x):
        x = x.detach().double().cpu().reshape(-1, n, 3)
        d = (x[:, :, None] - x[:, None]).norm(dim=-1)
        iu = torch.triu_indices(n, n, 1)
        pair = d[:, iu[0], iu[1]].reshape(-1).numpy()
        loge = np.log10(np.maximum(O.lj_energy(x.reshape(-1, 3 * n), n).numpy() + 100.0, 1.0))
        return pair, loge

    kw = dict(start_resampling_step=0, end_resampling_step=10 ** 9, resampling_interval=1, time_range=time_range)
    tgt = LennardJonesEnergy(dimensionality=3 * n, n_particles=n)
    sched_g = ConstantAnnealingFactorSchedule(gam)
    integ = _build_integrator(n, sdE, sdS, S, chunk, **kw)
    integ.noise_fn = lambda step, x: noise[step].float().cuda()
    integ.u0_fn = lambda step: u0[step]
    x, _, uniq, _, _ = integ.integrate_sde(x1.float().cuda(), tgt, sched_g, inverse_temperature=0.9)
    assert list(uniq) == list(uniq_ref)
    p_ref, e_ref = observables(x_ref)
    p_gpu, e_gpu = observables(x)
    assert O.w2_1d(p_gpu, p_ref) <= 0.02 * p_ref.std(), O.w2_1d(p_gpu, p_ref) / p_ref.std()
    assert O.w2_1d(e_gpu, e_ref) <= 0.02 * e_ref.std(), O.w2_1d(e_gpu, e_ref) / e_ref.std()
    # independent draws (in-kernel Philox noise, library u0): statistically compatible histograms
    integ = _build_integrator(n, sdE, sdS, S, chunk, fused_noise=True, noise_seed=7, **kw)
    xs, _, uniq_s, _, _ = integ.integrate_sde(x1.float().cuda(), tgt, sched_g, inverse_temperature=0.9)
    p_s, e_s = observables(xs)
    assert torch.isfinite(xs).all() and min(uniq_s) > N // 8
    assert O.w2_1d(p_s, p_ref) <= 1.0 * p_ref.std(), O.w2_1d(p_s, p_ref) / p_ref.std()
    assert O.w2_1d(e_s, e_ref) <= 1.0 * e_ref.std(), O.w2_1d(e_s, e_ref) / e_ref.std()


def test_loop_linear_gamma_vs_reference_golden():
    """integrate_sde under a LinearAnnealingFactorSchedule (gamma'(t) != 0: the dgamma/dt * U term of the FK drift, sdes.py:227,
    is live) on the reference's own 24-step trajectory, resampling every step: teacher-forced states to 1e-4, free-running
    ancestor counts identical."""
    from pita_b200.annealing_factor_schedules import LinearAnnealingFactorSchedule
    from pita_b200.lennardjones_energy import LennardJonesEnergy
    g = golden("loop_n13_linear.npz")
    n, N, S, chunk = int(g["n"]), int(g["N"]), int(g["S"]), int(g["chunk"])
    sdE, sdS = state_from_golden(g, "E."), state_from_golden(g, "S.")
    ga = [float(v) for v in g["gamma_args"]]
    x1, noises, u0s = _replay_reference_stream(g, sdE, sdS, gamma_sched=O.LinearGamma(ga[0], ga[1], ga[2], ga[3]), resample_at_end=False)
    gam = LinearAnnealingFactorSchedule(ga[0], ga[1], t_start=ga[2], t_end=ga[3])
    kw = dict(start_resampling_step=0, end_resampling_step=S, resampling_interval=1, resample_at_end=False)
    tgt = LennardJonesEnergy(dimensionality=3 * n, n_particles=n, temperature=1.0)
    integ = _build_integrator(n, sdE, sdS, S, chunk, **kw)
    integ.noise_fn = lambda step, x: noises[step].float().cuda()
    integ.u0_fn = lambda step: u0s[step]
    integ.prepare(N, 3 * n, "cuda")
    times = torch.linspace(1.0, 0.0, S + 1)[:-1]
    dt = 1.0 / S
    xs = np.concatenate([g["x1"][None].astype(np.float32), g["x_steps"]])
    seen_dgamma = False
    for step in range(S):
        seen_dgamma |= float(gam.dgamma_dt(times[step])) != 0.0
        x_in = torch.from_numpy(xs[step]).cuda()
        a_in = torch.zeros(N, device="cuda")  # resampling every step: a is reset before every step
        x_out, a_out, _, _ = integ._fk_step(float(times[step]), step, x_in, a_in, dt, float(np.float32(np.sqrt(dt))),
                                            float(g["beta"]), n, gam, tgt, 1)
        assert_close(x_out, xs[step + 1], f"x after step {step} (teacher forced)")
    assert seen_dgamma
    integ = _build_integrator(n, sdE, sdS, S, chunk, **kw)
    integ.noise_fn = lambda step, x: noises[step].float().cuda()
    integ.u0_fn = lambda step: u0s[step]
    x, logw, uniq, _, _ = integ.integrate_sde(x1.float().cuda(), tgt, gam, inverse_temperature=float(g["beta"]))
    assert list(uniq) == list(g["num_unique"]), (uniq, list(g["num_unique"]))
    assert_close(x, g["x_final"], "x_final vs reference", rtol=2e-3)


@pytest.mark.parametrize("tag", ["pin", "precond"])
def test_sde_f_variants_vs_reference_golden(tag):
    """VEReverseSDE.f with pin_energy=True (energy_net.py:41-48) and with precondition_beta=True on both wrappers
    (energy_net.py:38-39, score_net.py:37-38) against the reference's fp64 SDETerms."""
    from pita_b200.annealing_factor_schedules import ConstantAnnealingFactorSchedule
    from pita_b200.energy_net import EnergyNet
    from pita_b200.lennardjones_energy import LennardJonesEnergy
    from pita_b200.noise_schedules import ElucidatingNoiseSchedule
    from pita_b200.score_net import ScoreNet
    from pita_b200.sdes import VEReverseSDE
    g = golden("fk_n13_%s.npz" % tag)
    n, t, beta = int(g["n"]), float(g["t"]), float(g["beta"])
    pin, pre = bool(int(g["pin"])), bool(int(g["precondition_beta"]))
    sched = ElucidatingNoiseSchedule(float(g["sigma_min"]), 80.0, 7.0)
    en = EnergyNet(make_net(n, state_from_golden(g, "E.")), precondition_beta=pre)
    sn = ScoreNet(make_net(n, state_from_golden(g, "S.")), precondition_beta=pre)
    sde = VEReverseSDE(sched, energy_net=en, score_net=sn, pin_energy=pin)
    tgt = LennardJonesEnergy(dimensionality=3 * n, n_particles=n, temperature=1.0)
    x = torch.from_numpy(g["x"]).float().cuda()
    assert_close(tgt(x), g["target_logp"], "target log-prob")
    terms = sde.f(torch.tensor(t), x, torch.tensor(beta), ConstantAnnealingFactorSchedule(float(g["gamma"])), 1.0, tgt,
                  resampling_interval=1)
    assert_close(terms.drift_X, g["drift_X"], "drift_X")
    assert_close(terms.divergence_score, g["divergence_score"], "div_b")
    assert_close(terms.cross_term, g["cross_term"], "cross_term")
    assert_close(terms.dUt_dt, g["dUt_dt"], "dUt_dt")
    assert_close(terms.drift_A, g["drift_A"], "drift_A")
    B = x.shape[0]
    tt = torch.full((B,), t, device="cuda")
    ht = torch.full((B,), float(sched.h(torch.tensor(t, dtype=torch.float64))), device="cuda")
    assert_close(en.forward_energy(ht, x, beta, pin=pin, energy_function=tgt, t=tt), g["U"], "forward_energy")
    assert_close(en.forward(ht, x, beta, pin=pin, energy_function=tgt, t=tt), g["gradU"], "grad U")


def test_quantile_chunks_above_the_smem_sort():
    """inference_batch_size above 8192 (e.g. the reference's default batch_size=None = the whole shard): same chunk
    partition as the reference (sde_integration.py:312-343), each chunk clamped at its own torch.quantile (sdes.py:230)."""
    from pita_b200 import ops
    gen = torch.Generator().manual_seed(3)
    B = 30000
    raw = torch.randn(B, generator=gen) * 7
    a = torch.randn(B, generator=gen)
    for chunk in (B, 12000):
        a_out, drift = ops.fk_quantile_accumulate(raw.cuda(), a.cuda(), chunk, 0.9, 0.01, False, want_drift=True)
        ref = torch.cat([torch.clamp(raw[lo:lo + chunk], max=torch.quantile(raw[lo:lo + chunk], 0.9)) for lo in range(0, B, chunk)])
        assert_close(drift, ref, "drift_A", rtol=1e-6)
        assert_close(a_out, a.double() + ref.double() * 0.01, "a_next", rtol=1e-6)
        only, _ = ops.fk_quantile_accumulate(raw.cuda(), None, chunk, 0.9, 0.0, False)
        assert_close(only, ref, "clamped values", rtol=1e-6)


def test_prior_sample():
    """Prior / MeanFreePrior.sample (energies/base_prior.py:77-83): shape, centre of mass removed per sample, per-coordinate
    variance scale^2 (n-1)/n, log_prob == the reference's closed form."""
    import math

    from pita_b200.base_prior import Prior
    n, scale, N = 13, 2.5, 20000
    pr = Prior(scale, n_particles=n, spatial_dim=3, device="cuda")
    torch.manual_seed(0)
    x = pr.sample(N)
    assert x.shape == (N, 3 * n) and x.is_cuda and x.dtype == torch.float32
    assert x.reshape(N, n, 3).mean(1).abs().max().item() < 1e-5
    var = x.double().var().item()
    assert abs(var / (scale ** 2 * (n - 1) / n) - 1.0) < 0.02
    lp = pr.log_prob(x[:7])
    ref = -0.5 * (x[:7].double() ** 2).sum(1) / scale ** 2 - 0.5 * (n - 1) * 3 * math.log(2 * math.pi * scale ** 2)
    assert_close(lp, ref, "log_prob", rtol=1e-5)
    free = Prior(scale, n_particles=n, spatial_dim=3, device="cuda", should_mean_free=False)
    assert free.sample(5).shape == (5, 3 * n)


def test_generate_samples_with_the_real_integrator():
    """SURVEY §8f-2 on the GPU: `generate_samples` (energytemp_module.py:237-298) driving the real prior and the real
    integrator: the first pass resamples, the second (log-weight) pass runs the first `inference_batch_size` prior samples with
    resampling off.  Both passes are replayed through the fp64 oracle from the SAME prior samples / noise / offsets."""
    from functools import partial

    from pita_b200.annealing_factor_schedules import ConstantAnnealingFactorSchedule
    from pita_b200.base_prior import Prior
    from pita_b200.lennardjones_energy import LennardJonesEnergy
    from pita_b200.noise_schedules import ElucidatingNoiseSchedule
    from pita_b200.sampling import generate_samples
    n, N, S, chunk, time_range, gam, beta = 13, 72, 6, 24, 0.2, 4.0 / 3.0, 0.9
    sdE = O.random_egnn_state(seed=71, dtype=torch.float64, coord_gain=0.3)
    sdS = O.random_egnn_state(seed=72, dtype=torch.float64, coord_gain=0.3)
    gen = torch.Generator().manual_seed(5)
    noise = {s: torch.randn(N, 3 * n, generator=gen, dtype=torch.float64) for s in range(S)}
    u0 = {s: float(torch.rand(1, generator=gen, dtype=torch.float64)) for s in range(S + 1)}
    integ = _build_integrator(n, sdE, sdS, S, chunk, start_resampling_step=0, end_resampling_step=10 ** 9, resampling_interval=1,
                              time_range=time_range)
    integ.noise_fn = lambda step, x: noise[step][: x.shape[0]].float().cuda()
    integ.u0_fn = lambda step: u0[step]
    drawn = []

    def recording_prior(scale, device):
        pr = Prior(scale, n_particles=n, spatial_dim=3, device=device)
        sample = pr.sample
        pr.sample = lambda k: drawn.append(sample(k)) or drawn[-1]
        return pr

    sched = ElucidatingNoiseSchedule(0.05, 80.0, 7.0)
    torch.manual_seed(11)
    out = generate_samples(weighted_sde_integrator=integ, energy_function=LennardJonesEnergy(dimensionality=3 * n, n_particles=n),
                           num_samples=N, noise_schedule=sched, annealing_factor_schedule=partial(ConstantAnnealingFactorSchedule),
                           partial_prior=recording_prior, t_start=torch.tensor(time_range), device="cuda", inference_batch_size=chunk,
                           num_integration_steps=S, inverse_temp=beta, annealing_factor=gam, return_logweights=True)
    samples, not_resampled, logw, uniq, terms, acc = out
    assert len(drawn) == 1 and drawn[0].shape == (N, 3 * n)
    x1 = drawn[0].double().cpu()
    # the prior really has the reference's scale sqrt(h(t_start) / gamma)
    osched = O.EDMSchedule(0.05)
    scale = float((osched.h(torch.tensor(time_range, dtype=torch.float64)) / gam) ** 0.5)
    assert abs(float(x1.std()) / (scale * (1 - 1 / n) ** 0.5) - 1.0) < 0.1
    assert samples.shape == (N, 3 * n) and not_resampled.shape == (chunk, 3 * n) and logw.shape[1] == chunk

    def replay(x_start, interval):
        cursor = {}

        def noise_fn(step, xc):
            lo = cursor.get(step, 0)
            cursor[step] = lo + xc.shape[0]
            return noise[step][lo:lo + xc.shape[0]]

        cfg = O.LoopConfig(n=n, steps=S, chunk=chunk, beta=beta, resampling_interval=interval, time_range=time_range)
        return O.integrate(sdE, sdS, osched, O.ConstGamma(gam), cfg, x_start, noise_fn, lambda s: u0[s])

    x_ref, _, uniq_ref = replay(x1, 1)
    assert list(uniq) == list(uniq_ref) and min(uniq) < N
    assert_close(samples, x_ref, "generate_samples: resampled pass", rtol=1e-3)
    x_nr, logw_nr, uniq_nr = replay(x1[:chunk], S + 1)
    assert list(uniq_nr) == [chunk] * S
    assert_close(not_resampled, x_nr, "generate_samples: log-weight pass particles", rtol=1e-3)
    assert_close(logw, logw_nr, "generate_samples: log-weights", rtol=1e-3)
