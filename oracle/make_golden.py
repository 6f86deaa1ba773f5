"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python oracle/make_golden.py
The reference is a Python tree and cannot travel to the GPU box, so its outputs on seeded
synthetic inputs are committed as small fixtures; this script is the committed recipe.
Every fixture stores inputs, weights and the reference's outputs (fp64 = reference modules
cast with .double(); fp32 = the reference as shipped).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from _ref_import import import_reference  # noqa: E402
import pita_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)
ref = import_reference()


def make_net(n, seed, strong):
    torch.manual_seed(seed)
    net = ref.egnn.EGNN_dynamics(n_particles=n, n_dimension=3, hidden_nf=32, n_layers=3,
                                 act_fn=torch.nn.SiLU(), recurrent=True, tanh=True, attention=True,
                                 condition_time=True, condition_temperature=True, agg="sum")
    if strong:  # make the coordinate branch numerically visible (random init has gain 1e-3)
        with torch.no_grad():
            for l in range(3):
                getattr(net.egnn, f"gcl_{l}").coord_mlp[2].weight.mul_(300.0)
    return net


def sd_np(net, prefix):
    return {prefix + k: v.detach().double().numpy() for k, v in net.state_dict().items()}


class FakeTrainer:
    world_size = 1
    global_rank = 0
    num_nodes = 1


class FakeLM:
    trainer = FakeTrainer()

    @staticmethod
    def all_gather(v):
        if isinstance(v, dict):
            return {k: (None if t is None else t.unsqueeze(0)) for k, t in v.items()}
        return v.unsqueeze(0)


class LJTarget:
    """LennardJonesEnergy.__call__ (lennardjones_energy.py:213-227) around the reference's own
    LennardJonesPotential; the reference class itself eagerly loads external .npy datasets."""
    is_molecule = True
    n_spatial_dim = 3

    def __init__(self, n, temperature=1.0, dtype=torch.float32):
        self.n_particles = n
        self.temperature = temperature
        self.pot = ref.lj.LennardJonesPotential(dim=3 * n, n_particles=n, two_event_dims=False,
                                                temperature=temperature)

    def __call__(self, samples, return_force=False):
        with torch.enable_grad():
            s = samples.detach().clone().requires_grad_(True)
            lp = self.pot._log_prob(s).squeeze(-1)
            if return_force:
                f = torch.autograd.grad(lp.sum(), s)[0]
                return lp.detach(), f.detach()
            return lp.detach()


def golden_egnn_and_fk():
    for n, B in ((13, 6), (55, 3)):
        for strong in (False, True):
            tag = f"n{n}_{'strong' if strong else 'init'}"
            net_e = make_net(n, 12345, strong)
            net_s = make_net(n, 54321 if strong else 12345, strong)
            gen = torch.Generator().manual_seed(1000 + n)
            sched = ref.noise.ElucidatingNoiseSchedule(sigma_min=0.05, sigma_max=80, rho=7)
            t = 0.37 if strong else 0.81
            x32 = O.md_shaped_coords(B, n, seed=7 + n) * (1.0 + float(sched.h(torch.tensor(t))) ** 0.5 * 0.3)
            x32 = x32 + 0.3 * torch.randn(B, 3 * n, generator=gen)
            x32 = ref.data_utils.remove_mean(x32, n, 3)
            beta = 0.75
            out = {"n": n, "t": t, "beta": beta, "x": x32.double().numpy(), "sigma_min": 0.05,
                   "gamma": 4.0 / 3.0}
            out.update(sd_np(net_e, "E."))
            out.update(sd_np(net_s, "S."))
            # raw EGNN forward, fp32 as shipped and fp64
            tc = torch.linspace(-0.4, 0.9, B)
            bb = torch.linspace(0.5, 1.5, B)
            out["egnn_tcond"], out["egnn_beta"] = tc.double().numpy(), bb.double().numpy()
            out["egnn_out_f32"] = net_s(tc, x32, bb).detach().numpy()
            for dt_name, dt in (("f64", torch.float64),):
                torch.set_default_dtype(dt)
                ne, ns_ = net_e.double(), net_s.double()
                x = x32.double()
                out["egnn_out_f64"] = ns_(tc.double(), x, bb.double()).detach().numpy()
                en = ref.energy_net.EnergyNet(ne)
                sn = ref.score_net.ScoreNet(ns_)
                sde = ref.sdes.VEReverseSDE(sched, energy_net=en, score_net=sn, pin_energy=False,
                                            debias_inference=True,
                                            cdf=lambda h, xx, b_, _f=sn.forward: ref.utils.compute_divergence_exact(_f, h, xx, b_))
                sde.trainer = FakeTrainer()
                gs = ref.anneal.ConstantAnnealingFactorSchedule(4.0 / 3.0)
                terms = sde.f(torch.tensor(t), x.clone(), torch.tensor(beta), gs, 1.0, None, resampling_interval=1)
                out["drift_X"] = terms.drift_X.detach().numpy()
                out["drift_A"] = terms.drift_A.detach().numpy()
                out["div_b"] = terms.divergence_score.detach().numpy()
                out["cross"] = terms.cross_term.detach().numpy()
                out["dUt_dt"] = terms.dUt_dt.detach().numpy()
                tt = torch.full((B,), t)
                ht = sched.h(tt)
                xr = x.clone().requires_grad_(True)
                out["U"] = en.forward_energy(ht, xr, torch.tensor(beta)).detach().numpy()
                out["gradU"] = en.forward(ht, xr, torch.tensor(beta)).detach().numpy()
                out["score"] = sn.forward(ht, x, torch.tensor(beta)).detach().numpy()
                out["div_score"] = ref.utils.compute_divergence_exact(sn.forward, ht, x, torch.tensor(beta)).numpy()
                # not-debiased branch (sdes.py:117-128)
                sde_nd = ref.sdes.VEReverseSDE(sched, energy_net=en, score_net=sn, debias_inference=False,
                                               cdf=lambda *a: None)
                out["drift_X_nodebias"] = sde_nd.f(torch.tensor(t), x.clone(), torch.tensor(beta), gs, 1.0,
                                                   None).drift_X.detach().numpy()
                torch.set_default_dtype(torch.float32)
            np.savez_compressed(os.path.join(OUT, f"fk_{tag}.npz"), **out)
            print("wrote fk_%s" % tag, "drift_A", out["drift_A"][:3])


def golden_resample():
    out = {}
    cases = [(16, 1, 3.0), (1000, 2, 3.0), (4096, 3, 0.5), (65536, 4, 3.0), (100003, 6, 2.0), (1 << 18, 5, 3.0)]
    for N, seed, scale in cases:
        g = torch.Generator().manual_seed(seed)
        logits = torch.randn(N, generator=g) * scale
        torch.manual_seed(100 + seed)
        u0 = float(torch.rand(size=(1,), dtype=torch.float64))
        torch.manual_seed(100 + seed)
        ids, _ = ref.utils.sample_cat_sys(N, logits)
        w = torch.clip(torch.softmax(logits, dim=-1), 1e-6, 1.0)
        # claim used by the oracle: torch CPU cumsum(fp32) == fp32(cumsum in fp64)
        bins_t = torch.cumsum(w, dim=-1).numpy()
        bins_n = np.cumsum(w.numpy().astype(np.float64)).astype(np.float32)
        out[f"cumsum_claim_{N}"] = np.array(int(np.array_equal(bins_t, bins_n)))
        if N <= 65536:
            out[f"logits_{N}"] = logits.numpy()
        out[f"weights_{N}"] = w.numpy()
        out[f"u0_{N}"] = np.array(u0)
        if N <= 65536:
            out[f"ids_{N}"] = ids.astype(np.int32)
        else:  # keep the fixture small: strided sample + run-length summary
            out[f"ids_{N}_stride64"] = ids[::64].astype(np.int32)
            out[f"ids_{N}_sum"] = np.array(int(ids.sum()))
            out[f"ids_{N}_unique"] = np.array(len(np.unique(ids)))
        print("resample N=%d u0=%.6f cumsum-claim=%s unique=%d sumw=%.4f" % (
            N, u0, bool(out[f"cumsum_claim_{N}"]), len(np.unique(ids)), float(w.double().sum())))
    # edge-case offsets
    g = torch.Generator().manual_seed(9)
    logits = torch.randn(257, generator=g) * 2
    w = torch.clip(torch.softmax(logits, dim=-1), 1e-6, 1.0)
    for name, u0 in (("zero", 0.0), ("almost1", 1.0 - 2.0 ** -53), ("half", 0.5)):
        u = (torch.tensor([u0], dtype=torch.float64) + 1 / 257 * torch.arange(257)) % 1.0
        bins = torch.cumsum(w, dim=-1)
        ids = np.digitize(u, bins.cpu(), right=True)
        ids[ids == 257] = 256
        out[f"edge_{name}_ids"] = ids.astype(np.int32)
    out["edge_logits"] = logits.numpy()
    out["case_sizes"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(OUT, "resample.npz"), **out)


def golden_lj():
    out = {}
    for n, B in ((13, 64), (55, 16)):
        x = O.md_shaped_coords(B, n, seed=21 + n)
        for T in (1.0, 2.5):
            tgt = LJTarget(n, temperature=T)
            lp, f = tgt(x, return_force=True)
            pot64 = ref.lj.LennardJonesPotential(dim=3 * n, n_particles=n, two_event_dims=False, temperature=T)
            xd = x.double().requires_grad_(True)
            lp64 = pot64._log_prob(xd).squeeze(-1)
            f64 = torch.autograd.grad(lp64.sum(), xd)[0]
            out[f"x_{n}"] = x.numpy()
            out[f"logp_f32_{n}_T{T}"], out[f"force_f32_{n}_T{T}"] = lp.numpy(), f.numpy()
            out[f"logp_f64_{n}_T{T}"], out[f"force_f64_{n}_T{T}"] = lp64.detach().numpy(), f64.numpy()
        # in-repo independent restatement energy2 (sampling/sample_lj13.py:24-30), eps-free
        v = x.double().reshape(B, n, 3)
        d = torch.vmap(torch.pdist)(v)
        e2 = 2 * ((1 / d) ** 12 - 2 * (1 / d) ** 6).sum(-1) + 0.5 * (v - v.mean(1, keepdim=True)).pow(2).sum((-2, -1))
        out[f"energy2_neg_{n}"] = (-e2).numpy()
    np.savez_compressed(os.path.join(OUT, "lj.npz"), **out)
    print("wrote lj")


def golden_loop():
    """integrate_sde end to end (sde_integration.py:98-212), fp64 default dtype for a tight pin.  S=40 steps keep
    the explicit Euler step inside its stability region (dt * dlog h/dt < 1), so fp32 and fp64 runs stay comparable;
    the state after every step is recorded (by wrapping the reference's own step method) for teacher-forced checks."""
    n, N, S, chunk = 13, 32, 40, 16
    torch.set_default_dtype(torch.float64)
    try:
        net_e = make_net(n, 12345, True).double()
        net_s = make_net(n, 54321, True).double()
        sched = ref.noise.ElucidatingNoiseSchedule(sigma_min=0.05, sigma_max=80, rho=7)
        en, sn = ref.energy_net.EnergyNet(net_e), ref.score_net.ScoreNet(net_s)
        sde = ref.sdes.VEReverseSDE(sched, energy_net=en, score_net=sn, pin_energy=False, debias_inference=True,
                                    cdf=lambda h, xx, b_, _f=sn.forward: ref.utils.compute_divergence_exact(_f, h, xx, b_))
        sde.trainer = FakeTrainer()
        gam = 4.0 / 3.0
        integ = ref.integ.WeightedSDEIntegrator(
            sde=sde, num_integration_steps=S, start_resampling_step=2, end_resampling_step=36,
            lightning_module=FakeLM(), partial_annealing_factor_schedule=None, resampling_interval=2,
            num_negative_time_steps=0, post_mcmc_steps=0, batch_size=chunk, resample_at_end=True,
            diffusion_scale=1.0)
        states = []
        inner = integ.ddp_batched_euler_maruyama_step

        def recording_step(t, x, a, dt, step, **kw):
            out = inner(t, x, a, dt, step, **kw)
            states.append((ref.data_utils.remove_mean(out[0], n, 3).detach().clone(), out[1].detach().clone()))
            return out

        integ.ddp_batched_euler_maruyama_step = recording_step
        tgt = LJTarget(n, temperature=1.0)
        torch.manual_seed(2024)
        scale = float((sched.h(torch.tensor(1.0)) / gam) ** 0.5)
        x1 = ref.prior.Prior(scale, n_particles=n, spatial_dim=3).sample(N)
        x1_in = x1.clone()
        x, logw, uniq, terms, acc = integ.integrate_sde(
            x1, tgt, ref.anneal.ConstantAnnealingFactorSchedule(gam), inverse_temperature=torch.tensor(0.75))
        out = {"n": n, "N": N, "S": S, "chunk": chunk, "seed": 2024, "gamma": gam, "beta": 0.75,
               "start": 2, "end": 36, "interval": 2,
               "x1": x1_in.detach().numpy(), "x_final": x.detach().numpy(), "logweights": logw.detach().numpy(),
               "num_unique": np.array(uniq), "prior_scale": scale,
               "x_steps": torch.stack([s_[0] for s_ in states]).numpy().astype(np.float32),
               "a_steps": torch.stack([s_[1] for s_ in states]).numpy()}
        out.update(sd_np(net_e, "E."))
        out.update(sd_np(net_s, "S."))
        np.savez_compressed(os.path.join(OUT, "loop_n13.npz"), **out)
        print("wrote loop", uniq)
    finally:
        torch.set_default_dtype(torch.float32)


def golden_schedules():
    """Noise schedules (noise_schedules.py:19-28, 64-125) and annealing-factor schedules (annealing_factor_schedules.py:20-109)
    of the reference on a grid of times, fp64 — pins the host-side mirrors in pita_b200/ and the oracle's restatements."""
    t = torch.linspace(0.0, 1.0, 41, dtype=torch.float64)
    out = {"t": t.numpy()}
    for tag, sch in (("edm005", ref.noise.ElucidatingNoiseSchedule(0.05, 80, 7)), ("edm001", ref.noise.ElucidatingNoiseSchedule(0.01, 80, 7)),
                     ("geo", ref.noise.GeometricNoiseSchedule(0.01, 3.0)), ("lin", ref.noise.LinearNoiseSchedule(2.5))):
        out[tag + ".h"] = torch.as_tensor(sch.h(t), dtype=torch.float64).numpy()
        out[tag + ".g"] = torch.as_tensor(sch.g(t), dtype=torch.float64).numpy()
        if hasattr(sch, "dh_dt"):
            out[tag + ".dh_dt"] = torch.as_tensor(sch.dh_dt(t), dtype=torch.float64).numpy()
        if hasattr(sch, "t"):
            out[tag + ".t_of_h"] = torch.as_tensor(sch.t(sch.h(t)), dtype=torch.float64).numpy()
    for tag, sch in (("const", ref.anneal.ConstantAnnealingFactorSchedule(4.0 / 3.0)),
                     ("linear", ref.anneal.LinearAnnealingFactorSchedule(1.5, 1.0, t_start=0.8, t_end=0.2)),
                     ("sigmoid", ref.anneal.SigmoidAnnealingFactorSchedule(1.5, 1.0, t_start=0.9, t_end=0.1, sharpness=8.0))):
        out[tag + ".gamma"] = torch.as_tensor(sch.gamma(t), dtype=torch.float64).numpy()
        out[tag + ".dgamma_dt"] = torch.as_tensor(sch.dgamma_dt(t), dtype=torch.float64).numpy()
    np.savez_compressed(os.path.join(OUT, "schedules.npz"), **out)
    print("wrote schedules", {k: v.shape for k, v in out.items() if k.endswith(".h")})


def golden_laplacian():
    """No-score-net branch of VEReverseSDE.f (sdes.py:150-153, 204-216): b = -grad U g^2/2 and
    div b = -laplacian(U) g^2/2 through compute_laplacian_exact.  LJ-13, strong coordinate gain, fp64."""
    n, B, t, beta = 13, 4, 0.37, 0.75
    net_e = make_net(n, 12345, True)
    gen = torch.Generator().manual_seed(1313)
    sched = ref.noise.ElucidatingNoiseSchedule(sigma_min=0.05, sigma_max=80, rho=7)
    x32 = O.md_shaped_coords(B, n, seed=20) * (1.0 + float(sched.h(torch.tensor(t))) ** 0.5 * 0.3)
    x32 = ref.data_utils.remove_mean(x32 + 0.3 * torch.randn(B, 3 * n, generator=gen), n, 3)
    out = {"n": n, "t": t, "beta": beta, "x": x32.double().numpy(), "sigma_min": 0.05, "gamma": 4.0 / 3.0}
    out.update(sd_np(net_e, "E."))
    torch.set_default_dtype(torch.float64)
    try:
        en = ref.energy_net.EnergyNet(net_e.double())
        sde = ref.sdes.VEReverseSDE(sched, energy_net=en, score_net=None, pin_energy=False, debias_inference=True,
                                    cdf=lambda *a: None)  # (the constructor dereferences score_net.forward when cdf is None, sdes.py:113)
        sde.trainer = FakeTrainer()
        gs = ref.anneal.ConstantAnnealingFactorSchedule(4.0 / 3.0)
        terms = sde.f(torch.tensor(t), x32.double().clone(), torch.tensor(beta), gs, 1.0, None, resampling_interval=1)
        out["drift_X"] = terms.drift_X.detach().numpy()
        out["drift_A"] = terms.drift_A.detach().numpy()
        out["div_b"] = terms.divergence_score.detach().numpy()
        out["cross"] = terms.cross_term.detach().numpy()
    finally:
        torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(OUT, "fk_n13_laplacian.npz"), **out)
    print("wrote laplacian", out["div_b"])


def golden_ad2():
    """Alanine-dipeptide EGNN (SURVEY §8 row a8'): EGNN_dynamics_AD2_cat, 22 atoms, hidden 64, 5 layers, condition_beta,
    random init (seed 12345) with the coordinate gain raised like the LJ "strong" fixtures; weights are rounded to fp32 for
    the fixture and the reference is evaluated in fp64 FROM those rounded weights."""
    from _ref_import import import_reference_ad2
    ad2 = import_reference_ad2()
    n, B = 22, 6
    torch.manual_seed(12345)
    net = ad2.EGNN_dynamics_AD2_cat(n_particles=n, n_dimensions=3, hidden_nf=64, n_layers=5, act_fn=torch.nn.SiLU(),
                                    recurrent=True, attention=True, tanh=True, agg="sum", condition_beta=True)
    with torch.no_grad():
        for l in range(5):
            getattr(net.egnn, f"gcl_{l}").coord_mlp[2].weight.mul_(300.0)
    net = net.double()
    with torch.no_grad():
        for p_ in net.parameters():
            p_.copy_(p_.float().double())
    gen = torch.Generator().manual_seed(22)
    # MD-shaped synthetic coordinates in the reference's normalised units (lattice recipe / 0.164-style scale)
    side = 3
    sites = torch.stack(torch.meshgrid(*[torch.arange(side, dtype=torch.float64)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n] * 1.1
    x = sites.reshape(1, 3 * n).repeat(B, 1) + 0.08 * torch.randn(B, 3 * n, generator=gen, dtype=torch.float64)
    x = (x.reshape(B, n, 3) - x.reshape(B, n, 3).mean(1, keepdim=True)).reshape(B, 3 * n)
    t = torch.linspace(0.05, 0.95, B, dtype=torch.float64)
    beta = torch.full((B,), 1.25, dtype=torch.float64)
    with torch.no_grad():
        vel = net(t, x, beta)
    out = {"n": n, "x": x.numpy(), "t": t.numpy(), "beta": beta.numpy(), "vel": vel.numpy()}
    out.update({"W." + k: v.detach().float().numpy() for k, v in net.state_dict().items()})
    # the same network inside the reference's EnergyNet / ScoreNet / VEReverseSDE.f (fp64; energy and score share the weights)
    torch.set_default_dtype(torch.float64)
    try:
        sched = ref.noise.ElucidatingNoiseSchedule(sigma_min=0.05, sigma_max=80, rho=7)
        en, sn = ref.energy_net.EnergyNet(net), ref.score_net.ScoreNet(net)
        sde = ref.sdes.VEReverseSDE(sched, energy_net=en, score_net=sn, pin_energy=False, debias_inference=True,
                                    cdf=lambda h, xx, b_, _f=sn.forward: ref.utils.compute_divergence_exact(_f, h, xx, b_))
        sde.trainer = FakeTrainer()
        gs = ref.anneal.ConstantAnnealingFactorSchedule(4.0 / 3.0)
        t_fk, beta_fk = 0.37, 0.75
        xs = x * (1.0 + float(sched.h(torch.tensor(t_fk))) ** 0.5 * 0.3)
        terms = sde.f(torch.tensor(t_fk), xs.clone(), torch.tensor(beta_fk), gs, 1.0, None, resampling_interval=1)
        ht = sched.h(torch.full((B,), t_fk))
        xr = xs.clone().requires_grad_(True)
        out.update({"fk_t": t_fk, "fk_beta": beta_fk, "fk_gamma": 4.0 / 3.0, "fk_x": xs.numpy(), "sigma_min": 0.05,
                    "drift_X": terms.drift_X.detach().numpy(), "drift_A": terms.drift_A.detach().numpy(),
                    "div_b": terms.divergence_score.detach().numpy(), "cross": terms.cross_term.detach().numpy(),
                    "dUt_dt": terms.dUt_dt.detach().numpy(),
                    "U": en.forward_energy(ht, xr, torch.tensor(beta_fk)).detach().numpy(),
                    "gradU": en.forward(ht, xr, torch.tensor(beta_fk)).detach().numpy(),
                    "score": sn.forward(ht, xs, torch.tensor(beta_fk)).detach().numpy(),
                    "div_score": ref.utils.compute_divergence_exact(sn.forward, ht, xs, torch.tensor(beta_fk)).numpy()})
    finally:
        torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(OUT, "egnn_ad2_n22.npz"), **out)
    print("wrote ad2", float(vel.abs().max()))


class _RecordDraws:
    """Records every torch.randn_like / torch.rand_like the reference makes (the post-processing draws its proposal noise and
    its accept/reject uniforms from the global generator), so that the CUDA path can be fed the same draws."""

    def __enter__(self):
        self.randn, self.rand = [], []
        self._rn, self._r = torch.randn_like, torch.rand_like

        def randn_like(x, *a, **k):
            v = self._rn(x, *a, **k)
            self.randn.append(v.detach().clone())
            return v

        def rand_like(x, *a, **k):
            v = self._r(x, *a, **k)
            self.rand.append(v.detach().clone())
            return v

        torch.randn_like, torch.rand_like = randn_like, rand_like
        return self

    def __exit__(self, *exc):
        torch.randn_like, torch.rand_like = self._rn, self._r


def golden_post():
    """Post-processing on the target (sde_integration.py:28-45, 353-470): mala_proposal,
    metropolis_hastings_mala, metropolis_hastings_mala_adaptive and negative_time_descent of the UNMODIFIED reference on
    LJ-13, fp64, with one non-finite particle (filtered and moved to the end by the reference, :367-400).  All random draws
    are recorded.  The accept/reject margins |log u - log ratio| are asserted > 1e-3 so that an fp32 evaluation takes the
    same decisions."""
    n, N, steps = 13, 40, 6
    torch.set_default_dtype(torch.float64)
    try:
        tgt = LJTarget(n, temperature=1.0)
        x0 = O.md_shaped_coords(N, n, seed=77).double()
        x0 = ref.data_utils.remove_mean(x0, n, 3)
        bad = 5
        x0[bad, 4] = float("inf")
        out = {"n": n, "N": N, "steps": steps, "x0": x0.numpy(), "bad_row": bad}

        def mk(**kw):
            args = dict(sde=None, num_integration_steps=10, start_resampling_step=0, end_resampling_step=10, lightning_module=None,
                        partial_annealing_factor_schedule=None, num_negative_time_steps=steps, post_mcmc_steps=steps,
                        dt_negative_time=2e-4)
            args.update(kw)
            return ref.integ.WeightedSDEIntegrator(**args)

        xv = torch.cat([x0[:bad], x0[bad + 1:]])
        torch.manual_seed(31)
        with _RecordDraws() as d:
            xp, lqf, lqb = ref.integ.mala_proposal(xv.clone(), tgt, 2e-4)
        out.update({"prop.x": xv.numpy(), "prop.dt": 2e-4, "prop.noise": d.randn[0].numpy(), "prop.x_prop": xp.numpy(),
                    "prop.log_q_fwd": lqf.numpy(), "prop.log_q_bwd": lqb.numpy()})

        for tag, adaptive in (("mala", False), ("mala_adaptive", True)):
            integ = mk()
            torch.manual_seed(32 + int(adaptive))
            with _RecordDraws() as d:
                if adaptive:
                    xo, rates = integ.metropolis_hastings_mala_adaptive(x0.clone(), tgt, dt_init=2e-4, return_acceptance_rate=True)
                else:
                    xo, rates = integ.metropolis_hastings_mala(x0.clone(), tgt, return_acceptance_rate=True)
            assert len(d.randn) == steps and len(d.rand) == steps
            out.update({tag + ".x": xo.numpy(), tag + ".rates": np.array(rates), tag + ".noise": torch.stack(d.randn).numpy(),
                        tag + ".uniform": torch.stack(d.rand).numpy()})
            # decision margins, recomputed from the recorded draws
            xc, dt = xv.clone(), 2e-4
            lp = tgt(xc)
            margin = 1e9
            for k in range(steps):
                _, g = tgt(xc, return_force=True)
                xpr = xc + 0.5 * dt * g + np.sqrt(dt) * d.randn[k]
                lpp, gp = tgt(xpr, return_force=True)
                lqf = -((xpr - (xc + 0.5 * dt * g)) ** 2).sum(1) / (2 * dt)
                lqb = -((xc - (xpr + 0.5 * dt * gp)) ** 2).sum(1) / (2 * dt)
                ratio = lpp - lp + lqb - lqf
                lu = torch.log(d.rand[k])
                margin = min(margin, float((lu - ratio).abs().min()))
                acc = (lu < ratio)
                xc = torch.where(acc[:, None], xpr, xc)
                lp = torch.where(acc, lpp, lp)
                xc = ref.data_utils.remove_mean(xc, n, 3)
                if adaptive:
                    dt = dt * 1.1 if float(acc.double().mean()) > 0.55 else dt / 1.1
            assert torch.allclose(xc, xo[:N - 1], rtol=0, atol=1e-12), "restated MALA loop differs from the reference"
            assert margin > 1e-3, margin
            print("wrote %s rates %s margin %.3g" % (tag, np.round(rates, 3), margin))
        # the reference's non-adaptive MALA with return_acceptance_rate=False is a no-op: `acceptance_rate` is unbound at the
        # print (:386), the NameError is swallowed by the except (:401) before the update lines run
        integ = mk()
        torch.manual_seed(40)
        xo, none = integ.metropolis_hastings_mala(x0.clone(), tgt, return_acceptance_rate=False)
        out["mala_norate.x"] = xo.numpy()
        assert none is None

        for tag, lang in (("descent", False), ("langevin", True)):
            integ = mk(do_langevin=lang)
            torch.manual_seed(50)
            with _RecordDraws() as d:
                xo = integ.negative_time_descent(xv.clone(), tgt)
            out[tag + ".x"] = xo.numpy()
            if lang:
                out[tag + ".noise"] = torch.stack(d.randn).numpy()
        np.savez_compressed(os.path.join(OUT, "post_n13.npz"), **out)
        print("wrote post_n13")
    finally:
        torch.set_default_dtype(torch.float32)


def golden_fk_variants():
    """VEReverseSDE.f branches the default fixtures do not reach: pin_energy=True (energy_net.py:41-48: the
    model energy is mixed with the clamped target energy by (1-t)^3) and precondition_beta=True on both wrappers
    (energy_net.py:38-39, score_net.py:37-38).  LJ-13, strong coordinate gain, fp64."""
    n, B, beta = 13, 5, 0.75
    sched = ref.noise.ElucidatingNoiseSchedule(sigma_min=0.05, sigma_max=80, rho=7)
    for tag, t, pin, pre in (("pin", 0.12, True, False), ("precond", 0.37, False, True)):
        net_e, net_s = make_net(n, 12345, True), make_net(n, 54321, True)
        gen = torch.Generator().manual_seed(4242 + int(pin))
        x32 = O.md_shaped_coords(B, n, seed=31) * (1.0 + float(sched.h(torch.tensor(t))) ** 0.5 * 0.3)
        x32 = ref.data_utils.remove_mean(x32 + (0.02 if pin else 0.3) * torch.randn(B, 3 * n, generator=gen), n, 3)
        out = {"n": n, "t": t, "beta": beta, "x": x32.double().numpy(), "sigma_min": 0.05, "gamma": 4.0 / 3.0, "pin": int(pin),
               "precondition_beta": int(pre)}
        out.update(sd_np(net_e, "E."))
        out.update(sd_np(net_s, "S."))
        torch.set_default_dtype(torch.float64)
        try:
            en = ref.energy_net.EnergyNet(net_e.double(), precondition_beta=pre)
            sn = ref.score_net.ScoreNet(net_s.double(), precondition_beta=pre)
            sde = ref.sdes.VEReverseSDE(sched, energy_net=en, score_net=sn, pin_energy=pin, debias_inference=True,
                                        cdf=lambda h, xx, b_, _f=sn.forward: ref.utils.compute_divergence_exact(_f, h, xx, b_))
            sde.trainer = FakeTrainer()
            gs = ref.anneal.ConstantAnnealingFactorSchedule(4.0 / 3.0)
            tgt = LJTarget(n, temperature=1.0)
            terms = sde.f(torch.tensor(t), x32.double().clone(), torch.tensor(beta), gs, 1.0, tgt, resampling_interval=1)
            for k in ("drift_X", "drift_A", "divergence_score", "cross_term", "dUt_dt"):
                out[k] = getattr(terms, k).detach().numpy()
            tt = torch.full((B,), t)
            xr = x32.double().clone().requires_grad_(True)
            out["U"] = en.forward_energy(sched.h(tt), xr, torch.tensor(beta), pin=pin, energy_function=tgt, t=tt).detach().numpy()
            out["gradU"] = en.forward(sched.h(tt), xr, torch.tensor(beta), pin=pin, energy_function=tgt, t=tt).detach().numpy()
            out["target_logp"] = tgt(x32.double()).numpy()
        finally:
            torch.set_default_dtype(torch.float32)
        np.savez_compressed(os.path.join(OUT, f"fk_n13_{tag}.npz"), **out)
        print("wrote fk_n13_%s drift_A %s U %s" % (tag, out["drift_A"][:3], out["U"][:3]))


def golden_loop_linear():
    """integrate_sde with a LinearAnnealingFactorSchedule (annealing_factor_schedules.py:34-70): gamma'(t) != 0, so the
    dgamma/dt * U term of the FK drift (sdes.py:227) is exercised end to end; resampling every step, no end resample."""
    n, N, S, chunk = 13, 16, 24, 8
    torch.set_default_dtype(torch.float64)
    try:
        net_e = make_net(n, 12345, True).double()
        net_s = make_net(n, 54321, True).double()
        sched = ref.noise.ElucidatingNoiseSchedule(sigma_min=0.05, sigma_max=80, rho=7)
        en, sn = ref.energy_net.EnergyNet(net_e), ref.score_net.ScoreNet(net_s)
        sde = ref.sdes.VEReverseSDE(sched, energy_net=en, score_net=sn, pin_energy=False, debias_inference=True,
                                    cdf=lambda h, xx, b_, _f=sn.forward: ref.utils.compute_divergence_exact(_f, h, xx, b_))
        sde.trainer = FakeTrainer()
        integ = ref.integ.WeightedSDEIntegrator(
            sde=sde, num_integration_steps=S, start_resampling_step=0, end_resampling_step=S, lightning_module=FakeLM(),
            partial_annealing_factor_schedule=None, resampling_interval=1, num_negative_time_steps=0, post_mcmc_steps=0,
            batch_size=chunk, resample_at_end=False, diffusion_scale=1.0)
        states = []
        inner = integ.ddp_batched_euler_maruyama_step

        def recording_step(t, x, a, dt, step, **kw):
            o = inner(t, x, a, dt, step, **kw)
            states.append((ref.data_utils.remove_mean(o[0], n, 3).detach().clone(), o[1].detach().clone()))
            return o

        integ.ddp_batched_euler_maruyama_step = recording_step
        tgt = LJTarget(n, temperature=1.0)
        torch.manual_seed(2025)
        gsch = ref.anneal.LinearAnnealingFactorSchedule(1.5, 1.0, t_start=0.9, t_end=0.2)
        scale = float((sched.h(torch.tensor(1.0)) / float(gsch.gamma(torch.tensor(1.0)))) ** 0.5)
        x1 = ref.prior.Prior(scale, n_particles=n, spatial_dim=3).sample(N)
        x1_in = x1.clone()
        x, logw, uniq, terms, acc = integ.integrate_sde(x1, tgt, gsch, inverse_temperature=torch.tensor(0.75))
        out = {"n": n, "N": N, "S": S, "chunk": chunk, "seed": 2025, "beta": 0.75, "start": 0, "end": S, "interval": 1,
               "gamma_args": np.array([1.5, 1.0, 0.9, 0.2]), "x1": x1_in.detach().numpy(), "x_final": x.detach().numpy(),
               "logweights": logw.detach().numpy(), "num_unique": np.array(uniq), "prior_scale": scale,
               "x_steps": torch.stack([s_[0] for s_ in states]).numpy().astype(np.float32),
               "a_steps": torch.stack([s_[1] for s_ in states]).numpy()}
        out.update(sd_np(net_e, "E."))
        out.update(sd_np(net_s, "S."))
        np.savez_compressed(os.path.join(OUT, "loop_n13_linear.npz"), **out)
        print("wrote loop_linear", uniq)
    finally:
        torch.set_default_dtype(torch.float32)


if __name__ == "__main__":
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["resample", "lj", "fk", "loop", "ad2", "laplacian", "schedules", "post", "variants", "loop_linear"]
    if "laplacian" in which:
        golden_laplacian()
    if "schedules" in which:
        golden_schedules()
    if "ad2" in which:
        golden_ad2()
    if "resample" in which:
        golden_resample()
    if "lj" in which:
        golden_lj()
    if "fk" in which:
        golden_egnn_and_fk()
    if "loop" in which:
        golden_loop()
    if "post" in which:
        golden_post()
    if "variants" in which:
        golden_fk_variants()
    if "loop_linear" in which:
        golden_loop_linear()
