"""Host-side mirrors of the reference interface (pita_b200/*.py) — everything that runs on the host and needs no GPU:
noise / annealing-factor schedules (folded into kernel arguments as fp64 host scalars), the EGNN parameter container and its
packed-weight layout, SDETerms, shard bookkeeping, and the rule that the product never computes on the CPU.
Reference values: tests/golden/schedules.npz and the fk_* fixtures, written by oracle/make_golden.py from the unmodified
reference.  CPU only."""
import copy
import os

import numpy as np
import pytest
import torch

import pita_oracle as O


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _cmp(got, ref, what, rtol=1e-12):
    got = np.asarray(torch.as_tensor(got, dtype=torch.float64))
    ref = np.asarray(ref, dtype=np.float64)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin), what + ": finiteness differs"
    err = np.abs(got[fin] - ref[fin]) / np.maximum(np.abs(ref[fin]), 1e-300)
    assert err.max() <= rtol, f"{what}: max rel err {err.max():.3e}"


def test_noise_schedules_match_reference(golden_dir):
    """noise_schedules.py:19-28, 64-125 on a grid of 41 times (fp64)."""
    from pita_b200 import noise_schedules as ns
    g = _g(golden_dir, "schedules.npz")
    t = torch.from_numpy(g["t"])
    cases = {"edm005": ns.ElucidatingNoiseSchedule(0.05, 80, 7), "edm001": ns.ElucidatingNoiseSchedule(0.01, 80, 7),
             "geo": ns.GeometricNoiseSchedule(0.01, 3.0), "lin": ns.LinearNoiseSchedule(2.5)}
    for tag, sch in cases.items():
        _cmp(sch.h(t), g[tag + ".h"], tag + ".h")
        _cmp(sch.g(t), g[tag + ".g"], tag + ".g")
        if tag + ".dh_dt" in g.files:
            _cmp(sch.dh_dt(t), g[tag + ".dh_dt"], tag + ".dh_dt")
        if tag + ".t_of_h" in g.files:
            _cmp(sch.t(sch.h(t)), g[tag + ".t_of_h"], tag + ".t(h)", rtol=1e-9)
    # the oracle's restatement (used by every loop test) against the same reference values
    for tag, smin in (("edm005", 0.05), ("edm001", 0.01)):
        o = O.EDMSchedule(smin)
        _cmp(o.h(t), g[tag + ".h"], "oracle " + tag + ".h")
        _cmp(o.g(t), g[tag + ".g"], "oracle " + tag + ".g")
        _cmp(o.dh_dt(t), g[tag + ".dh_dt"], "oracle " + tag + ".dh_dt")
    # g^2 == dh/dt (what the fused step kernel assumes when it is handed g2 and dh_dt separately)
    e = cases["edm005"]
    _cmp(e.g(t) ** 2, np.asarray(e.dh_dt(t)), "g^2 = dh/dt", rtol=1e-12)


def test_annealing_factor_schedules_match_reference(golden_dir):
    """annealing_factor_schedules.py:20-109."""
    from pita_b200 import annealing_factor_schedules as af
    g = _g(golden_dir, "schedules.npz")
    t = torch.from_numpy(g["t"])
    cases = {"const": (af.ConstantAnnealingFactorSchedule(4.0 / 3.0), O.ConstGamma(4.0 / 3.0)),
             "linear": (af.LinearAnnealingFactorSchedule(1.5, 1.0, t_start=0.8, t_end=0.2), O.LinearGamma(1.5, 1.0, 0.8, 0.2)),
             "sigmoid": (af.SigmoidAnnealingFactorSchedule(1.5, 1.0, t_start=0.9, t_end=0.1, sharpness=8.0),
                         O.SigmoidGamma(1.5, 1.0, 0.9, 0.1, 8.0))}
    for tag, (mirror, oracle) in cases.items():
        for obj, who in ((mirror, "pita_b200"), (oracle, "oracle")):
            _cmp(obj.gamma(t), g[tag + ".gamma"], f"{who} {tag}.gamma")
            _cmp(obj.dgamma_dt(t), g[tag + ".dgamma_dt"], f"{who} {tag}.dgamma_dt")
        # python floats are accepted like tensors (the reference wraps them in a float32 torch.tensor, :25-26)
        assert float(mirror.gamma(0.5)) == pytest.approx(float(g[tag + ".gamma"][20]), rel=1e-6)


def test_egnn_container_loads_reference_state_dict_and_packs(golden_dir):
    """Parameter names / shapes are the reference's, so its checkpoints load; the packed device layout
    (csrc/egnn_common.cuh::pk) has the size the C ABI reports and is refreshed when a parameter changes in place."""
    from pita_b200 import _native
    from pita_b200.egnn_temp_conditioned import EGNN_dynamics, pack_state_dict
    g = _g(golden_dir, "fk_n13_init.npz")
    sd = {k[2:]: torch.from_numpy(g[k]).float() for k in g.files if k.startswith("S.")}
    net = EGNN_dynamics(n_particles=13, n_dimension=3, hidden_nf=32, n_layers=3, act_fn=torch.nn.SiLU(), recurrent=True,
                        tanh=True, attention=True, condition_time=True, condition_temperature=True, agg="sum")
    assert set(net.state_dict().keys()) == set(sd.keys())
    assert all(net.state_dict()[k].shape == v.shape for k, v in sd.items())
    net.load_state_dict(sd)  # strict
    assert sum(p.numel() for p in net.parameters()) == 22533  # LJ config
    pack = pack_state_dict(net.state_dict(), 32, 3, "cpu")
    assert pack.dtype == torch.float32 and pack.numel() == _native.load().pita_egnn_pack_floats(32, 3)
    # header of the pack = the node embedding as the kernels read it: column 0, column 1 of Linear(2 -> 32), then its bias
    w, b = sd["egnn.embedding.weight"], sd["egnn.embedding.bias"]
    assert torch.equal(pack[:32], w[:, 0]) and torch.equal(pack[32:64], w[:, 1]) and torch.equal(pack[64:96], b)
    # deepcopy-able (energytemp_module.py:99) and the pack cache notices in-place updates (EMA swap, optimizer step)
    twin = copy.deepcopy(net)
    p0 = net.packed_weights("cpu")
    assert net.packed_weights("cpu") is p0
    with torch.no_grad():
        next(net.parameters()).add_(1.0)
    assert net.packed_weights("cpu") is not p0
    assert torch.equal(twin.packed_weights("cpu"), p0)
    # configurations the native path does not build are refused loudly, not approximated
    with pytest.raises(NotImplementedError):
        EGNN_dynamics(n_particles=22, n_dimension=3, hidden_nf=64, n_layers=5, recurrent=True, tanh=True, attention=True,
                      condition_time=True, condition_temperature=True)


def test_sde_terms_container():
    """SDETerms (sdes.py:34-92): optional fields survive .cpu() and .concatenate()."""
    from pita_b200.sdes import SDETerms
    a = SDETerms(drift_X=torch.ones(2, 3), drift_A=torch.zeros(2), divergence_score=torch.arange(2.0))
    b = SDETerms(drift_X=2 * torch.ones(4, 3), drift_A=torch.ones(4), divergence_score=torch.arange(4.0))
    c = SDETerms.concatenate([SDETerms.cpu(a), b])
    assert c.drift_X.shape == (6, 3) and c.drift_A.shape == (6,) and c.divergence_score.shape == (6,)
    assert c.cross_term is None and c.dUt_dt is None and c.diffusion is None
    with pytest.raises(ValueError):
        SDETerms.concatenate([])


def test_product_refuses_cpu_tensors():
    """No CPU fallback anywhere on the path: the public entry points raise on host tensors instead of computing."""
    from pita_b200.lennardjones_energy import LennardJonesEnergy
    from pita_b200.sde_integration import WeightedSDEIntegrator
    from pita_b200.utils import num_unique_from_changes, sample_cat_sys
    tgt = LennardJonesEnergy(dimensionality=39, n_particles=13)
    with pytest.raises(Exception):
        tgt(torch.zeros(4, 39))
    with pytest.raises(Exception):
        sample_cat_sys(8, torch.zeros(8))
    integ = WeightedSDEIntegrator(sde=None, num_integration_steps=2, start_resampling_step=0, end_resampling_step=2)
    with pytest.raises(RuntimeError):
        integ.integrate_sde(torch.zeros(4, 39), tgt, None)
    # len(np.unique(ids)) of a wrapped systematic resample == number of cyclic index changes, 0 changes meaning 1 ancestor
    assert num_unique_from_changes(0) == 1 and num_unique_from_changes(7) == 7
    # resampling schedule predicate (sde_integration.py:292): (step+1) % interval == 0 inside [start, end)
    integ.start_resampling_step, integ.end_resampling_step = 2, 6
    assert [s for s in range(8) if integ._will_resample(s, 2)] == [3, 5]
    assert not any(integ._will_resample(s, -1) for s in range(8))


def test_generate_samples_call_contract():
    """energytemp_module.py:237-298: prior scale sqrt(h(t_start) / gamma(t_start)); first pass on all prior samples; with
    return_logweights a second pass on the first `inference_batch_size` prior samples with resampling switched off by
    resampling_interval = num_integration_steps + 1; the reference's return tuples.  Host orchestration only — the
    integrator and the prior are stand-ins that record how they are called."""
    from functools import partial

    from pita_b200.annealing_factor_schedules import ConstantAnnealingFactorSchedule
    from pita_b200.noise_schedules import ElucidatingNoiseSchedule
    from pita_b200.sampling import generate_samples, prior_scale

    calls = []

    class FakeIntegrator:
        def integrate_sde(self, x1, energy_function, annealing_factor_schedule, inverse_temperature=1.0,
                          annealing_factor_score=1.0, resampling_interval=None):
            calls.append(dict(n=x1.shape[0], interval=resampling_interval, beta=inverse_temperature,
                              gamma=float(annealing_factor_schedule.gamma(0.3)), score=annealing_factor_score, x1=x1))
            return x1 + 1.0, torch.zeros(5, x1.shape[0]), [x1.shape[0]] * 5, ["terms"], [0.5]

    class FakePrior:
        def __init__(self, scale, device):
            self.scale, self.device = scale, device

        def sample(self, n):
            return torch.arange(n * 6, dtype=torch.float32).reshape(n, 6) * float(self.scale)

    sched = ElucidatingNoiseSchedule(0.05, 80, 7)
    kw = dict(weighted_sde_integrator=FakeIntegrator(), energy_function="target", num_samples=10, noise_schedule=sched,
              annealing_factor_schedule=partial(ConstantAnnealingFactorSchedule), partial_prior=FakePrior, t_start=torch.tensor(1.0),
              device="cpu", inference_batch_size=4, num_integration_steps=100, inverse_temp=0.75, annealing_factor=4.0 / 3.0)
    out = generate_samples(**kw)
    assert len(out) == 4 and len(calls) == 1
    assert calls[0]["n"] == 10 and calls[0]["interval"] is None and calls[0]["beta"] == 0.75
    assert calls[0]["gamma"] == pytest.approx(4.0 / 3.0) and calls[0]["score"] == pytest.approx(4.0 / 3.0)  # score factor defaults to it
    scale = float(prior_scale(sched, ConstantAnnealingFactorSchedule(4.0 / 3.0), torch.tensor(1.0)))
    assert scale == pytest.approx((80.0 ** 2 / (4.0 / 3.0)) ** 0.5, rel=1e-5)  # h(1) = sigma_max^2
    assert torch.equal(out[0], calls[0]["x1"] + 1.0) and out[1] == [10] * 5 and out[2] == ["terms"] and out[3] == [0.5]
    calls.clear()
    out = generate_samples(return_logweights=True, annealing_factor_score=1.1, **kw)
    assert len(out) == 6 and len(calls) == 2
    assert calls[0]["n"] == 10 and calls[0]["interval"] is None and calls[0]["score"] == 1.1
    assert calls[1]["n"] == 4 and calls[1]["interval"] == 101  # no resampling in the log-weight pass
    assert torch.equal(calls[1]["x1"], calls[0]["x1"][:4])     # the SAME prior samples, first inference_batch_size of them
    samples, not_resampled, logw, uniq, terms, acc = out
    assert samples.shape == (10, 6) and not_resampled.shape == (4, 6) and logw.shape == (5, 4) and uniq == [10] * 5


def test_ad2_class_state_dict_matches_the_reference(golden_dir):
    """EGNN_dynamics_AD2_cat (egnn_dynamics_ad2_cat.py:11-65) as a drop-in: same parameter names and shapes as the reference
    module (read from the fixture the unmodified reference wrote), so reference checkpoints load; unsupported configurations
    raise instead of falling back; the weight pack has the size the C ABI states."""
    import numpy as np

    from pita_b200 import ops
    from pita_b200.egnn_dynamics_ad2_cat import EGNN_dynamics_AD2_cat, atom_types_ad2, pack_state_dict_ad2
    g = np.load(os.path.join(golden_dir, "egnn_ad2_n22.npz"))
    ref = {k[2:]: g[k].shape for k in g.files if k.startswith("W.")}
    net = EGNN_dynamics_AD2_cat(n_particles=22, n_dimensions=3, hidden_nf=64, n_layers=5, act_fn=torch.nn.SiLU(), recurrent=True,
                                attention=True, tanh=True, agg="sum", condition_beta=True)
    mine = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert mine == {k: tuple(s) for k, s in ref.items()}
    assert list(net.state_dict().keys()) == [k[2:] for k in g.files if k.startswith("W.")]  # creation order too
    net.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("W.")})
    t = atom_types_ad2(22)
    assert t.max().item() == 20 and torch.nn.functional.one_hot(t).shape == (22, 21)
    for bad in (dict(n_particles=21), dict(hidden_nf=32), dict(n_layers=4), dict(condition_beta=False), dict(agg="mean")):
        kw = dict(n_particles=22, n_dimensions=3, hidden_nf=64, n_layers=5, condition_beta=True)
        kw.update(bad)
        with pytest.raises(NotImplementedError):
            EGNN_dynamics_AD2_cat(**kw)
    w = pack_state_dict_ad2(net.state_dict(), 64, 5, "cpu")
    assert w.numel() == ops.egnn_pack_floats(64, 5) and w.dtype == torch.float32
    # spot checks of the layout (csrc/egnn_ad2.cu, namespace ad2::pk): embedding stored [feature][64]; first block = A^T of layer 0
    emb = net.state_dict()["egnn.embedding.weight"]
    assert torch.equal(w[: 23 * 64].reshape(23, 64), emb.t())
    A = net.state_dict()["egnn.gcl_0.edge_mlp.0.weight"][:, :64]
    off = 23 * 64 + 64
    assert torch.equal(w[off: off + 4096].reshape(64, 64), A.t())
