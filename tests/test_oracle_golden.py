"""Pins the CPU oracle (oracle/pita_oracle.py) against outputs of the unmodified reference
(tests/golden/*.npz, written by oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

import pita_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _sd(g, prefix, dtype=torch.float64):
    return {k[len(prefix):]: torch.from_numpy(g[k]).to(dtype) for k in g.files if k.startswith(prefix)}


def _close(a, b, rtol, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    err = np.abs(a - b) / np.maximum(np.abs(b), 1.0)
    assert err.max() <= rtol, f"{what}: max rel err {err.max():.3e} > {rtol}"


FK_CASES = ["fk_n13_init.npz", "fk_n13_strong.npz", "fk_n55_init.npz", "fk_n55_strong.npz"]


@pytest.mark.parametrize("name", FK_CASES)
def test_egnn_forward_matches_reference(golden_dir, name):
    g = _load(golden_dir, name)
    n = int(g["n"])
    sd = _sd(g, "S.")
    x = torch.from_numpy(g["x"])
    out = O.egnn_velocity(sd, torch.from_numpy(g["egnn_tcond"]), x, torch.from_numpy(g["egnn_beta"]), n)
    _close(out.numpy(), g["egnn_out_f64"], 1e-11, "egnn fp64")
    # the last layer's node update is dead code for the velocity output
    out2 = O.egnn_velocity(sd, torch.from_numpy(g["egnn_tcond"]), x, torch.from_numpy(g["egnn_beta"]), n, skip_dead=True)
    assert torch.equal(out, out2)
    # fp32 evaluation of the oracle vs the reference as shipped (fp32): rounding-level agreement
    sd32 = _sd(g, "S.", torch.float32)
    out32 = O.egnn_velocity(sd32, torch.from_numpy(g["egnn_tcond"]).float(), x.float(),
                            torch.from_numpy(g["egnn_beta"]).float(), n)
    scale = np.abs(g["egnn_out_f64"]).max()
    assert np.abs(out32.numpy() - g["egnn_out_f32"]).max() <= 2e-5 * max(scale, 1e-3) + 1e-7


@pytest.mark.parametrize("name", FK_CASES)
def test_fk_terms_match_reference(golden_dir, name):
    g = _load(golden_dir, name)
    n = int(g["n"])
    sdE, sdS = _sd(g, "E."), _sd(g, "S.")
    x = torch.from_numpy(g["x"])
    sched = O.EDMSchedule(float(g["sigma_min"]))
    gam = O.ConstGamma(float(g["gamma"]))
    d = O.fk_drift(sdE, sdS, sched, gam, float(g["t"]), x, float(g["beta"]), n)
    _close(d.energy, g["U"], 1e-10, "U")
    _close(d.grad_u, g["gradU"], 1e-9, "gradU")
    _close(d.score, g["score"], 1e-9, "score")
    _close(d.drift_x, g["drift_X"], 1e-9, "drift_X")
    _close(d.div_b, g["div_b"], 1e-9, "div_b")
    _close(d.cross, g["cross"], 1e-9, "cross")
    _close(d.du_dt, g["dUt_dt"], 1e-9, "dUt_dt")
    _close(d.drift_a, g["drift_A"], 1e-9, "drift_A")
    tt = torch.full((x.shape[0],), float(g["t"]), dtype=torch.float64)
    div = O.exact_divergence(lambda h1, x1: O.model_score(sdS, h1, x1, float(g["beta"]), n), sched.h(tt), x)
    _close(div, g["div_score"], 1e-9, "div_score")
    nd = O.fk_drift(sdE, sdS, sched, gam, float(g["t"]), x, float(g["beta"]), n, debias=False)
    _close(nd.drift_x, g["drift_X_nodebias"], 1e-9, "drift_X (not debiased)")


def test_resampler_matches_reference(golden_dir):
    g = _load(golden_dir, "resample.npz")
    for N in g["case_sizes"]:
        N = int(N)
        assert int(g[f"cumsum_claim_{N}"]) == 1
        ids = O.systematic_indices(g[f"weights_{N}"], float(g[f"u0_{N}"]))
        if f"ids_{N}" in g.files:
            assert np.array_equal(ids, g[f"ids_{N}"].astype(np.int64))
            # and from the logits (softmax restated by torch on this CPU)
            ids2 = O.systematic_resample(torch.from_numpy(g[f"logits_{N}"]), float(g[f"u0_{N}"]))
            assert (ids2 != ids).mean() < 1e-3
        else:
            assert np.array_equal(ids[::64], g[f"ids_{N}_stride64"].astype(np.int64))
            assert int(ids.sum()) == int(g[f"ids_{N}_sum"])
            assert len(np.unique(ids)) == int(g[f"ids_{N}_unique"])
    w = O.clipped_softmax(torch.from_numpy(g["edge_logits"])).numpy()
    for name, u0 in (("zero", 0.0), ("almost1", 1.0 - 2.0 ** -53), ("half", 0.5)):
        assert np.array_equal(O.systematic_indices(w, u0), g[f"edge_{name}_ids"].astype(np.int64)), name


def test_lj_matches_reference(golden_dir):
    g = _load(golden_dir, "lj.npz")
    for n in (13, 55):
        x = torch.from_numpy(g[f"x_{n}"])
        for T in (1.0, 2.5):
            lp, f = O.lj_logprob_force(x.double(), n, temperature=T)
            _close(lp, g[f"logp_f64_{n}_T{T}"], 1e-12, "lj logp")
            _close(f, g[f"force_f64_{n}_T{T}"], 1e-11, "lj force")
            lp32, f32 = O.lj_logprob_force(x, n, temperature=T)
            _close(lp32, g[f"logp_f32_{n}_T{T}"], 2e-5, "lj logp fp32")
        # eps-free in-repo restatement (sampling/sample_lj13.py:24-30) agrees up to the bgflow eps
        lp0 = -O.lj_energy(x.double(), n, eps_sqrt=0.0)
        _close(lp0, g[f"energy2_neg_{n}"], 1e-11, "energy2")


def test_loop_matches_reference(golden_dir):
    g = _load(golden_dir, "loop_n13.npz")
    n, N, S, chunk = int(g["n"]), int(g["N"]), int(g["S"]), int(g["chunk"])
    sdE, sdS = _sd(g, "E."), _sd(g, "S.")
    sched = O.EDMSchedule(0.05)
    gam = O.ConstGamma(float(g["gamma"]))
    torch.manual_seed(int(g["seed"]))
    x1 = O.mean_free_prior(N, n, float(g["prior_scale"]), dtype=torch.float64)
    _close(x1, g["x1"], 1e-13, "prior")
    cfg = O.LoopConfig(n=n, steps=S, chunk=chunk, beta=float(g["beta"]), resampling_interval=int(g["interval"]),
                       start_resampling_step=int(g["start"]), end_resampling_step=int(g["end"]), resample_at_end=True)
    x, logw, uniq = O.integrate(
        sdE, sdS, sched, gam, cfg, x1,
        noise_fn=lambda step, xc: torch.randn_like(xc),
        u0_fn=lambda step: float(torch.rand(size=(1,), dtype=torch.float64)))
    assert list(uniq) == list(g["num_unique"])
    _close(logw, g["logweights"], 1e-8, "logweights")
    _close(x, g["x_final"], 1e-8, "x_final")


def test_ad2_egnn_oracle_vs_reference_golden(golden_dir):
    """SURVEY §8 row a8' (alanine dipeptide EGNN: 22 atoms, hidden 64, 5 layers, 21 one-hot node-type columns ++ t ++ beta): the oracle
    restatement reproduces the unmodified reference's forward (fixture generated by oracle/make_golden.py ad2) in fp64.
    The CUDA path for this network is not built yet — this pins the oracle for it."""
    g = _load(golden_dir, "egnn_ad2_n22.npz")
    sd = _sd(g, "W.")
    n = int(g["n"])
    assert O.egnn_layer_count(sd) == 5 and sd["egnn.embedding.weight"].shape == (64, 23)
    vel = O.egnn_velocity_ad2(sd, torch.from_numpy(g["t"]), torch.from_numpy(g["x"]), torch.from_numpy(g["beta"]), n)
    ref = torch.from_numpy(g["vel"])
    assert (vel - ref).abs().max().item() <= 1e-12 * max(1.0, ref.abs().max().item())
    # the node-type table (egnn_dynamics_ad2_cat.py:67-73): three hydrogen triples share a class; the largest label is 20,
    # so one_hot has 21 columns and in_node_nf = 21 + t + beta = 23
    types = O.ad2_atom_types(22)
    assert types.max().item() == 20 and len(set(types.tolist())) == 16


def test_laplacian_branch_oracle_vs_reference_golden(golden_dir):
    """VEReverseSDE.f without a score net (b = -grad U g^2/2, div b = -laplacian(U) g^2/2 through
    compute_laplacian_exact, sdes.py:150-153, 204-216).  Oracle vs the unmodified reference in fp64 (fixture:
    oracle/make_golden.py laplacian).  Pins the oracle of csrc/egnn_lap.cu (tests/test_gpu_laplacian.py)."""
    g = _load(golden_dir, "fk_n13_laplacian.npz")
    n = int(g["n"])
    d = O.fk_drift(_sd(g, "E."), None, O.EDMSchedule(float(g["sigma_min"])), O.ConstGamma(float(g["gamma"])), float(g["t"]),
                   torch.from_numpy(g["x"]), float(g["beta"]), n)
    _close(d.div_b, g["div_b"], 1e-9, "div_b (laplacian)")
    _close(d.cross, g["cross"], 1e-9, "cross term")
    _close(d.drift_x, g["drift_X"], 1e-9, "drift_X")
    _close(d.drift_a, g["drift_A"], 1e-9, "drift_A")
