"""EGNN kernels (forward, energy reverse pass, score + exact divergence) vs the reference's golden outputs
and the fp64 autograd oracle, through the C-ABI.  Tolerance: 1e-4 relative (north star), fp32 kernels."""
import numpy as np
import pytest
import torch

import pita_oracle as O
from helpers import assert_close, golden, make_net, state_from_golden

pytestmark = pytest.mark.gpu

FK_CASES = ["fk_n13_init.npz", "fk_n13_strong.npz", "fk_n55_init.npz", "fk_n55_strong.npz"]


@pytest.mark.parametrize("name", FK_CASES)
def test_forward_vs_reference_golden(name):
    g = golden(name)
    n = int(g["n"])
    net = make_net(n, state_from_golden(g, "S."))
    x = torch.from_numpy(g["x"]).float().cuda()
    out = net(torch.from_numpy(g["egnn_tcond"]).float().cuda(), x, torch.from_numpy(g["egnn_beta"]).float().cuda())
    assert_close(out, g["egnn_out_f64"], "EGNN_dynamics.forward vs reference fp64")


# divergence evaluation modes and their stated bounds: fp32 SIMT, 3xTF32 tensor cores (round-1 forward-mode kernel) and the
# bilinear engine (round 2, the default) meet the 1e-4 north-star tolerance; plain TF32 is the labelled looser path (stated bound on the divergence: 5e-2 relative, include/pita_b200.h;
# the score of that mode is still evaluated in 3xTF32 and held to 1e-4).
DIV_TOL = {"fp32": 1e-4, "3xtf32": 1e-4, "tf32": 5e-2, "bilinear": 1e-4}


@pytest.mark.parametrize("mode", ["bilinear", "fp32", "3xtf32", "tf32"])
@pytest.mark.parametrize("name", FK_CASES)
def test_energy_score_divergence_vs_reference_golden(name, mode):
    from pita_b200.energy_net import EnergyNet
    from pita_b200.noise_schedules import ElucidatingNoiseSchedule
    from pita_b200.score_net import ScoreNet
    g = golden(name)
    n, t, beta = int(g["n"]), float(g["t"]), float(g["beta"])
    en = EnergyNet(make_net(n, state_from_golden(g, "E.")))
    sn = ScoreNet(make_net(n, state_from_golden(g, "S.")), div_mode=mode)
    x = torch.from_numpy(g["x"]).float().cuda()
    sched = ElucidatingNoiseSchedule(float(g["sigma_min"]), 80.0, 7.0)
    B = x.shape[0]
    ht = torch.full((B,), float(sched.h(torch.tensor(t, dtype=torch.float64))), device="cuda")
    U, gU, dh = en._terms(ht, x, beta, True, True)
    assert_close(U, g["U"], "forward_energy")
    assert_close(gU, g["gradU"], "grad_x U")
    dh_dt = float(sched.dh_dt(torch.tensor(t, dtype=torch.float64)))
    assert_close(dh * dh_dt, g["dUt_dt"], "dU/dt")
    s, div = sn.score_and_divergence(ht, x, beta)
    assert_close(s, g["score"], "score")
    assert_close(div, g["div_score"], "div score (%s)" % mode, rtol=DIV_TOL[mode])
    assert_close(en.forward_energy(ht, x, beta), g["U"], "forward_energy (energy-only launch)")
    assert_close(sn.forward(ht, x, beta), g["score"], "score (no divergence launch)")


@pytest.mark.parametrize("name", FK_CASES)
def test_sde_f_terms_vs_reference_golden(name):
    """VEReverseSDE.f through the drop-in classes == reference SDETerms (sdes.py:130-239)."""
    from pita_b200.annealing_factor_schedules import ConstantAnnealingFactorSchedule
    from pita_b200.energy_net import EnergyNet
    from pita_b200.noise_schedules import ElucidatingNoiseSchedule
    from pita_b200.score_net import ScoreNet
    from pita_b200.sdes import VEReverseSDE
    g = golden(name)
    n, t, beta = int(g["n"]), float(g["t"]), float(g["beta"])
    sde = VEReverseSDE(ElucidatingNoiseSchedule(float(g["sigma_min"]), 80.0, 7.0),
                       energy_net=EnergyNet(make_net(n, state_from_golden(g, "E."))),
                       score_net=ScoreNet(make_net(n, state_from_golden(g, "S."))))
    x = torch.from_numpy(g["x"]).float().cuda()
    terms = sde.f(torch.tensor(t), x, torch.tensor(beta), ConstantAnnealingFactorSchedule(float(g["gamma"])), 1.0, None,
                  resampling_interval=1)
    assert_close(terms.drift_X, g["drift_X"], "drift_X")
    assert_close(terms.divergence_score, g["div_b"], "div_b")
    assert_close(terms.cross_term, g["cross"], "cross_term")
    assert_close(terms.dUt_dt, g["dUt_dt"], "dUt_dt")
    assert_close(terms.drift_A, g["drift_A"], "drift_A")
    sde.debias_inference = False
    nd = sde.f(torch.tensor(t), x, torch.tensor(beta), ConstantAnnealingFactorSchedule(float(g["gamma"])), 1.0, None)
    assert_close(nd.drift_X, g["drift_X_nodebias"], "drift_X (not debiased)")


@pytest.mark.parametrize("mode", ["bilinear", "fp32", "3xtf32", "tf32"])
@pytest.mark.parametrize("n,B", [(13, 37), (55, 5)])
@pytest.mark.parametrize("gain", [0.001, 0.3])
def test_kernels_vs_fp64_oracle_random(n, B, gain, mode):
    """Seeded random weights / MD-shaped inputs / per-particle noise levels vs the fp64 autograd oracle."""
    from pita_b200 import ops
    from pita_b200.egnn_temp_conditioned import pack_state_dict
    sdE = O.random_egnn_state(seed=100 + n, dtype=torch.float64, coord_gain=gain)
    sdS = O.random_egnn_state(seed=200 + n, dtype=torch.float64, coord_gain=gain)
    sched = O.EDMSchedule(0.05)
    t = torch.linspace(0.05, 0.98, B, dtype=torch.float64)
    ht = sched.h(t)
    x = O.md_shaped_coords(B, n, seed=n, dtype=torch.float64) * (1 + ht.sqrt()[:, None] * 0.5)
    x = O.centre(x, n)
    beta = 1.3
    tr, xr = t.clone().requires_grad_(True), x.clone().requires_grad_(True)
    E = O.model_energy(sdE, sched.h(tr), xr, beta, n)
    gx, gt = torch.autograd.grad(E.sum(), (xr, tr))
    s_ref = O.model_score(sdS, ht, x, beta, n)
    div_ref = O.exact_divergence(lambda h1, x1: O.model_score(sdS, h1, x1, beta, n), ht, x)
    wE, wS = pack_state_dict(sdE, 32, 3, "cuda"), pack_state_dict(sdS, 32, 3, "cuda")
    e, g, dh = ops.egnn_energy(wE, 32, 3, n, ht.float().cuda(), x.float().cuda(), beta)
    assert_close(e, E, "E")
    assert_close(g, gx, "grad E")
    # dE/dt is ill-conditioned at small t (cancellation between h^-3/2 U and h^-1/2 dU/dh): the reference's own fp32
    # evaluation is 1.1e-4 .. 8.1e-4 away from fp64 on exactly these inputs (measured with the oracle in fp32), so the
    # bar here is 5e-4; the golden-fixture tests above hold dU/dt to 1e-4.
    assert_close(dh.double().cpu() * sched.dh_dt(t), gt, "dE/dt", rtol=5e-4)
    s, d = ops.egnn_score_div(wS, 32, 3, n, ht.float().cuda(), x.float().cuda(), beta, mode=mode)
    # Score bar: 1e-4 norm-wise everywhere (measured <= 2.6e-6) and 1e-4 element-wise, except on the tensor-core (3xTF32) path
    # for the LJ-55 / strong-gain case, whose first particle (h = 0.0087) evaluates (c_s x + c_out F - x) / h with a
    # 1 / sqrt(h (1 + h)) = 10.7x amplification of every rounding in F: measured 1.0e-4 .. 1.3e-4 there (fp32 SIMT kernel
    # 0.9e-4, the reference's own fp32 arithmetic 2e-5; profiles/r1e_err_by_mode.jsonl).  north_star allows a stated looser
    # bound on a TF32 MLP path: 2.5e-4 element-wise for exactly this case.
    loose = mode != "fp32" and n == 55 and gain > 0.1
    assert_close(s, s_ref, "score", rtol=2.5e-4 if loose else 1e-4, norm_rtol=1e-4)
    assert_close(d, div_ref, "div (%s)" % mode, rtol=DIV_TOL[mode])


@pytest.mark.parametrize("mode", ["bilinear", "fp32", "3xtf32"])
@pytest.mark.parametrize("n", [13, 55])
def test_full_size_properties(n, mode):
    """Properties that hold at any size, checked at a size the oracle could not finish:
    determinism, permutation of the batch, rigid translation (score/energy-gradient are translation covariant
    only through the explicit x terms), batch-slice consistency with a small oracle-checked slice."""
    from pita_b200 import ops
    from pita_b200.egnn_temp_conditioned import pack_state_dict
    B = 2048 if n == 13 else 296
    if mode == "bilinear":
        B = 5500 if n == 13 else 1250  # more than two (phase A, phase B) launch pairs, ragged last batch
    sd = O.random_egnn_state(seed=7, dtype=torch.float64, coord_gain=0.3)
    w = pack_state_dict(sd, 32, 3, "cuda")
    sched = O.EDMSchedule(0.05)
    ht = torch.full((B,), float(sched.h(torch.tensor(0.4, dtype=torch.float64))))
    x = O.centre(O.md_shaped_coords(B, n, seed=3) * 1.5, n)
    s1, d1 = ops.egnn_score_div(w, 32, 3, n, ht.cuda(), x.cuda(), 1.0, mode=mode)
    s2, d2 = ops.egnn_score_div(w, 32, 3, n, ht.cuda(), x.cuda(), 1.0, mode=mode)
    assert torch.equal(s1, s2) and torch.equal(d1, d2), "not deterministic"
    perm = torch.randperm(B)
    s3, d3 = ops.egnn_score_div(w, 32, 3, n, ht.cuda(), x[perm].cuda(), 1.0, mode=mode)
    assert torch.equal(s3.cpu(), s1.cpu()[perm]) and torch.equal(d3.cpu(), d1.cpu()[perm]), "batch order dependence"
    e1, g1, h1 = ops.egnn_energy(w, 32, 3, n, ht.cuda(), x.cuda(), 1.0)
    e2, g2, h2 = ops.egnn_energy(w, 32, 3, n, ht.cuda(), x.cuda(), 1.0)
    assert torch.equal(e1, e2) and torch.equal(g1, g2) and torch.equal(h1, h2), "reverse pass not deterministic"
    k = 3
    div_ref = O.exact_divergence(lambda hh, xx: O.model_score(sd, hh, xx, 1.0, n), ht[:k].double(), x[:k].double())
    assert_close(d1[:k], div_ref, "div slice")


@pytest.mark.parametrize("precond", [False, True])
def test_denoisers_vs_oracle(precond):
    """ScoreNet.denoiser (score_net.py:21-43, with and without precondition_beta, return_score) and EnergyNet.denoiser /
    denoiser_and_energy (energy_net.py:64-79: x - h grad U, dU/dh, U) against the fp64 oracle (the inference half)."""
    from pita_b200.energy_net import EnergyNet
    from pita_b200.score_net import ScoreNet
    n, B = 13, 23
    sd = O.random_egnn_state(seed=5, dtype=torch.float64, coord_gain=0.3)
    net = make_net(n, sd)
    x = O.centre(O.md_shaped_coords(B, n, seed=9, dtype=torch.float64) * 1.3, n)
    ht = O.EDMSchedule(0.05).h(torch.linspace(0.3, 0.7, B, dtype=torch.float64))
    beta = 0.8
    s_ref = O.model_score(sd, ht, x, beta, n, precondition_beta=precond)
    d_ref = x + ht[:, None] * s_ref
    d_theta, score = ScoreNet(net, precondition_beta=precond).denoiser(ht.float().cuda(), x.float().cuda(), torch.tensor(beta, device="cuda"),
                                                                       return_score=True)
    assert_close(score, s_ref, "ScoreNet.denoiser score")
    assert_close(d_theta, d_ref, "ScoreNet.denoiser D_theta")
    hr, xr = ht.clone().requires_grad_(True), x.clone().requires_grad_(True)
    u = O.model_energy(sd, hr, xr, beta, n, precondition_beta=precond)
    gx, gh = torch.autograd.grad(u.sum(), (xr, hr))
    en = EnergyNet(net, precondition_beta=precond)
    den, du_dh, U = en.denoiser_and_energy(ht.float().cuda(), x.float().cuda(), torch.tensor(beta, device="cuda"))
    assert_close(U, u.detach(), "EnergyNet.denoiser_and_energy U")
    assert_close(du_dh, gh, "EnergyNet.denoiser_and_energy dU/dh", rtol=2e-4)
    assert_close(den, x - ht[:, None] * gx, "EnergyNet.denoiser_and_energy denoiser")
    assert_close(en.denoiser(ht.float().cuda(), x.float().cuda(), torch.tensor(beta, device="cuda")), x - ht[:, None] * gx, "EnergyNet.denoiser")


def test_bench_size_properties_lj55():
    """BASELINE.json's own size (LJ-55, 262 144 particles per GPU) through size-independent properties of the default engine:
    run-to-run determinism, independence of a particle's result from its position in the batch (first / last rows recomputed
    in a small batch, bit-exact), rotation invariance of the energy and of the divergence and rotation covariance of the score
    and of the energy gradient (E(3) equivariance of the EGNN), and a slice against the fp64 oracle."""
    from pita_b200 import ops
    from pita_b200.egnn_temp_conditioned import pack_state_dict
    n, B = 55, 1 << 18
    sd = O.random_egnn_state(seed=11, dtype=torch.float64, coord_gain=0.3)
    w = pack_state_dict(sd, 32, 3, "cuda")
    sched = O.EDMSchedule(0.05)
    gen = torch.Generator().manual_seed(0)
    base = O.centre(O.md_shaped_coords(4096, n, seed=5) * 1.5, n)
    x = (base.repeat(B // 4096, 1) * (1 + 0.05 * torch.rand(B, 1, generator=gen))).float().cuda()
    ht = torch.full((B,), float(sched.h(torch.tensor(0.45, dtype=torch.float64))), device="cuda")
    s1, d1 = ops.egnn_score_div(w, 32, 3, n, ht, x, 0.8)
    e1, g1, h1 = ops.egnn_energy(w, 32, 3, n, ht, x, 0.8)
    assert all(bool(torch.isfinite(v).all()) for v in (s1, d1, e1, g1, h1))
    s2, d2 = ops.egnn_score_div(w, 32, 3, n, ht, x, 0.8)
    assert torch.equal(s1, s2) and torch.equal(d1, d2), "not deterministic at bench size"
    for sl in (slice(0, 64), slice(B - 77, B)):
        ss, ds = ops.egnn_score_div(w, 32, 3, n, ht[sl], x[sl].contiguous(), 0.8)
        assert torch.equal(ss, s1[sl]) and torch.equal(ds, d1[sl]), "result depends on the position in the batch"
        es, gs, hs = ops.egnn_energy(w, 32, 3, n, ht[sl], x[sl].contiguous(), 0.8)
        assert torch.equal(es, e1[sl]) and torch.equal(gs, g1[sl]) and torch.equal(hs, h1[sl])
    # a proper rotation (about an arbitrary axis), applied to every atom
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=gen, dtype=torch.float64))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    R = q.float().cuda()
    sub = slice(0, B, 37)   # a strided subset is enough for the comparison, the kernels ran on everything above
    xr = (x[sub].reshape(-1, n, 3) @ R.T).reshape(-1, 3 * n).contiguous()
    sr, dr = ops.egnn_score_div(w, 32, 3, n, ht[sub], xr, 0.8)
    er, gr, hr = ops.egnn_energy(w, 32, 3, n, ht[sub], xr, 0.8)
    rot = lambda v: (v.reshape(-1, n, 3) @ R.T).reshape(-1, 3 * n)  # noqa: E731
    assert_close(dr, d1[sub], "divergence is rotation invariant", rtol=2e-4)
    assert_close(er, e1[sub], "energy is rotation invariant", rtol=2e-4)
    assert_close(hr, h1[sub], "dE/dh is rotation invariant", rtol=2e-4)
    assert_close(sr, rot(s1[sub]), "score is rotation covariant", rtol=2e-4)
    assert_close(gr, rot(g1[sub]), "grad E is rotation covariant", rtol=2e-4)
    k = 2
    div_ref = O.exact_divergence(lambda hh, xx: O.model_score(sd, hh, xx, 0.8, n), ht[:k].double().cpu(), x[:k].double().cpu())
    assert_close(d1[:k], div_ref, "div slice vs the fp64 oracle")


@pytest.mark.parametrize("n,B", [(13, 20), (55, 5)])
def test_bilinear_engine_tables_vs_oracle(n, B):
    """Every per-pair table the bilinear engine's phase A writes, its direct part and phase B's per-tile partial sums
    against oracle/egnn_bilinear.py (fp64): the intermediate quantities are pinned, not only the final divergence."""
    import tri_debug
    res = tri_debug.run(n, B, verbose=False)
    for k, v in res.items():
        assert v == v and v <= (1e-4 if k.startswith("phaseB") else 2e-5), "%s: %.3e" % (k, v)
