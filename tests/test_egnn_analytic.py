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


@pytest.mark.parametrize("n,B", [(13, 2), (7, 3)])
def test_cached_middle_layer_form_is_the_same_trace(n, B):
    """The form the score/divergence kernel evaluates (per-edge cache, du as a dot with v = Wc1^T (wc2 * silu'(zc)), sender
    sum before W3a) is algebraically identical to the plain tangent pass — and both match vmap(jacrev)."""
    sd = O.random_egnn_state(seed=9 + n, dtype=torch.float64, coord_gain=0.3)
    y = O.centre(O.md_shaped_coords(B, n, seed=n + 2, dtype=torch.float64) * 1.1, n)
    tc = torch.linspace(-0.3, 0.4, B, dtype=torch.float64)
    beta = torch.full((B,), 0.9, dtype=torch.float64)
    plain = A.trace_dxL_dy(sd, tc, y, beta, n)
    cached = A.trace_dxL_dy(sd, tc, y, beta, n, cached=True)
    assert _rel(cached, plain) < 1e-12
    from torch.func import jacrev, vmap

    def xl(t1, y1, b1):  # x_L = vel + y up to the mean removal, whose trace contribution is -3 (translation invariance)
        return O.egnn_velocity(sd, t1[None], y1[None], b1[None], n)[0]

    jac = vmap(jacrev(xl, argnums=1))(tc, y, beta)
    tr_vel = jac.diagonal(dim1=-2, dim2=-1).sum(-1)
    assert _rel(cached - 3 * n, tr_vel) < 1e-10  # tr d remove_mean(vel)/dy == tr d x_L/dy - 3n


@pytest.mark.parametrize("n,B", [(13, 2), (6, 3)])
def test_rank_structured_middle_layer(n, B):
    """the middle layer through the rank structure of its inputs (two dense products per generic edge,
    the k-independent third one cached, cf_a(i) factored out of the sender sum) gives the same trace; the function itself
    asserts the rank-one claim and the factored aggregate on the way."""
    sd = O.random_egnn_state(seed=19 + n, dtype=torch.float64, coord_gain=0.3)
    y = O.centre(O.md_shaped_coords(B, n, seed=n + 5, dtype=torch.float64) * 1.15, n)
    tc = torch.linspace(-0.2, 0.5, B, dtype=torch.float64)
    beta = torch.full((B,), 1.1, dtype=torch.float64)
    assert _rel(A.trace_dxL_dy_ranked(sd, tc, y, beta, n), A.trace_dxL_dy(sd, tc, y, beta, n)) < 1e-11


@pytest.mark.parametrize("n,B", [(13, 2), (5, 3)])
def test_bilinear_form_of_the_trace(n, B):
    """oracle/egnn_bilinear.py (the algebra of csrc/egnn_tri_*.cu): forward-mode tangent through layer 0, reverse-mode
    cotangent through layer 2, bilinear pairing on the middle layer's edges — first with unified per-k node tables, then in
    the kernel's shape (per-pair tables, generic / S / R items, one dense product per (edge, k)).  Both equal the tangent
    form, which test_cached_middle_layer_form_is_the_same_trace ties to vmap(jacrev)."""
    import egnn_bilinear as BL
    sd = O.random_egnn_state(seed=29 + n, dtype=torch.float64, coord_gain=0.3)
    y = O.centre(O.md_shaped_coords(B, n, seed=n + 7, dtype=torch.float64) * 1.15, n)
    tc = torch.linspace(-0.2, 0.5, B, dtype=torch.float64)
    beta = torch.full((B,), 1.1, dtype=torch.float64)
    ref = A.trace_dxL_dy(sd, tc, y, beta, n)
    assert _rel(BL.trace_bilinear_dense(sd, tc, y, beta, n), ref) < 1e-11
    assert _rel(BL.trace_bilinear_kernel_form(sd, tc, y, beta, n), ref) < 1e-11
    tab = BL.phase_a_tables(sd, tc, y, beta, n)
    per_recv = BL.phase_b_items(tab, n, per_receiver=True)
    assert _rel(tab["direct_node"].sum(1) + per_recv.sum(1), ref) < 1e-11


def test_bilinear_form_tolerates_tf32_operands():
    """The n^3 products of the bilinear form carry < 1e-3 of the trace: truncating their operand rows to TF32 (what the
    tensor core does to phase B's generic items) moves the LJ-55 'strong' fixture's trace by < 1e-5 relative."""
    import egnn_bilinear as BL
    from helpers import golden, state_from_golden
    g = golden("fk_n13_strong.npz")
    sd = state_from_golden(g, "S.")
    x = torch.from_numpy(g["x"]).double()[:2]
    sched = O.EDMSchedule(float(g["sigma_min"]))
    ht = sched.h(torch.full((2,), float(g["t"]), dtype=torch.float64))
    beta = torch.full((2,), float(g["beta"]), dtype=torch.float64)
    y = (1 + ht)[:, None] ** -0.5 * x

    def trunc(z):
        i = z.float().contiguous().view(torch.int32) & ~0x1FFF
        return i.view(torch.float32).double()

    exact = BL.trace_bilinear_kernel_form(sd, 0.125 * torch.log(ht), y, beta, 13)
    rough = BL.trace_bilinear_kernel_form(sd, 0.125 * torch.log(ht), y, beta, 13, rnd=trunc)
    assert _rel(rough, exact) < 1e-5
