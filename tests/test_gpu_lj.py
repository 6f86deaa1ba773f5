"""LJ energy+force kernel vs the oracle / golden fixtures (through the C-ABI)."""
import numpy as np
import pytest
import torch

import pita_oracle as O
from helpers import assert_close, golden

pytestmark = pytest.mark.gpu


def _run(x, n, T=1.0):
    from pita_b200.lennardjones_energy import LennardJonesEnergy
    e = LennardJonesEnergy(dimensionality=3 * n, n_particles=n, spatial_dim=3, temperature=T)
    return e(x.cuda(), return_force=True)


@pytest.mark.parametrize("n", [13, 55])
@pytest.mark.parametrize("T", [1.0, 2.5])
def test_lj_golden(n, T):
    g = golden("lj.npz")
    x = torch.from_numpy(g[f"x_{n}"])
    lp, f = _run(x, n, T)
    assert_close(lp, g[f"logp_f64_{n}_T{T}"], "logp vs reference fp64")
    assert_close(f, g[f"force_f64_{n}_T{T}"], "force vs reference fp64")


@pytest.mark.parametrize("n,B", [(13, 1), (13, 33), (13, 4097), (55, 1), (55, 7), (55, 1000)])
def test_lj_vs_oracle_ragged_batches(n, B):
    x = O.md_shaped_coords(B, n, seed=B + n)
    lp_ref, f_ref = O.lj_logprob_force(x.double(), n, temperature=1.7)
    lp, f = _run(x, n, 1.7)
    assert_close(lp, lp_ref, "logp")
    assert_close(f, f_ref, "force")


@pytest.mark.parametrize("n", [13, 55])
def test_lj_energy_only_and_empty(n):
    from pita_b200 import ops
    x = O.md_shaped_coords(10, n, seed=1).cuda()
    lp, f = ops.lj_energy_force(x, n, need_force=False)
    assert f is None
    lp2, _ = ops.lj_energy_force(x, n)
    assert torch.equal(lp, lp2)
    e, f0 = ops.lj_energy_force(x[:0], n)
    assert e.numel() == 0 and f0.numel() == 0


def test_lj_unsupported_n_raises():
    from pita_b200 import ops
    from pita_b200.lennardjones_energy import LennardJonesEnergy
    with pytest.raises(NotImplementedError):
        LennardJonesEnergy(dimensionality=66, n_particles=22)
    with pytest.raises(RuntimeError):
        ops.lj_energy_force(torch.zeros(2, 66, device="cuda"), 22)


@pytest.mark.parametrize("n", [13, 55])
def test_lj_full_size_properties(n):
    """Size-independent properties at benchmark scale: translation invariance, zero net force from the pair term
    (sum_i f_i = -(1/T) sum_i (x_i - com) = 0), determinism."""
    B = 1 << 18
    x = O.md_shaped_coords(B, n, seed=5).cuda()
    lp, f = _run(x, n)
    lp2, f2 = _run(x, n)
    assert torch.equal(lp, lp2) and torch.equal(f, f2)
    net = f.reshape(B, n, 3).sum(1)
    assert net.abs().max().item() < 5e-3 * max(1.0, f.abs().max().item() * 1e-3)
    shift = torch.tensor([0.25, -0.5, 0.125], device="cuda").repeat(n)
    lp3, f3 = _run(x + shift, n)
    assert_close(lp3, lp, "translation invariance (logp)", rtol=2e-3)  # fp32: the shift changes every rounding
    # spot check against the fp64 oracle on a slice
    lp_ref, f_ref = O.lj_logprob_force(x[:512].cpu().double(), n)
    assert_close(lp[:512], lp_ref, "logp slice")
    assert_close(f[:512], f_ref, "force slice")
