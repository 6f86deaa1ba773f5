"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-4  # north-star tolerance for fp32 paths: |d| <= 1e-4 * max(|ref|, 1) and ||d||/||ref|| <= 1e-4


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def state_from_golden(g, prefix, dtype=torch.float64):
    return {k[len(prefix):]: torch.from_numpy(g[k]).to(dtype) for k in g.files if k.startswith(prefix)}


def assert_close(got, ref, what, rtol=RTOL, norm_rtol=None):
    got = got.detach().double().cpu().numpy() if torch.is_tensor(got) else np.asarray(got, dtype=np.float64)
    ref = ref.detach().double().cpu().numpy() if torch.is_tensor(ref) else np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    assert np.isfinite(got).all(), f"{what}: non-finite values"
    err = np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)
    assert err.max() <= rtol, f"{what}: max elementwise rel err {err.max():.3e} > {rtol:g}"
    nr = np.linalg.norm(ref)
    if nr > 0:
        nerr = np.linalg.norm(got - ref) / nr
        lim = norm_rtol if norm_rtol is not None else rtol
        assert nerr <= lim, f"{what}: norm-wise rel err {nerr:.3e} > {lim:g}"


def make_net(n, sd, device="cuda"):
    """pita_b200 EGNN_dynamics carrying the given (reference-named) weights."""
    from pita_b200.egnn_temp_conditioned import EGNN_dynamics

    net = EGNN_dynamics(n_particles=n, n_dimension=3, hidden_nf=32, n_layers=3, act_fn=torch.nn.SiLU(), recurrent=True,
                        tanh=True, attention=True, condition_time=True, condition_temperature=True, agg="sum")
    net.load_state_dict({k: v.float() for k, v in sd.items()})
    return net.to(device)
