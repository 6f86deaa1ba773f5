"""Drop-in for `src.models.components.egnn_dynamics_ad2_cat.EGNN_dynamics_AD2_cat` (reference :11-218) over
`src.models.components.egnn.EGNN` (reference egnn.py:108-184): the alanine-dipeptide EGNN denoiser of
configs/model/net/egnn_dynamics_ad2_cat.yaml (22 atoms, hidden 64, 5 layers, SiLU, recurrent, tanh, attention, agg = sum,
node features = one_hot(atom type) ++ t [++ beta]).

Same constructor kwargs, same parameter names / creation order (state_dicts and seeded initialisation are interchangeable),
`forward(t, xs, beta) -> [B, n*d]`.  The compute is `pita_egnn_forward` / `pita_egnn_energy` / `pita_egnn_score_div` with
(hidden, layers, n) = (64, 5, 22): csrc/egnn_ad2.cu.  Anything else raises — there is no PyTorch fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .egnn_temp_conditioned import _EGNNParams


def atom_types_ad2(n_particles: int = 22) -> torch.Tensor:
    """reference :67-73 — one class per atom except the three hydrogen triples that share one."""
    if n_particles != 22:
        raise NotImplementedError("only the 22-atom alanine dipeptide typing is built")
    t = torch.arange(22)
    t[[0, 2, 3]] = 2
    t[[19, 20, 21]] = 20
    t[[11, 12, 13]] = 12
    return t


def pack_state_dict_ad2(sd, hidden: int, layers: int, device) -> torch.Tensor:
    """Flat fp32 weight buffer in the layout csrc/egnn_ad2.cu (namespace ad2::pk) expects: embedding weight as
    [feature][H] (21 one-hot columns, t, beta), bias, then per layer the fourteen H x H blocks (`*_f` = W^T, `*_b` = W) and ten
    H-vectors — the same order as the LJ pack (egnn_temp_conditioned.pack_state_dict)."""
    H = hidden
    f = lambda k: sd[k].detach().to(dtype=torch.float32, device="cpu")  # noqa: E731
    emb = f("egnn.embedding.weight")  # [H][23]
    if emb.shape != (H, 23):
        raise NotImplementedError("EGNN_dynamics_AD2_cat is built for condition_beta=True (23 node features); got %r" % (tuple(emb.shape),))
    parts = [emb.t().contiguous().reshape(-1), f("egnn.embedding.bias")]
    for l in range(layers):
        pre = "egnn.gcl_%d." % l
        W1 = f(pre + "edge_mlp.0.weight")
        A, Bm = W1[:, :H], W1[:, H:2 * H]
        W2, Wc1 = f(pre + "edge_mlp.2.weight"), f(pre + "coord_mlp.0.weight")
        W3 = f(pre + "node_mlp.0.weight")
        W3h, W3a, W4 = W3[:, :H], W3[:, H:], f(pre + "node_mlp.2.weight")
        for M in (A, Bm):
            parts.append(M.t().contiguous().reshape(-1))
        for M in (A, Bm):
            parts.append(M.contiguous().reshape(-1))
        for M in (W2, Wc1, W3h, W3a, W4):
            parts.append(M.t().contiguous().reshape(-1))
            parts.append(M.contiguous().reshape(-1))
        ba = torch.zeros(H)
        ba[0] = f(pre + "att_mlp.0.bias")[0]
        parts += [W1[:, 2 * H], W1[:, 2 * H + 1], f(pre + "edge_mlp.0.bias"), f(pre + "edge_mlp.2.bias"),
                  f(pre + "att_mlp.0.weight")[0], ba, f(pre + "coord_mlp.0.bias"), f(pre + "coord_mlp.2.weight")[0],
                  f(pre + "node_mlp.0.bias"), f(pre + "node_mlp.2.bias")]
    flat = torch.cat([p.reshape(-1) for p in parts]).contiguous()
    assert flat.numel() == ops.egnn_pack_floats(hidden, layers), (flat.numel(), ops.egnn_pack_floats(hidden, layers))
    return flat.to(device)


class EGNN_dynamics_AD2_cat(nn.Module):
    def __init__(self, n_particles, n_dimensions, hidden_nf=64, act_fn=torch.nn.SiLU(), n_layers=5, recurrent=True, attention=True,
                 tanh=True, atom_encoding_filename: str = "atom_types_ecoding.npy", data_dir="data/alanine", pdb_filename="",
                 agg="sum", M=128, condition_beta=False):
        super().__init__()
        unsupported = []
        if n_particles != 22: unsupported.append("n_particles=%r" % (n_particles,))
        if n_dimensions != 3: unsupported.append("n_dimensions=%r" % (n_dimensions,))
        if hidden_nf != 64: unsupported.append("hidden_nf=%r" % (hidden_nf,))
        if n_layers != 5: unsupported.append("n_layers=%r" % (n_layers,))
        if not isinstance(act_fn, torch.nn.SiLU): unsupported.append("act_fn=%r" % (act_fn,))
        if not (recurrent and attention and tanh): unsupported.append("recurrent/attention/tanh must all be True")
        if agg != "sum": unsupported.append("agg=%r" % (agg,))
        if not condition_beta: unsupported.append("condition_beta=False (the sampling loop conditions on beta)")
        if unsupported:
            raise NotImplementedError("pita_b200.EGNN_dynamics_AD2_cat builds only the egnn_dynamics_ad2_cat.yaml configuration "
                                      "natively; unsupported: " + ", ".join(unsupported))
        self._n_particles = n_particles
        self._n_dimensions = n_dimensions
        self.h_initial = torch.nn.functional.one_hot(atom_types_ad2(n_particles))  # reference :40
        self.condition_beta = condition_beta
        h_size = self.h_initial.size(1) + 1 + (1 if condition_beta else 0)
        self.egnn = _EGNNParams(h_size, hidden_nf, n_layers, act_fn)
        self.hidden_nf = hidden_nf
        self.n_layers = n_layers
        self.counter = 0
        self.M = M
        self._pack = None
        self._pack_key = None

    def packed_weights(self, device) -> torch.Tensor:
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._pack is None or self._pack_key != key:
            self._pack = pack_state_dict_ad2(self.state_dict(), self.hidden_nf, self.n_layers, device)
            self._pack_key = key
        return self._pack

    def forward(self, t, xs, beta):
        self.counter += 1
        out = ops.egnn_forward(self.packed_weights(xs.device), self.hidden_nf, self.n_layers, self._n_particles, t, xs, beta)
        return out.to(xs.dtype)
