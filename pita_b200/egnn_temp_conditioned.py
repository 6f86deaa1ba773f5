"""Drop-in for `src.models.components.egnn_temp_conditioned.EGNN_dynamics` (reference :7-117).

Same constructor kwargs, same parameter names / creation order (so `state_dict`s and seeded
initialisation are interchangeable), `forward(t, xs, beta) -> [B, n*d]`.  The compute is the CUDA
kernel `pita_egnn_forward`; this module only owns the parameters and their packed device copy.
Only the configuration of configs/model/net/egnn_temp.yaml is built natively (hidden 32, 3 layers,
SiLU, recurrent, tanh, attention, agg=sum, time + temperature conditioning, 13 or 55 particles);
anything else raises — there is no PyTorch fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class _GCL(nn.Module):
    """Parameter container of one E_GCL block (reference :197-263); creation order matters for seeding."""

    def __init__(self, hidden_nf: int, act_fn: nn.Module, edges_in_d: int = 1):
        super().__init__()
        self.edge_mlp = nn.Sequential(nn.Linear(2 * hidden_nf + 1 + edges_in_d, hidden_nf), act_fn,
                                      nn.Linear(hidden_nf, hidden_nf), act_fn)
        self.node_mlp = nn.Sequential(nn.Linear(2 * hidden_nf, hidden_nf), act_fn, nn.Linear(hidden_nf, hidden_nf))
        last = nn.Linear(hidden_nf, 1, bias=False)
        torch.nn.init.xavier_uniform_(last.weight, gain=0.001)
        self.coord_mlp = nn.Sequential(nn.Linear(hidden_nf, hidden_nf), act_fn, last, nn.Tanh())
        self.att_mlp = nn.Sequential(nn.Linear(hidden_nf, 1), nn.Sigmoid())


class _EGNNParams(nn.Module):
    """Parameter container matching reference `EGNN` (:120-170)."""

    def __init__(self, in_node_nf: int, hidden_nf: int, n_layers: int, act_fn: nn.Module):
        super().__init__()
        self.hidden_nf = hidden_nf
        self.n_layers = n_layers
        self.embedding = nn.Linear(in_node_nf, hidden_nf)
        self.embedding_out = nn.Linear(hidden_nf, in_node_nf)  # unused by the velocity output, kept for state_dict parity
        for i in range(n_layers):
            self.add_module("gcl_%d" % i, _GCL(hidden_nf, act_fn))


def pack_state_dict(sd, hidden: int, layers: int, device) -> torch.Tensor:
    """Flat fp32 weight buffer in the layout csrc/egnn.cu (namespace pk) expects.

    `*_f` blocks hold W^T ([in][out]: lane `out` reads its weight row coalesced), `*_b` blocks hold W as
    torch stores it ([out][in]: the transposed product of the reverse pass)."""
    H = hidden
    f = lambda k: sd[k].detach().to(dtype=torch.float32, device="cpu")  # noqa: E731
    parts = [f("egnn.embedding.weight")[:, 0], f("egnn.embedding.weight")[:, 1], f("egnn.embedding.bias")]
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
    return flat.to(device)


class EGNN_dynamics(nn.Module):
    def __init__(self, n_particles, n_dimension, hidden_nf=64, act_fn=torch.nn.SiLU(), n_layers=4, recurrent=True,
                 attention=False, condition_time=True, tanh=False, agg="sum", energy=False, add_virtual=False,
                 condition_temperature=False, condition_on_temperature=None):
        super().__init__()
        # `condition_on_temperature` is injected by configs/model/energytemp.yaml:27-28 although the reference
        # constructor does not accept it; accepted and ignored here.
        unsupported = []
        if n_dimension != 3: unsupported.append("n_dimension=%r" % (n_dimension,))
        if n_particles not in (13, 55): unsupported.append("n_particles=%r" % (n_particles,))
        if hidden_nf != 32: unsupported.append("hidden_nf=%r" % (hidden_nf,))
        if n_layers != 3: unsupported.append("n_layers=%r" % (n_layers,))
        if not isinstance(act_fn, torch.nn.SiLU): unsupported.append("act_fn=%r" % (act_fn,))
        if not (recurrent and attention and tanh and condition_time and condition_temperature):
            unsupported.append("recurrent/attention/tanh/condition_time/condition_temperature must all be True")
        if agg != "sum" or energy or add_virtual: unsupported.append("agg/energy/add_virtual")
        if unsupported:
            raise NotImplementedError("pita_b200.EGNN_dynamics builds only the egnn_temp.yaml configuration natively; "
                                      "unsupported: " + ", ".join(unsupported))
        self.in_node_nf = 2
        self.egnn = _EGNNParams(self.in_node_nf, hidden_nf, n_layers, act_fn)
        self._n_particles = n_particles
        self._n_dimension = n_dimension
        self.hidden_nf = hidden_nf
        self.n_layers = n_layers
        self.condition_time = condition_time
        self.condition_temperature = condition_temperature
        self.counter = 0
        self._pack = None
        self._pack_key = None

    # ---- packed weights: refreshed whenever a parameter changed in place (optimizer step, load_state_dict,
    # EMA swap — energytemp_module.py:803-813) or moved device.
    def packed_weights(self, device) -> torch.Tensor:
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._pack is None or self._pack_key != key:
            self._pack = pack_state_dict(self.state_dict(), self.hidden_nf, self.n_layers, device)
            self._pack_key = key
        return self._pack

    def forward(self, t, xs, beta):
        self.counter += 1
        out = ops.egnn_forward(self.packed_weights(xs.device), self.hidden_nf, self.n_layers, self._n_particles, t, xs, beta)
        return out.to(xs.dtype)
