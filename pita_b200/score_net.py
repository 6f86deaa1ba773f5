"""`ScoreNet` with the reference's interface (models/components/score_net.py), on `pita_egnn_score_div`.
`score_and_divergence` additionally returns tr(d score/dx) — what the reference gets from
`compiled_divergence_fn(score_net.forward)` (utils.py:30-51) — from the same kernel launch."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import ops


class ScoreNet(nn.Module):
    def __init__(self, model: nn.Module, precondition_beta: Optional[bool] = False, div_mode: Optional[str] = None):
        super().__init__()
        self.model = model
        self.precondition_beta = precondition_beta
        self.div_mode = div_mode  # None -> PITA_DIV_MODE / "bilinear"; "bilinear" | "fp32" | "3xtf32" | "tf32"

    def score_and_divergence(self, h_t, x_t, beta, need_div=True):
        m = self.model
        s, d = ops.egnn_score_div(m.packed_weights(x_t.device), m.hidden_nf, m.n_layers, m._n_particles, h_t, x_t, beta,
                                  need_div=need_div, mode=self.div_mode)
        if self.precondition_beta:  # :37-38; the divergence is linear in the score
            b = ops._expand(beta, x_t.shape[0], x_t.device)
            s = s * b[:, None]
            d = d * b if d is not None else None
        return s, d

    def forward(self, h_t, x_t, beta):
        return self.score_and_divergence(h_t, x_t, beta, need_div=False)[0]

    def denoiser(self, h_t, x_t, beta, return_score=False):
        score = self.forward(h_t, x_t, beta)
        d_theta = x_t + ops._expand(h_t, x_t.shape[0], x_t.device)[:, None] * score
        return (d_theta, score) if return_score else d_theta

    def reinitialize(self, model):
        self.model = model
