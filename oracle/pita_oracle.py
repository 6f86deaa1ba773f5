"""TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of PITA's annealed-sampling hot path.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this module; nothing under `pita_b200/` does.  It is the checker, never
the thing shipped or measured as the product.

What it restates (reference = taraak/pita @ e987a7f, paths relative to pita/src/):
  * EGNN denoiser               models/components/egnn_temp_conditioned.py:56-93,172-194,265-356
  * EDM-preconditioned wrappers models/components/energy_net.py:14-62, score_net.py:13-43
  * exact divergence            models/components/utils.py:43-51 (vmap(jacrev) trace)
  * FK drift terms              models/components/sdes.py:117-251
  * Euler-Maruyama + FK loop    models/components/sde_integration.py:98-351
  * systematic resampler        models/components/utils.py:111-120
  * schedules                   noise_schedules.py:98-125, annealing_factor_schedules.py:20-109
  * centre-of-mass removal      utils/data_utils.py:4-26
  * Lennard-Jones target        energies/lennardjones_energy.py:34-39,121-155,213-227
    (+ bgflow geometry helpers, an un-vendored dependency: see oracle/_ref_import.py)
  * mean-free prior             energies/base_prior.py:77-83

It is written as pure functions over a weight dict on dense [B,n,n] pair tensors (the
reference builds edge lists and scatter_adds), runs in whatever dtype it is handed (fp64 for
ground truth, fp32 to mimic the reference), and uses autograd / torch.func for the derivatives
exactly where the reference does.

PINNING: the reference's own tests hold no golden vectors for this path, so
the oracle is pinned against outputs of the unmodified reference run in the build container:
`oracle/make_golden.py` imports /root/reference under import stubs and writes
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks this file against them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, Optional

import numpy as np
import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# schedules
# --------------------------------------------------------------------------------------
@dataclass
class EDMSchedule:
    """noise_schedules.py:98-125 (ElucidatingNoiseSchedule)."""

    sigma_min: float
    sigma_max: float = 80.0
    rho: float = 7.0

    @property
    def _a(self):
        return self.sigma_max ** (1.0 / self.rho)

    @property
    def _b(self):
        return self.sigma_min ** (1.0 / self.rho) - self.sigma_max ** (1.0 / self.rho)

    def h(self, t):
        return (self._a + (1 - t) * self._b) ** (2 * self.rho)

    def dh_dt(self, t):
        return -2 * self.rho * self._b * (self._a + (1 - t) * self._b) ** (2 * self.rho - 1)

    def g(self, t):
        return self.dh_dt(t) ** 0.5


@dataclass
class ConstGamma:
    """annealing_factor_schedules.py:20-32."""

    value: float

    def gamma(self, t):
        return torch.ones_like(t) * self.value

    def dgamma_dt(self, t):
        return torch.zeros_like(t)


@dataclass
class LinearGamma:
    """annealing_factor_schedules.py:35-71."""

    value: float
    start: float
    t_start: float = 1.0
    t_end: float = 0.0

    def _slope(self):
        return (self.value - self.start) / (self.t_end - self.t_start)

    def gamma(self, t):
        lin = self._slope() * (t - self.t_start) + self.start
        out = torch.where(t < self.t_end, torch.full_like(t, self.value), lin)
        return torch.where(t > self.t_start, torch.full_like(t, self.start), out)

    def dgamma_dt(self, t):
        inside = (t <= self.t_start) & (t >= self.t_end)
        return torch.where(inside, torch.full_like(t, self._slope()), torch.zeros_like(t))


@dataclass
class SigmoidGamma:
    """annealing_factor_schedules.py:74-109."""

    value: float
    start: float
    t_start: float = 1.0
    t_end: float = 0.0
    sharpness: float = 10.0

    def _smooth(self, t):
        centre = (self.t_start + self.t_end) / 2
        width = self.t_start - self.t_end
        return 1 / (1 + torch.exp(-self.sharpness * (centre - t) / width))

    def gamma(self, t):
        return self.start + (self.value - self.start) * self._smooth(t)

    def dgamma_dt(self, t):
        s = self._smooth(t)
        width = self.t_start - self.t_end
        return (self.value - self.start) * (self.sharpness / width) * s * (1 - s)


# --------------------------------------------------------------------------------------
# geometry helpers
# --------------------------------------------------------------------------------------
def centre(x: Tensor, n: int, d: int = 3) -> Tensor:
    """data_utils.py:4-26 remove_mean."""
    v = x.reshape(-1, n, d)
    return (v - v.mean(dim=1, keepdim=True)).reshape(x.shape)


def mean_free_prior(num: int, n: int, scale: float, gen: Optional[torch.Generator] = None,
                    dtype=torch.float32) -> Tensor:
    """base_prior.py:77-83."""
    z = torch.randn(num, n * 3, generator=gen, dtype=dtype) * scale
    return centre(z, n)


# --------------------------------------------------------------------------------------
# EGNN denoiser (temperature-conditioned), dense restatement
# --------------------------------------------------------------------------------------
def _silu(v):
    return v * torch.sigmoid(v)


def egnn_layer_count(sd: Dict[str, Tensor]) -> int:
    k = 0
    while f"egnn.gcl_{k}.edge_mlp.0.weight" in sd:
        k += 1
    return k


def egnn_velocity(sd: Dict[str, Tensor], tcond: Tensor, y: Tensor, beta: Tensor, n: int,
                  coords_range: float = 15.0, skip_dead: bool = False, node_feat: Optional[Tensor] = None) -> Tensor:
    """EGNN_dynamics.forward (egnn_temp_conditioned.py:56-93) with condition_time and
    condition_temperature, recurrent, tanh, attention, agg='sum', norm_diff (the configuration of
    configs/model/net/egnn_temp.yaml).  tcond, beta: [B]; y: [B, 3n] -> [B, 3n].

    skip_dead=True drops the last layer's node update, which cannot influence the output (the
    CUDA kernels skip it); used to check that claim.
    """
    B = y.shape[0]
    L = egnn_layer_count(sd)
    rng = coords_range / L  # egnn_temp_conditioned.py:143
    x0 = y.reshape(B, n, 3)
    # node features (:63-70, :78): the reference concatenates [t]*n and [beta]*n along the LAST
    # dim ([B,2n]) and then reshapes to [B*n, 2], so node k receives (f[2k], f[2k+1]) with
    # f = (t,...,t, beta,...,beta): the first n//2 nodes see (t,t), the last n//2 see (beta,beta)
    # and (for odd n) the middle node sees (t,beta).  Reproduced verbatim — parity requires it.
    flat = torch.cat([tcond[:, None].expand(B, n), beta[:, None].expand(B, n)], dim=-1)  # [B,2n]
    feat = flat.reshape(B, n, 2) if node_feat is None else node_feat  # node_feat: [B,n,F] (egnn_velocity_ad2)
    h = feat @ sd["egnn.embedding.weight"].T + sd["egnn.embedding.bias"]  # [B,n,H] (:179)
    off = ~torch.eye(n, dtype=torch.bool)
    offf = off.to(y.dtype)[None, :, :, None]

    def pair_diff(x):
        return x[:, :, None, :] - x[:, None, :, :]  # [B,i,j,3] = x_i - x_j

    ea = pair_diff(x0).pow(2).sum(-1, keepdim=True)  # edge_attr from input coords (:79)
    x = x0
    for l in range(L):
        pre = f"egnn.gcl_{l}."
        W1, b1 = sd[pre + "edge_mlp.0.weight"], sd[pre + "edge_mlp.0.bias"]
        W2, b2 = sd[pre + "edge_mlp.2.weight"], sd[pre + "edge_mlp.2.bias"]
        H = W2.shape[0]
        dlt = pair_diff(x)
        r2 = dlt.pow(2).sum(-1, keepdim=True)  # radial (:351)
        dhat = dlt / (torch.sqrt(r2 + 1e-8) + 1)  # (:353-354)
        hi = h[:, :, None, :].expand(B, n, n, H)  # source = h[row] (receiver i)
        hj = h[:, None, :, :].expand(B, n, n, H)  # target = h[col]
        z1 = torch.cat([hi, hj, r2, ea], dim=-1) @ W1.T + b1  # (:270-271)
        m = _silu(_silu(z1) @ W2.T + b2)
        att = torch.sigmoid(m @ sd[pre + "att_mlp.0.weight"].T + sd[pre + "att_mlp.0.bias"])
        m = m * att * offf  # (:273-275); diagonal is not an edge
        zc = _silu(m @ sd[pre + "coord_mlp.0.weight"].T + sd[pre + "coord_mlp.0.bias"])
        phi = torch.tanh(zc @ sd[pre + "coord_mlp.2.weight"].T) * rng  # (:297-298)
        x = x + (dhat * phi * offf).sum(dim=2)  # (:305-318) aggregate over senders j
        if not (skip_dead and l == L - 1):
            agg = m.sum(dim=2)  # (:284)
            z3 = torch.cat([h, agg], dim=-1) @ sd[pre + "node_mlp.0.weight"].T + sd[pre + "node_mlp.0.bias"]
            h = h + _silu(z3) @ sd[pre + "node_mlp.2.weight"].T + sd[pre + "node_mlp.2.bias"]  # (:289-291)
    vel = x - x0
    vel = vel - vel.mean(dim=1, keepdim=True)  # (:84)
    return vel.reshape(B, n * 3)


def ad2_atom_types(n: int = 22) -> Tensor:
    """EGNN_dynamics_AD2_cat.get_h_initial for alanine dipeptide (egnn_dynamics_ad2_cat.py:67-73): one class per
    atom except the three hydrogen triples that share a class; the largest label is 20, so one_hot has 21 columns."""
    if n != 22:
        raise NotImplementedError("only the 22-atom alanine dipeptide typing is restated")
    t = torch.arange(22)
    t[[0, 2, 3]] = 2
    t[[19, 20, 21]] = 20
    t[[11, 12, 13]] = 12
    return t


def egnn_velocity_ad2(sd: Dict[str, Tensor], t: Tensor, y: Tensor, beta: Tensor, n: int = 22,
                      condition_beta: bool = True, coords_range: float = 15.0) -> Tensor:
    """EGNN_dynamics_AD2_cat.forward (egnn_dynamics_ad2_cat.py:158-194) over egnn.EGNN (egnn.py:108-184): the same E_GCL
    stack as egnn_velocity (recurrent, tanh, attention, agg='sum'; H and L read from the weights: 64 / 5 in
    configs/model/net/egnn_ad2.yaml-style use), but every node carries one_hot(atom type) ++ [t] (++ [beta]) — here the
    time / temperature columns really are per node (:176-186), unlike the interleaving quirk of the LJ network.
    SURVEY §8 row a8'.  TEST INFRASTRUCTURE (the CUDA path is csrc/egnn_ad2.cu)."""
    B = y.shape[0]
    onehot = torch.nn.functional.one_hot(ad2_atom_types(n)).to(y.dtype)  # [n, 21]
    cols = [onehot[None].expand(B, n, onehot.shape[1]), t[:, None, None].expand(B, n, 1)]
    if condition_beta:
        cols.append(beta[:, None, None].expand(B, n, 1))
    return egnn_velocity(sd, t, y, beta, n, coords_range=coords_range, node_feat=torch.cat(cols, dim=-1))


def _velocity(sd, tcond, y, beta, n):
    """The denoiser the weights belong to: 2 embedding inputs = EGNN_dynamics (LJ), 23 = EGNN_dynamics_AD2_cat."""
    if sd["egnn.embedding.weight"].shape[1] == 23:
        return egnn_velocity_ad2(sd, tcond, y, beta, n)
    return egnn_velocity(sd, tcond, y, beta, n)


# --------------------------------------------------------------------------------------
# EDM preconditioning wrappers
# --------------------------------------------------------------------------------------
def _coeffs(ht: Tensor):
    c_s = 1 / (1 + ht)
    c_in = 1 / (1 + ht) ** 0.5
    c_out = ht ** 0.5 * c_in
    c_noise = 0.125 * torch.log(ht)
    return c_s, c_in, c_out, c_noise


def model_energy(sd, ht: Tensor, x: Tensor, beta, n: int, precondition_beta=False) -> Tensor:
    """EnergyNet.forward_energy with pin=False (energy_net.py:14-48)."""
    beta = beta * torch.ones(x.shape[0], dtype=x.dtype)
    c_s, c_in, c_out, c_noise = _coeffs(ht)
    yy = c_in[:, None] * x
    u = (_velocity(sd, c_noise, yy, beta, n) * yy).sum(dim=1)
    e = (1 - c_s) / (2 * ht) * torch.linalg.norm(x, dim=-1) ** 2 - c_out / (c_in * ht) * u
    if precondition_beta:
        e = e * beta
    return e


def model_score(sd, ht: Tensor, x: Tensor, beta, n: int, precondition_beta=False) -> Tensor:
    """ScoreNet.forward (score_net.py:13-43)."""
    beta = beta * torch.ones(x.shape[0], dtype=x.dtype)
    c_s, c_in, c_out, c_noise = _coeffs(ht)
    den = c_s[:, None] * x + c_out[:, None] * _velocity(sd, c_noise, c_in[:, None] * x, beta, n)
    if precondition_beta:
        den = den * beta[:, None] + (1 - beta[:, None]) * x
    return (den - x) / ht[:, None]


def exact_divergence(fn: Callable[[Tensor, Tensor], Tensor], ht: Tensor, x: Tensor) -> Tensor:
    """utils.py:43-51: per-sample full Jacobian by vmap(jacrev), then its trace."""
    from torch.func import jacrev, vmap

    def one(h1, x1):
        return fn(h1[None], x1[None])[0]

    jac = vmap(jacrev(one, argnums=1))(ht, x)
    return jac.diagonal(dim1=-2, dim2=-1).sum(-1).detach()


def exact_laplacian(fn: Callable[[Tensor, Tensor], Tensor], ht: Tensor, x: Tensor) -> Tensor:
    """utils.py:68-77 (compute_laplacian_exact): per-sample Hessian of a scalar field by vmap(hessian), then its trace.
    fn(h[1], x[1, D]) -> [1]."""
    from torch.func import hessian, vmap

    def one(h1, x1):
        return fn(h1[None], x1[None])[0]

    hes = vmap(hessian(one, argnums=1))(ht, x)
    return hes.diagonal(dim1=-2, dim2=-1).sum(-1).detach()


# --------------------------------------------------------------------------------------
# FK drift (VEReverseSDE.f) and diffusion
# --------------------------------------------------------------------------------------
@dataclass
class Drift:
    drift_x: Tensor
    drift_a: Tensor
    div_b: Optional[Tensor] = None
    cross: Optional[Tensor] = None
    du_dt: Optional[Tensor] = None
    energy: Optional[Tensor] = None
    grad_u: Optional[Tensor] = None
    score: Optional[Tensor] = None
    drift_a_raw: Optional[Tensor] = None


def quantile_clamp(v: Tensor, q: float = 0.9) -> Tensor:
    """sdes.py:230 — clamp at the chunk's own q-quantile (torch.quantile, linear interp)."""
    return torch.clamp(v, max=torch.quantile(v, q))


def fk_drift(sd_energy, sd_score, sched: EDMSchedule, gamma_sched, t: float, x: Tensor, beta: float,
             n: int, debias: bool = True) -> Drift:
    """VEReverseSDE.f for one chunk (sdes.py:130-239), pin_energy False; sd_score=None is the no-score-net (Laplacian)
    branch (oracle only — the CUDA path raises NotImplementedError for it)."""
    B = x.shape[0]
    tt = torch.full((B,), float(t), dtype=x.dtype)
    gam = gamma_sched.gamma(tt)
    g2 = sched.g(tt) ** 2
    if not debias:  # f_not_debiased (sdes.py:117-128)
        s = model_score(sd_score, sched.h(tt), x, beta, n)
        return Drift(drift_x=(gam[:, None] * s * g2[:, None]).detach(), drift_a=torch.zeros(B, dtype=x.dtype))
    with torch.enable_grad():
        xr = x.detach().clone().requires_grad_(True)
        tr = tt.clone().requires_grad_(True)
        ht = sched.h(tr)
        u = model_energy(sd_energy, ht, xr, beta, n)
        grad_u, du_dt = torch.autograd.grad(u.sum(), (xr, tr))
        s = None if sd_score is None else model_score(sd_score, ht, xr, beta, n).detach()
    ht = ht.detach()
    u = u.detach()
    if sd_score is None:  # no score net (sdes.py:150-153, 204-216): b = -grad U g^2/2, div b = -laplacian(U) g^2/2
        b = -grad_u * g2[:, None] / 2
        div_b = -exact_laplacian(lambda h1, x1: model_energy(sd_energy, h1, x1, beta, n), ht, x.detach()) * g2 / 2
    else:
        b = s * g2[:, None] / 2
        div_s = exact_divergence(lambda h1, x1: model_score(sd_score, h1, x1, beta, n), ht, x.detach())
        div_b = div_s * g2 / 2
    drift_x = gam[:, None] * (-grad_u) * g2[:, None] / 2 + gam[:, None] * b
    cross = (-grad_u * b).sum(-1)
    raw = gam * gam * cross + gam * div_b + gam * du_dt + gamma_sched.dgamma_dt(tt) * u
    return Drift(drift_x=drift_x, drift_a=quantile_clamp(raw), div_b=div_b, cross=cross, du_dt=du_dt,
                 energy=u, grad_u=grad_u, score=s, drift_a_raw=raw)


# --------------------------------------------------------------------------------------
# systematic resampling
# --------------------------------------------------------------------------------------
def clipped_softmax(logits: Tensor) -> Tensor:
    """utils.py:114 — softmax then clip to [1e-6, 1], NOT renormalised."""
    return torch.clip(torch.softmax(logits, dim=-1), 1e-6, 1.0)


def systematic_indices(weights32: np.ndarray, u0: float) -> np.ndarray:
    """utils.py:111-120 given the clipped weights (fp32) and the fp64 uniform offset u0.

    bins = fp32 values of the fp64-accumulated running sum (what torch.cumsum does on CPU for a
    float32 input — checked against the reference in make_golden.py); u_i = (u0 + i/N) mod 1 in
    fp64; ids_i = #bins strictly below u_i (np.digitize right=True), clamped to N-1.
    """
    w = np.asarray(weights32, dtype=np.float32)
    N = w.shape[0]
    bins = np.cumsum(w.astype(np.float64)).astype(np.float32)
    # `1 / bs * torch.arange(bs)` is evaluated in float32 (python float x int64 tensor -> default
    # dtype) before the float64 add (utils.py:113): the grid is fl32(fl32(1/N) * fl32(i)).
    grid = (np.float32(1.0 / N) * np.arange(N).astype(np.float32)).astype(np.float64)
    u = (np.float64(u0) + grid) % 1.0
    ids = np.searchsorted(bins.astype(np.float64), u, side="left")
    ids[ids == N] = N - 1
    return ids.astype(np.int64)


def systematic_resample(logits: Tensor, u0: float) -> np.ndarray:
    return systematic_indices(clipped_softmax(logits.float()).numpy(), u0)


# --------------------------------------------------------------------------------------
# Lennard-Jones target
# --------------------------------------------------------------------------------------
def lj_energy(x: Tensor, n: int, eps_sqrt: float = 1e-6, energy_factor: float = 1.0,
              oscillator_scale: float = 1.0) -> Tensor:
    """LennardJonesPotential._energy (lennardjones_energy.py:121-143), smooth=False:
    sum over ORDERED pairs of r^-12 - 2 r^-6 with r = sqrt(|x_i-x_j|^2 + 1e-6), plus the
    harmonic centre-of-mass term."""
    B = x.shape[0]
    v = x.reshape(B, n, 3)
    d2 = (v[:, :, None, :] - v[:, None, :, :]).pow(2).sum(-1)
    off = ~torch.eye(n, dtype=torch.bool)
    r = torch.sqrt(d2[:, off] + eps_sqrt)  # [B, n(n-1)]
    e = ((1.0 / r) ** 12 - 2 * (1.0 / r) ** 6).sum(-1) * energy_factor
    c = v - v.mean(dim=1, keepdim=True)
    return e + 0.5 * c.pow(2).sum(dim=(-2, -1)) * oscillator_scale


def lj_logprob_force(x: Tensor, n: int, temperature: float = 1.0):
    """LennardJonesEnergy.__call__(samples, return_force=True) (lennardjones_energy.py:213-227)."""
    with torch.enable_grad():
        xr = x.detach().clone().requires_grad_(True)
        lp = -lj_energy(xr, n) / temperature
        (f,) = torch.autograd.grad(lp.sum(), xr)
    return lp.detach(), f.detach()


# --------------------------------------------------------------------------------------
# the annealed FK loop
# --------------------------------------------------------------------------------------
@dataclass
class LoopConfig:
    n: int
    steps: int
    chunk: int
    beta: float = 1.0
    resampling_interval: int = 1
    start_resampling_step: int = 0
    end_resampling_step: int = 10 ** 9
    diffusion_scale: float = 1.0
    time_range: float = 1.0
    debias: bool = True
    resample_at_end: bool = False
    temperature: float = 1.0  # target temperature for the end resample


def integrate(sd_energy, sd_score, sched: EDMSchedule, gamma_sched, cfg: LoopConfig, x1: Tensor,
              noise_fn: Callable[[int, Tensor], Tensor], u0_fn: Callable[[int], float]):
    """WeightedSDEIntegrator.integrate_sde with world_size 1, no post-processing
    (sde_integration.py:98-185, 214-351).  noise_fn(step, x_chunk) supplies N(0,1) draws per chunk
    in the reference's call order; u0_fn(step) the resampler's uniform offset."""
    S = cfg.steps
    dtype = x1.dtype
    times = torch.linspace(cfg.time_range, 0.0, S + 1, dtype=dtype)[:-1]  # default dtype in the reference (:115-120)
    dt = cfg.time_range / S
    x = x1.clone()
    a = torch.zeros(x.shape[0], dtype=dtype)
    logw, uniq = [], []
    N = x.shape[0]
    for step in range(S):
        t = float(times[step])
        dx_list, da_list, diff_list = [], [], []
        for lo in range(0, N, cfg.chunk):  # chunk loop (sde_integration.py:312-343)
            xc = x[lo:lo + cfg.chunk]
            d = fk_drift(sd_energy, sd_score, sched, gamma_sched, t, xc, cfg.beta, cfg.n, cfg.debias)
            gt = sched.g(torch.full((xc.shape[0],), t, dtype=dtype))
            diff_list.append(cfg.diffusion_scale * gt[:, None] * noise_fn(step, xc))  # sdes.py:245-251
            dx_list.append(d.drift_x)
            da_list.append(d.drift_a)
        drift_x, drift_a, diff = torch.cat(dx_list), torch.cat(da_list), torch.cat(diff_list)
        x_next = x + drift_x * dt + diff * np.sqrt(dt)  # :347-349
        a_next = a + drift_a * dt
        if step < cfg.start_resampling_step:  # :278-282
            a_next = torch.zeros_like(a_next)
            x_next = x
        if step >= cfg.end_resampling_step:
            a_next = torch.zeros_like(a_next)
        n_unique = N
        do = not (cfg.resampling_interval == -1 or (step + 1) % cfg.resampling_interval != 0
                  or step < cfg.start_resampling_step or step >= cfg.end_resampling_step)
        if do:  # :292-295
            ids = systematic_resample(a_next, u0_fn(step))
            x_next = x_next[torch.from_numpy(ids)]
            a_next = torch.zeros_like(a_next)
            n_unique = len(np.unique(ids))
        x = centre(x_next, cfg.n).detach()  # :148
        a = a_next.detach()
        logw.append(a)
        uniq.append(n_unique)
    did = cfg.resampling_interval != -1 and cfg.resampling_interval < S
    if cfg.resample_at_end and did:  # :158-183
        t_end = float(times[min(cfg.end_resampling_step, S - 1)])
        tt = torch.full((N,), t_end, dtype=dtype)
        target_lp = -lj_energy(x, cfg.n) / cfg.temperature
        me = model_energy(sd_energy, sched.h(tt), x, cfg.beta, cfg.n)
        a_end = quantile_clamp(target_lp + me * gamma_sched.gamma(tt) + a)
        ids = systematic_resample(a_end, u0_fn(S))
        x = x[torch.from_numpy(ids)]
        logw.append(a_end.detach())
        uniq.append(len(np.unique(ids)))
    return x, torch.stack(logw), uniq


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------
def md_shaped_coords(num: int, n: int, seed: int, spacing: float = 1.1, jitter: float = 0.08,
                     dtype=torch.float32) -> Tensor:
    """First n sites of a simple-cubic lattice (spacing 1.1 ~ r_m) + N(0, 0.08^2) jitter, COM removed."""
    side = int(math.ceil(n ** (1.0 / 3.0) - 1e-9))
    sites = [(i, j, k) for i in range(side) for j in range(side) for k in range(side)][:n]
    base = torch.tensor(sites, dtype=torch.float64) * spacing
    gen = torch.Generator().manual_seed(seed)
    x = base[None] + jitter * torch.randn(num, n, 3, generator=gen, dtype=torch.float64)
    x = x - x.mean(dim=1, keepdim=True)
    return x.reshape(num, n * 3).to(dtype)


def random_egnn_state(n_layers: int = 3, hidden: int = 32, seed: int = 12345, dtype=torch.float32,
                      coord_gain: float = 0.001, in_nf: int = 2) -> Dict[str, Tensor]:
    """Random-init weights with the reference's parameter names and init distributions
    (nn.Linear default init; xavier_uniform gain 0.001 on the last coord layer,
    egnn_temp_conditioned.py:245-246).  coord_gain can be raised in tests so that the coordinate
    branch is numerically visible."""
    gen = torch.Generator().manual_seed(seed)

    def lin(out_f, in_f, bias=True):
        bound = 1.0 / math.sqrt(in_f)
        w = (torch.rand(out_f, in_f, generator=gen, dtype=torch.float64) * 2 - 1) * bound
        b = (torch.rand(out_f, generator=gen, dtype=torch.float64) * 2 - 1) * bound if bias else None
        return w, b

    sd: Dict[str, Tensor] = {}
    H = hidden
    w, b = lin(H, in_nf)  # 2 = [t, beta] (LJ); 23 = one_hot(atom type) ++ [t, beta] (alanine dipeptide)
    sd["egnn.embedding.weight"], sd["egnn.embedding.bias"] = w, b
    w, b = lin(2, H)
    sd["egnn.embedding_out.weight"], sd["egnn.embedding_out.bias"] = w, b
    for l in range(n_layers):
        pre = f"egnn.gcl_{l}."
        for name, (o, i) in {"edge_mlp.0": (H, 2 * H + 2), "edge_mlp.2": (H, H), "node_mlp.0": (H, 2 * H),
                             "node_mlp.2": (H, H), "coord_mlp.0": (H, H), "att_mlp.0": (1, H)}.items():
            w, b = lin(o, i)
            sd[pre + name + ".weight"], sd[pre + name + ".bias"] = w, b
        bound = coord_gain * math.sqrt(6.0 / (H + 1))
        sd[pre + "coord_mlp.2.weight"] = (torch.rand(1, H, generator=gen, dtype=torch.float64) * 2 - 1) * bound
    return {k: v.to(dtype) for k, v in sd.items()}


def w2_1d(a: np.ndarray, b: np.ndarray) -> float:
    """1-D Wasserstein-2 between equal-size samples (energy-histogram parity, SURVEY §8d gate 3)."""
    a, b = np.sort(np.asarray(a, dtype=np.float64)), np.sort(np.asarray(b, dtype=np.float64))
    m = min(len(a), len(b))
    qa = np.quantile(a, (np.arange(m) + 0.5) / m)
    qb = np.quantile(b, (np.arange(m) + 0.5) / m)
    return float(np.sqrt(np.mean((qa - qb) ** 2)))
