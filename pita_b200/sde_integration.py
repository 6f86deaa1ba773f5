"""`WeightedSDEIntegrator` with the reference's constructor and `integrate_sde` contract
(models/components/sde_integration.py:48-470), re-designed so that the whole annealed Feynman-Kac loop
stays on the GPU:

  per step (rank-local shard, no Python chunk loop):
     pita_egnn_energy      U_t, grad U_t, dU_t/dh          (energy net, fwd + hand-derived reverse)
     pita_egnn_score_div   s_t, div s_t                    (score net, fwd + forward-mode tangents)
     pita_sde_fk_step      drift assembly + Euler-Maruyama + remove_mean + raw FK weight drift, fused
     pita_fk_quantile_accumulate   per-chunk 0.9-quantile clamp (chunk = inference batch size) + a += dA*dt
     [resampling steps]    all-gather(a) -> softmax/clip -> fp64 scan -> search -> peer-memory gather
  no per-step device->host copy: log-weights / unique counts are stacked on the device and the SDE term
  diagnostics the reference ships to the CPU every step (:289,297) are opt-in (`record_sde_terms`).

Differences a caller can observe, all deliberate: `sde_terms_all` is empty unless `record_sde_terms=True`;
`integrate_sde` returns tensors detached; `x1.requires_grad` is not touched; with world_size > 1 the
full [N, D] set is only materialised once, at the end.
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import Callable, Optional

import numpy as np
import torch

from . import ops
from .distributed import ShardedResampler, shard_bounds, world_info
from .sdes import SDETerms, VEReverseSDE
from .utils import draw_u0


@contextmanager
def conditional_no_grad(condition):
    if condition:
        with torch.no_grad():
            yield
    else:
        yield


def mala_proposal(x, energy_function, dt):
    """reference :28-45 — returns (x_prop, log q(x'|x), log q(x|x')) using the fused kernels."""
    n = energy_function.n_particles
    _, grad = energy_function(x, return_force=True)
    x_prop, log_q_fwd = ops.mala_propose(x, grad, torch.randn_like(x), n, float(dt))
    _, grad_prop = energy_function(x_prop, return_force=True)
    back_mean = x_prop + 0.5 * dt * grad_prop
    log_q_bwd = -((x - back_mean) ** 2).sum(dim=1) / (2 * dt)
    return x_prop, log_q_fwd, log_q_bwd


class WeightedSDEIntegrator:
    def __init__(self, sde: VEReverseSDE, num_integration_steps: int, start_resampling_step: int,
                 end_resampling_step: int, lightning_module=None, partial_annealing_factor_schedule=None,
                 reverse_time: bool = True, diffusion_scale=1.0, time_range=1.0, resampling_interval=-1,
                 num_negative_time_steps=100, post_mcmc_steps=100, adaptive_mcmc=False, batch_size=None, no_grad=True,
                 resample_at_end=False, dt_negative_time=1e-4, do_langevin=False, should_mean_free=True,
                 # --- pita_b200 extensions (keyword-only in spirit; defaults reproduce the reference) ---
                 record_sde_terms: bool = False, collect_logweights: bool = True, fused_noise: bool = False,
                 noise_seed: int = 0, exchange: str = "auto", process_group=None) -> None:
        self.sde = sde
        self.num_integration_steps = num_integration_steps
        self.start_resampling_step = start_resampling_step
        self.end_resampling_step = end_resampling_step
        self.reverse_time = reverse_time
        self.diffusion_scale = diffusion_scale
        self.resampling_interval = resampling_interval
        self.time_range = time_range
        self.num_negative_time_steps = num_negative_time_steps
        self.post_mcmc_steps = post_mcmc_steps
        self.adaptive_mcmc = adaptive_mcmc
        self.dt_negative_time = dt_negative_time
        self.batch_size = batch_size
        self.no_grad = no_grad
        self.resample_at_end = resample_at_end
        self.do_langevin = do_langevin
        self.lightning_module = lightning_module
        self.should_mean_free = should_mean_free
        self.start_time = time_range if reverse_time else 0.0
        self.end_time = time_range - self.start_time
        self.record_sde_terms = record_sde_terms
        self.collect_logweights = collect_logweights
        self.fused_noise = fused_noise
        self.noise_seed = noise_seed
        self.exchange = exchange
        self.process_group = process_group
        # hooks used by the parity tests to inject the reference's random draws
        self.noise_fn: Optional[Callable[[int, torch.Tensor], torch.Tensor]] = None
        self.u0_fn: Optional[Callable[[int], float]] = None
        self.mcmc_noise_fn: Optional[Callable[[int, torch.Tensor], torch.Tensor]] = None    # MALA proposal noise of step k
        self.mcmc_uniform_fn: Optional[Callable[[int, torch.Tensor], torch.Tensor]] = None  # MALA accept/reject uniforms
        self.descent_noise_fn: Optional[Callable[[int, torch.Tensor], torch.Tensor]] = None  # Langevin noise of descent step k
        self._resampler = None
        self._resampler_key = None
        self._noise_call_seed = 0
        self._u0_gen = None

    # ------------------------------------------------------------------------------------------
    def maybe_remove_mean(self, x, energy_function):
        if self.should_mean_free:
            return ops.remove_mean(x, energy_function.n_particles)
        return x

    def _world(self):
        lm = self.lightning_module
        if lm is not None and getattr(lm, "trainer", None) is not None:
            return lm.trainer.world_size, lm.trainer.global_rank
        return world_info(self.process_group)

    def prepare(self, n_local: int, row_floats: int, device):
        """Sets up the sharded resampler and, with more than one rank, a uniform-offset stream shared by all ranks.

        The reference lets every rank draw u0 from its own CPU generator and relies on `seed_everything` having made
        them identical (utils.py:112).  With one rank the global CPU generator is used exactly like
        that; with several, rank 0 draws one seed from it and broadcasts it, so ranks cannot disagree on u0."""
        key = (n_local, row_floats, str(device), id(self.process_group), self.exchange)
        if self._resampler is None or self._resampler_key != key:  # symmetric-memory buffers are rendezvoused once, then reused
            self._resampler = ShardedResampler(n_local, row_floats, device, group=self.process_group, exchange=self.exchange)
            self._resampler_key = key
        # in-kernel Philox noise: a fresh key per integrate_sde call (drawn from torch's CPU generator, like the reference's
        # torch.randn stream advances from call to call), so two passes / epochs never reuse the same Brownian increments
        self._noise_call_seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64)) if self.fused_noise else 0
        self._u0_gen = None
        if self._resampler.world > 1:
            import torch.distributed as dist
            seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64)
            seed_dev = seed.to(device)
            dist.broadcast(seed_dev, src=0, group=self.process_group)
            self._u0_gen = torch.Generator().manual_seed(int(seed_dev.item()))

    def _draw_u0(self, step: int) -> float:
        if self.u0_fn is not None:
            return self.u0_fn(step)
        if self._u0_gen is not None:
            return float(torch.rand(size=(1,), dtype=torch.float64, generator=self._u0_gen))
        return draw_u0()

    def _step_scalars(self, t32: float, schedule):
        """Host-side scalar algebra of one step, folded into kernel arguments (float64 on the host)."""
        ns = self.sde.noise_schedule
        tt = torch.tensor(t32, dtype=torch.float64)
        g = float(ns.g(tt))
        dh = float(ns.dh_dt(tt)) if hasattr(ns, "dh_dt") else g * g
        return dict(h=float(ns.h(tt)), g=g, g2=g * g, dh_dt=dh, gamma=float(schedule.gamma(tt)),
                    dgamma=float(schedule.dgamma_dt(tt)))

    # ------------------------------------------------------------------------------------------
    def integrate_sde(self, x1: torch.Tensor, energy_function, annealing_factor_schedule, inverse_temperature=1.0,
                      annealing_factor_score=1.0, resampling_interval=None):
        if not x1.is_cuda:
            raise RuntimeError("pita_b200.WeightedSDEIntegrator runs on CUDA tensors only (no CPU fallback)")
        if resampling_interval is None:
            resampling_interval = self.resampling_interval
        n = energy_function.n_particles
        S = self.num_integration_steps
        world, rank = self._world()
        N_total, D = x1.shape
        lo, hi = shard_bounds(N_total, world, rank)
        if self.batch_size is None:
            self.batch_size = hi - lo
        beta = float(inverse_temperature)
        times = torch.linspace(self.start_time, self.end_time, S + 1)[:-1]  # reference :115-120 (float32)
        dt = self.time_range / S
        sqrt_dt = float(np.float32(np.sqrt(dt)))

        x = ops.N.as_f32(x1[lo:hi]).clone()
        a = torch.zeros(hi - lo, device=x.device, dtype=torch.float32)
        self.prepare(hi - lo, D, x.device)
        logweights, uniq, sde_terms_all = [], [], []

        with torch.no_grad():
            for step in range(S):
                x, a, n_unique, terms = self._fk_step(float(times[step]), step, x, a, dt, sqrt_dt, beta, n,
                                                      annealing_factor_schedule, energy_function, resampling_interval)
                if self.collect_logweights:
                    logweights.append(a)
                uniq.append(n_unique)
                if terms is not None:
                    sde_terms_all.append(terms)

            did_resampling = resampling_interval != -1 and resampling_interval < S
            if self.resample_at_end and did_resampling:  # reference :158-183
                t_end = float(times[min(self.end_resampling_step, S - 1)])
                sc = self._step_scalars(t_end, annealing_factor_schedule)
                target_logprob = energy_function(x)
                ht = torch.full((x.shape[0],), sc["h"], device=x.device, dtype=torch.float32)
                tt = torch.full((x.shape[0],), t_end, device=x.device, dtype=torch.float32)
                model_energy = self.sde.energy_net.forward_energy(ht, x, beta, pin=self.sde.pin_energy,
                                                                  energy_function=energy_function, t=tt)
                a_next = target_logprob + model_energy * sc["gamma"] + a
                a_full = self._resampler.gather_logweights(a_next)
                a_full = torch.clamp(a_full, max=torch.quantile(a_full, 0.9))  # one global quantile (:179)
                u0 = self._draw_u0(S)
                xin = self._stage_for_exchange(x)
                x, changes = self._resampler.resample(xin, a_full, u0)
                if self.collect_logweights:
                    logweights.append(a_full[lo:hi] if world > 1 else a_full)
                uniq.append(changes)

        # ---- hand results back in the reference's shapes: full [N, D] particle set, [S(+1), N] log-weights
        if world > 1:
            x = self._all_gather_rows(x)
            logweights = [self._all_gather_rows(w) for w in logweights]
        logweights = torch.stack(logweights) if logweights else torch.zeros(0, N_total, device=x.device)
        num_unique_idxs = [v if isinstance(v, int) else max(int(v.item()), 1) for v in uniq]

        if self.num_negative_time_steps > 0:
            x = self.negative_time_descent(x, energy_function)
        acceptance_rate_list = []
        if self.post_mcmc_steps > 0:
            if self.adaptive_mcmc:
                x, acceptance_rate_list = self.metropolis_hastings_mala_adaptive(
                    x, energy_function, dt_init=self.dt_negative_time, return_acceptance_rate=True)
            else:
                x, acceptance_rate_list = self.metropolis_hastings_mala(x, energy_function, return_acceptance_rate=True)
        return x, logweights, num_unique_idxs, sde_terms_all, acceptance_rate_list

    # ------------------------------------------------------------------------------------------
    def _stage_for_exchange(self, x):
        buf = self._resampler.particle_buffer()
        if buf is None:
            return x
        buf.copy_(x)
        return buf

    def _all_gather_rows(self, t):
        import torch.distributed as dist
        world, _ = self._world()
        out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.process_group)
        return out

    def sharded_step(self, t32, step, x, a, dt, sqrt_dt, beta, n, schedule, energy_function, resampling_interval):
        """One Euler-Maruyama + FK step on the rank-local shard: the sharded form of the reference's
        `ddp_batched_euler_maruyama_step` (:214-351), which takes and returns the FULL particle set on every rank.  Here
        x [N/W, D] and a [N/W] are this rank's block; only a resampling step communicates.  Returns
        (x_next, a_next, n_unique or its device counter, SDETerms or None).  `prepare()` must have been called."""
        return self._fk_step(t32, step, x, a, dt, sqrt_dt, beta, n, schedule, energy_function, resampling_interval)

    def _fk_step(self, t32, step, x, a, dt, sqrt_dt, beta, n, schedule, energy_function, resampling_interval):
        """One Euler-Maruyama + FK step on the rank-local shard (reference :214-351)."""
        sc = self._step_scalars(t32, schedule)
        B = x.shape[0]
        frozen = step < self.start_resampling_step  # :278-280 — particles stay put, weights stay 0
        zero_a = frozen or step >= self.end_resampling_step
        debias = self.sde.debias_inference
        terms = None
        if frozen:
            x_next, _ = ops.sde_fk_step(x, None, None, None, None, None, None, n, g2=0.0, gamma=0.0, dgamma_dt=0.0, dh_dt=0.0,
                                        dt=dt, sqrt_dt=sqrt_dt, noise_scale=0.0, debias=False, freeze_x=True,
                                        remove_mean=self.should_mean_free, want_a_raw=False)
            return x_next, torch.zeros_like(a), B * self._world()[0], None
        ht = torch.full((B,), sc["h"], device=x.device, dtype=torch.float32)
        no_score_net = self.sde.score_net is None   # the Laplacian branch (reference sdes.py:169-170, 204-216)
        if no_score_net:
            if not debias:
                raise AssertionError("f_not_debiased needs a score net (reference sdes.py:118)")
            score = div = None
        else:
            score, div = self.sde.score_net.score_and_divergence(ht, x, beta, need_div=debias)
        if debias:
            if self.sde.pin_energy:
                tt = torch.full((B,), t32, device=x.device, dtype=torch.float32)
                U, grad_u, dU_dt = self.sde.energy_net.energy_grad_dh(ht, x, beta, pin=True, energy_function=energy_function,
                                                                      t=tt, dh_dt=sc["dh_dt"])
                dE_dh, dh_dt = dU_dt, 1.0
            else:
                U, grad_u, dE_dh = self.sde.energy_net._terms(ht, x, beta, True, True)
                dh_dt = sc["dh_dt"]
            if no_score_net:  # b = -grad U g^2/2 and div b = -laplacian(U) g^2/2: the fused step sees them as its score / div
                score = -grad_u
                tt = torch.full((B,), t32, device=x.device, dtype=torch.float32) if self.sde.pin_energy else None
                div = -self.sde.energy_net.laplacian(ht, x, beta, pin=self.sde.pin_energy, t=tt)
        else:
            U = grad_u = dE_dh = None
            dh_dt = 0.0
        if self.fused_noise and self.noise_fn is None:
            noise = None
        else:
            noise = self.noise_fn(step, x) if self.noise_fn else torch.randn_like(x)
        out = self._resampler.particle_buffer() if self._will_resample(step, resampling_interval) else None
        x_next, a_raw = ops.sde_fk_step(x, grad_u, score, noise, div, dE_dh, U, n, g2=sc["g2"], gamma=sc["gamma"],
                                        dgamma_dt=sc["dgamma"], dh_dt=dh_dt, dt=dt, sqrt_dt=sqrt_dt,
                                        noise_scale=self.diffusion_scale * sc["g"], debias=debias, freeze_x=False,
                                        remove_mean=self.should_mean_free, seed=self.noise_seed ^ self._noise_call_seed,
                                        offset=step * 65536 + self._world()[1], want_a_raw=debias, out=out)
        if debias:
            a_next, drift_a = ops.fk_quantile_accumulate(a_raw, a, self.batch_size, 0.9, dt, zero_a,
                                                         want_drift=self.record_sde_terms)
        else:
            a_next, drift_a = torch.zeros_like(a), torch.zeros_like(a)
        if self.record_sde_terms:
            terms = SDETerms(drift_X=None, drift_A=drift_a, divergence_score=None if div is None else div * sc["g2"] / 2,
                             dUt_dt=None if dE_dh is None else dE_dh * dh_dt)
        if not self._will_resample(step, resampling_interval):
            return x_next, a_next, B * self._world()[0], terms
        # ---- resample on the global weights (reference :292-295)
        a_full = self._resampler.gather_logweights(a_next)
        u0 = self._draw_u0(step)
        x_res, changes = self._resampler.resample(x_next, a_full, u0)
        return x_res, torch.zeros_like(a_next), changes, terms

    def _will_resample(self, step, resampling_interval):
        return not (resampling_interval == -1 or (step + 1) % resampling_interval != 0
                    or step < self.start_resampling_step or step >= self.end_resampling_step)

    # ------------------------------------------------------------------------------------------
    # post-processing on the target (reference :353-470)
    def negative_time_descent(self, x, energy_function):
        n = energy_function.n_particles
        for k in range(self.num_negative_time_steps):
            _, drift = energy_function(x, return_force=True)
            noise = None
            if self.do_langevin:
                noise = self.descent_noise_fn(k, x) if self.descent_noise_fn else torch.randn_like(x)
            x = ops.descent_step(x, drift, noise, n, float(self.dt_negative_time), self.should_mean_free)
        return x

    def _mala(self, x, energy_function, dt, adaptive, return_acceptance_rate):
        n = energy_function.n_particles
        x_curr = ops.N.as_f32(x).clone()
        logp_curr = energy_function(x_curr)
        valid = torch.isfinite(logp_curr)
        x_valid, x_invalid = x_curr[valid].contiguous(), x_curr[~valid]
        logp_valid = logp_curr[valid].contiguous()
        rates = []
        if not adaptive and not return_acceptance_rate:
            # Reference parity (:386, :401): without return_acceptance_rate the non-adaptive loop prints an unbound
            # `acceptance_rate`, the NameError is swallowed by its try/except before the update lines run, and the particles
            # come back unchanged and in their original order.  integrate_sde always asks for the rates (:203-207).
            return x_curr, None
        for k in range(self.post_mcmc_steps):
            if x_valid.shape[0] == 0:
                continue
            _, grad = energy_function(x_valid, return_force=True)
            xi = self.mcmc_noise_fn(k, x_valid) if self.mcmc_noise_fn else torch.randn_like(x_valid)
            x_prop, log_q_fwd = ops.mala_propose(x_valid, grad, xi, n, float(dt))
            logp_prop, grad_prop = energy_function(x_prop, return_force=True)
            uni = self.mcmc_uniform_fn(k, logp_valid) if self.mcmc_uniform_fn else torch.rand_like(logp_valid)
            acc = ops.mala_accept(x_valid, logp_valid, x_prop, logp_prop, grad_prop, log_q_fwd,
                                  uni, n, float(dt),
                                  bool(energy_function.is_molecule and self.should_mean_free) if not adaptive
                                  else bool(energy_function.is_molecule))
            rate = acc.mean().item()
            if return_acceptance_rate:
                rates.append(rate)
            if adaptive:  # reference :440-443
                dt = dt * 1.1 if rate > 0.55 else dt / 1.1
        x_curr = torch.cat([x_valid, x_invalid], dim=0)  # reference :400 (valid rows first)
        return x_curr, (rates if return_acceptance_rate else None)

    def metropolis_hastings_mala(self, x, energy_function, return_acceptance_rate=False):
        return self._mala(x, energy_function, self.dt_negative_time, False, return_acceptance_rate)

    def metropolis_hastings_mala_adaptive(self, x, energy_function, dt_init, return_acceptance_rate=False):
        return self._mala(x, energy_function, dt_init, True, return_acceptance_rate)
