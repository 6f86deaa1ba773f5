"""`generate_samples` — the only caller of the annealed loop in the reference (models/energytemp_module.py:237-298;
SURVEY §8f-2), as a free function over the same ingredients the Lightning module holds in `hparams`:

    prior scale  sqrt(h(t_start) / gamma(t_start))                      (:250-259)
    samples      integrate_sde(prior.sample(N), ...)                    (:268-280)
    log-weights  a SECOND pass without resampling on the first `inference_batch_size` prior samples, selected by
                 resampling_interval = num_integration_steps + 1        (:281-297)

Pure host orchestration: every tensor operation happens inside `prior.sample` and `integrate_sde` (CUDA kernels through the
C ABI; no CPU fallback there).  Differences from the reference, both deliberate: the prior samples are not cloned before the
first pass (`integrate_sde` copies its shard anyway), and the second pass reuses the integrator's resident workspace.
"""
from __future__ import annotations

from typing import Callable, Optional


def prior_scale(noise_schedule, annealing_factor_schedule, t_start):
    """sqrt(h(t_start) / gamma(t_start)) — energytemp_module.py:251-257."""
    return (noise_schedule.h(t_start) / annealing_factor_schedule.gamma(t_start)) ** 0.5


def generate_samples(*, weighted_sde_integrator, energy_function, num_samples: int, noise_schedule,
                     annealing_factor_schedule: Callable, partial_prior: Callable, t_start, device,
                     inference_batch_size: int, num_integration_steps: int, inverse_temp: Optional[float] = 1.0,
                     annealing_factor: Optional[float] = 1.0, annealing_factor_score: Optional[float] = None,
                     return_logweights: bool = False):
    """Returns `(samples, num_unique_idxs, sde_terms, acceptance_rate_list)`, or with `return_logweights`
    `(samples, samples_not_resampled, logweights, num_unique_idxs, sde_terms, acceptance_rate_list)` — the reference's tuples.

    `annealing_factor_schedule` and `partial_prior` are the Hydra partials of the reference config
    (`annealing_factor_schedule(annealing_factor=...)`, `partial_prior(scale=..., device=...)`)."""
    schedule = annealing_factor_schedule(annealing_factor=annealing_factor)
    prior = partial_prior(scale=prior_scale(noise_schedule, schedule, t_start), device=device)
    prior_samples = prior.sample(num_samples)
    if annealing_factor_score is None:
        annealing_factor_score = annealing_factor
    samples, _, num_unique_idxs, sde_terms, acceptance_rate_list = weighted_sde_integrator.integrate_sde(
        x1=prior_samples, energy_function=energy_function, inverse_temperature=inverse_temp,
        annealing_factor_schedule=schedule, annealing_factor_score=annealing_factor_score)
    if not return_logweights:
        return samples, num_unique_idxs, sde_terms, acceptance_rate_list
    # re-integrate without resampling to get log-weights; fewer samples are enough (:281-289)
    samples_not_resampled, logweights, _, _, _ = weighted_sde_integrator.integrate_sde(
        x1=prior_samples[:inference_batch_size], energy_function=energy_function,
        resampling_interval=num_integration_steps + 1, inverse_temperature=inverse_temp,
        annealing_factor_schedule=schedule, annealing_factor_score=annealing_factor_score)
    return samples, samples_not_resampled, logweights, num_unique_idxs, sde_terms, acceptance_rate_list
