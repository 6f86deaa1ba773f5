"""Worker of tests/test_gpu_multi.py — one process per GPU (torchrun, NCCL).  Checks, at world size W:
  1. ShardedResampler (p2p peer-memory gather over NVLink, and the NCCL all-gather fallback) reproduces the global
     systematic resample of the oracle bit for bit, rank slices concatenated in rank order;
  2. WeightedSDEIntegrator.integrate_sde with particles sharded over W ranks reproduces the world-size-1 oracle loop
     (injected noise / offsets; chunk divides the shard so chunk-local quantiles coincide, sde_integration.py:227-233).
Test infrastructure: imports the oracle as the checker."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import pita_oracle as O  # noqa: E402
from helpers import assert_close, make_net  # noqa: E402


def check_resampler(world, rank, dev, exchange):
    from pita_b200.distributed import ShardedResampler, shard_bounds
    n_local, D = 4096, 165
    N = n_local * world
    lo, hi = shard_bounds(N, world, rank)
    rs = ShardedResampler(n_local, D, dev, exchange=exchange)
    assert rs.exchange == exchange, (rs.exchange, getattr(rs, "_why", None))
    for seed in (0, 1, 2):
        g = torch.Generator().manual_seed(seed)
        x_full = torch.randn(N, D, generator=g)
        a_full = torch.randn(N, generator=g) * 3.0
        u0 = float(torch.rand(1, dtype=torch.float64, generator=g))
        a_g = rs.gather_logweights(a_full[lo:hi].to(dev))
        assert torch.equal(a_g.cpu(), a_full)
        buf = rs.particle_buffer()
        xl = x_full[lo:hi].to(dev)
        if buf is not None:
            buf.copy_(xl)
            xl = buf
        x_new, changes = rs.resample(xl, a_g, u0)
        # indices are bit-exact GIVEN the weights (SURVEY §8d gate 1): take the device's clipped softmax, as the 1-GPU tests do
        from pita_b200 import ops
        ids = O.systematic_indices(ops.softmax_clip(a_g).cpu().numpy(), u0)
        want = x_full[torch.from_numpy(ids[lo:hi])]
        assert torch.equal(x_new.cpu(), want), "rank %d: sharded resample (%s) differs from the global one" % (rank, exchange)
        assert max(int(changes.item()), 1) == len(np.unique(ids))
    return True


def check_loop(world, rank, dev, exchange):
    from pita_b200.annealing_factor_schedules import ConstantAnnealingFactorSchedule
    from pita_b200.energy_net import EnergyNet
    from pita_b200.lennardjones_energy import LennardJonesEnergy
    from pita_b200.noise_schedules import ElucidatingNoiseSchedule
    from pita_b200.score_net import ScoreNet
    from pita_b200.sde_integration import WeightedSDEIntegrator
    from pita_b200.sdes import VEReverseSDE
    n, N, S, chunk, time_range = 13, 64, 8, 16, 0.2
    assert (N // world) % chunk == 0
    sdE = O.random_egnn_state(seed=31 + n, dtype=torch.float64, coord_gain=0.3)
    sdS = O.random_egnn_state(seed=32 + n, dtype=torch.float64, coord_gain=0.3)
    gam, sched = 4.0 / 3.0, O.EDMSchedule(0.05)
    gen = torch.Generator().manual_seed(N)
    scale = float((sched.h(torch.tensor(time_range, dtype=torch.float64)) / gam) ** 0.5)
    x1 = O.centre(O.md_shaped_coords(N, n, seed=N, dtype=torch.float64) + scale * torch.randn(N, 3 * n, generator=gen, dtype=torch.float64), n)
    noise = {s: torch.randn(N, 3 * n, generator=gen, dtype=torch.float64) for s in range(S)}
    u0 = {s: float(torch.rand(1, generator=gen, dtype=torch.float64)) for s in range(S + 1)}
    cfg = O.LoopConfig(n=n, steps=S, chunk=chunk, beta=0.9, resampling_interval=1, time_range=time_range)
    cursor = {}

    def noise_fn(step, xc):
        lo = cursor.get(step, 0)
        cursor[step] = lo + xc.shape[0]
        return noise[step][lo:lo + xc.shape[0]]

    x_ref, logw_ref, uniq_ref = O.integrate(sdE, sdS, sched, O.ConstGamma(gam), cfg, x1, noise_fn, lambda s: u0[s])
    sde = VEReverseSDE(ElucidatingNoiseSchedule(0.05, 80.0, 7.0), energy_net=EnergyNet(make_net(n, sdE, dev)),
                       score_net=ScoreNet(make_net(n, sdS, dev)), debias_inference=True)
    integ = WeightedSDEIntegrator(sde=sde, num_integration_steps=S, lightning_module=None, batch_size=chunk, num_negative_time_steps=0,
                                  post_mcmc_steps=0, start_resampling_step=0, end_resampling_step=10 ** 9, resampling_interval=1,
                                  time_range=time_range, exchange=exchange)
    per = N // world
    integ.noise_fn = lambda step, x: noise[step][rank * per:(rank + 1) * per].float().to(dev)
    integ.u0_fn = lambda step: u0[step]
    tgt = LennardJonesEnergy(dimensionality=3 * n, n_particles=n)
    x, logw, uniq, _, _ = integ.integrate_sde(x1.float().to(dev), tgt, ConstantAnnealingFactorSchedule(gam), inverse_temperature=0.9)
    assert x.shape == (N, 3 * n) and logw.shape == (S, N)
    assert list(uniq) == list(uniq_ref), (uniq, uniq_ref)
    assert_close(x, x_ref, "x_final (world %d)" % world, rtol=1e-3)
    assert integ._resampler.exchange == exchange
    return True


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = {}
    for exchange in ("allgather", "p2p"):
        try:
            res["resampler_" + exchange] = check_resampler(world, rank, dev, exchange)
            res["loop_" + exchange] = check_loop(world, rank, dev, exchange)
        except Exception as exc:  # noqa: BLE001
            res["error_" + exchange] = repr(exc)[:400]
    flag = torch.tensor([int(any(k.startswith("error") for k in res))], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print("MULTI_GPU_RESULT " + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    main()
