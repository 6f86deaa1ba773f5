"""world_size-2 (gloo, CPU) test of the sharded resampling host logic (pita_b200/distributed.py).

The N>1 path keeps particles block-partitioned over ranks (reference slice: sde_integration.py:227-233)
and only communicates in a resampling step: all-gather of the log-weights, the same global systematic
resample on every rank restricted to the rank's own offspring slots, then an exchange of ancestor rows.
On the GPU the three compute steps are CUDA kernels; here the compute backend is replaced by the CPU
oracle (test infrastructure) so that the *host* logic — shard bounds, slot ranges, the all-gather
exchange, the change count that reproduces len(np.unique(choice)) — is exercised without a GPU.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """CPU stand-in for CudaResampleBackend built on the oracle (checker only)."""

    def __init__(self):
        import pita_oracle as O
        self.O = O

    def softmax_clip(self, logits):
        return self.O.clipped_softmax(logits)

    def systematic(self, w, u0, lo, hi):
        ids = self.O.systematic_indices(w.numpy(), u0)
        n = len(ids)
        prev = np.roll(ids, 1)
        changes = int((ids[lo:hi] != prev[lo:hi]).sum())
        assert n == w.numel()
        return torch.from_numpy(ids[lo:hi].copy()), torch.tensor([changes], dtype=torch.int64)

    def gather_tensor(self, src, ids):
        return src[ids].clone()


def _worker(rank, world, port, n_local, D, seeds, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pita_b200.distributed import ShardedResampler, shard_bounds, world_info
    assert world_info() == (world, rank)
    N = n_local * world
    lo, hi = shard_bounds(N, world, rank)
    assert (lo, hi) == (rank * n_local, (rank + 1) * n_local)
    rs = ShardedResampler(n_local, D, "cpu", backend=OracleBackend())
    assert rs.exchange == "allgather" and rs.particle_buffer() is None
    res = {}
    for seed in seeds:
        g = torch.Generator().manual_seed(seed)
        x_full = torch.randn(N, D, generator=g)
        a_full = torch.randn(N, generator=g) * 3.0
        u0 = float(torch.rand(1, dtype=torch.float64, generator=g))
        a_gathered = rs.gather_logweights(a_full[lo:hi].clone())
        assert torch.equal(a_gathered, a_full)
        x_new, changes = rs.resample(x_full[lo:hi].clone(), a_gathered, u0)
        res[seed] = (x_new.numpy(), int(changes.item()))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **{"x%d" % s: v[0] for s, v in res.items()},
             **{"c%d" % s: np.int64(v[1]) for s, v in res.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_resample_matches_global(tmp_path, world):
    import pita_oracle as O
    n_local, D, seeds = 96, 39, [0, 1, 2]
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_local, D, seeds, str(tmp_path)), nprocs=world, join=True)
    N = n_local * world
    for seed in seeds:
        g = torch.Generator().manual_seed(seed)
        x_full = torch.randn(N, D, generator=g)
        a_full = torch.randn(N, generator=g) * 3.0
        u0 = float(torch.rand(1, dtype=torch.float64, generator=g))
        ids = O.systematic_resample(a_full, u0)  # reference semantics on the whole population
        want = x_full[torch.from_numpy(ids)].numpy()
        got = np.concatenate([np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))["x%d" % seed] for r in range(world)])
        assert np.array_equal(got, want), "sharded resample differs from the global one"
        changes = [int(np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))["c%d" % seed]) for r in range(world)]
        assert len(set(changes)) == 1, "all ranks must hold the all-reduced change count"
        assert max(changes[0], 1) == len(np.unique(ids))  # sde_integration.py:295


def test_shard_bounds_rejects_ragged():
    from pita_b200.distributed import shard_bounds
    assert shard_bounds(8, 2, 1) == (4, 8)
    with pytest.raises(ValueError):
        shard_bounds(9, 2, 0)
