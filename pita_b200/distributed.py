"""Multi-GPU plumbing for the only step of the path that is not embarrassingly parallel: FK resampling.

Particles are block-partitioned across ranks exactly like the reference's slice
(sde_integration.py:227-233): rank r owns global rows [r*N/W, (r+1)*N/W).  The reference all-gathers x, a
and six diagnostic tensors EVERY step (:248-258); here particles stay sharded and only a resampling step
communicates:
  1. all-gather of the log-weights a (4 B/particle, NCCL over NVLink);
  2. every rank evaluates the same global clipped-softmax + fp64 prefix sum and searches the ancestors of
     ITS OWN offspring slots (pita_softmax_clip / pita_resample_systematic);
  3. ancestor rows are fetched by ONE gather kernel that reads peer HBM directly over NVLink
     (pita_gather_rows with per-rank base pointers from torch symmetric memory) — or, when symmetric
     memory is unavailable, from an NCCL all-gather of x;
  4. the number of distinct ancestors is the all-reduced count of index changes.
The compute backend is injectable so the host logic can be exercised on CPU/gloo by the tests (which
inject the oracle); the default backend is the CUDA library and raises if it is missing.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def world_info(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    if n_total % world != 0:
        raise ValueError("number of particles (%d) must be divisible by the number of ranks (%d); the reference "
                         "silently drops the remainder (sde_integration.py:227)" % (n_total, world))
    per = n_total // world
    return rank * per, (rank + 1) * per


class CudaResampleBackend:
    """Default compute backend: the C-ABI kernels."""

    def softmax_clip(self, logits):
        from . import ops
        return ops.softmax_clip(logits)

    def systematic(self, w, u0, lo, hi):
        from . import ops
        ids, ch = ops.resample_systematic(w, u0, lo, hi, count_changes=True)
        return ids, ch

    def gather_ptrs(self, src_ptrs, rows_per_rank, ids, row_floats, out):
        from . import ops
        return ops.gather_rows(src_ptrs, rows_per_rank, ids, row_floats, out=out)

    def gather_tensor(self, src, ids):
        from . import ops
        return ops.gather_rows([src.data_ptr()], src.shape[0], ids, src.shape[1])


class ShardedResampler:
    """Systematic resampling of N = W * n_local particles kept sharded over W ranks."""

    def __init__(self, n_local: int, row_floats: int, device, group=None, exchange: str = "auto", backend=None):
        self.group = group
        self.world, self.rank = world_info(group)
        self.n_local, self.row_floats, self.device = n_local, row_floats, torch.device(device)
        self.n_total = n_local * self.world
        self.backend = backend or CudaResampleBackend()
        self.exchange = exchange
        self._symm = None
        self._bufs = None
        self._cur = 0
        if self.world > 1 and exchange in ("auto", "p2p") and self.device.type == "cuda":
            try:
                self._setup_symmetric()
                self.exchange = "p2p"
            except Exception as exc:  # noqa: BLE001
                if exchange == "p2p":
                    raise
                self._symm = None
                self.exchange = "allgather"
                self._why = repr(exc)
        elif self.world > 1:
            self.exchange = "allgather"
        else:
            self.exchange = "local"

    # -- symmetric (peer-mapped) particle buffers: two, so that peers can still read the old generation
    def _setup_symmetric(self):
        import torch.distributed._symmetric_memory as symm_mem
        grp = self.group if self.group is not None else dist.group.WORLD
        self._bufs, self._symm = [], []
        for _ in range(2):
            t = symm_mem.empty(self.n_local, self.row_floats, dtype=torch.float32, device=self.device)
            self._bufs.append(t)
            self._symm.append(symm_mem.rendezvous(t, grp))

    def particle_buffer(self) -> Optional[torch.Tensor]:
        """Buffer the caller should write the pre-resampling particles into (p2p mode), else None."""
        return self._bufs[self._cur] if self.exchange == "p2p" else None

    def gather_logweights(self, a_local: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return a_local
        a_full = torch.empty(self.n_total, device=a_local.device, dtype=a_local.dtype)
        dist.all_gather_into_tensor(a_full, a_local.contiguous(), group=self.group)
        return a_full

    def resample(self, x_local: torch.Tensor, a_full: torch.Tensor, u0: float) -> Tuple[torch.Tensor, torch.Tensor]:
        """x_local: this rank's [n_local, D] rows (in p2p mode: the tensor returned by particle_buffer()).
        a_full: the gathered (and possibly clamped) log-weights of all N particles.  Returns the rank's
        resampled rows and a 1-element int64 tensor holding the global number of cyclic index changes
        (== number of distinct ancestors, 0 meaning 1) — left on the device so the loop never syncs."""
        lo, hi = shard_bounds(self.n_total, self.world, self.rank)
        w = self.backend.softmax_clip(a_full)
        ids, changes = self.backend.systematic(w, u0, lo, hi)
        if self.exchange == "p2p":
            cur = self._cur
            if x_local.data_ptr() != self._bufs[cur].data_ptr():
                self._bufs[cur].copy_(x_local)
            hdl = self._symm[cur]
            hdl.barrier()  # every rank's pre-resampling rows are in place
            out = self._bufs[1 - cur]
            x_new = self.backend.gather_ptrs([int(p) for p in hdl.buffer_ptrs], self.n_local, ids, self.row_floats, out)
            hdl.barrier()  # peers are done reading this generation
            self._cur = 1 - cur
        elif self.exchange == "allgather":
            x_full = torch.empty(self.n_total, self.row_floats, device=x_local.device, dtype=x_local.dtype)
            dist.all_gather_into_tensor(x_full, x_local.contiguous(), group=self.group)
            x_new = self.backend.gather_tensor(x_full, ids)
        else:
            x_new = self.backend.gather_tensor(x_local.contiguous(), ids)
        if self.world > 1:
            dist.all_reduce(changes, op=dist.ReduceOp.SUM, group=self.group)
        return x_new, changes
