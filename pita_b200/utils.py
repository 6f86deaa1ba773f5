"""`sample_cat_sys` with the reference's signature (models/components/utils.py:111-120).

The uniform offset is drawn exactly as the reference draws it (`torch.rand(1, float64)` on the CPU
generator) so a seeded run consumes the same random stream; softmax/clip, the fp64 prefix sum, the
search and the unique count run on the GPU.  Returns a device int64 tensor (the reference returns a
host NumPy array after a device->host copy of the bins)."""
from __future__ import annotations

import torch

from . import ops


def draw_u0() -> float:
    return float(torch.rand(size=(1,), dtype=torch.float64))


def sample_cat_sys(bs, logits, u0=None, return_unique=False):
    if logits.shape[-1] != bs:
        raise RuntimeError("sample_cat_sys: bs must equal logits.shape[-1]")
    if u0 is None:
        u0 = draw_u0()
    w = ops.softmax_clip(logits)
    ids, changes = ops.resample_systematic(w, u0, count_changes=return_unique)
    if return_unique:
        return ids, None, changes
    return ids, None


def num_unique_from_changes(changes: int) -> int:
    """cyclic value changes of a rotated non-decreasing index sequence == number of distinct indices (0 -> 1)."""
    return max(int(changes), 1)
