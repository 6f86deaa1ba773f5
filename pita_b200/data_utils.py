"""`remove_mean` with the reference's signature (utils/data_utils.py:4-26), on the CUDA kernel."""
from . import ops


def remove_mean(samples, n_particles, n_dimensions):
    if n_dimensions != 3:
        raise NotImplementedError("pita_b200.remove_mean is built for 3 spatial dimensions")
    return ops.remove_mean(samples, n_particles).to(samples.dtype)
