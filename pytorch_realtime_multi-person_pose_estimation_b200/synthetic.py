"""Seeded synthetic weights / inputs for benchmarks and smoke runs (no checkpoints exist offline):
He-normal weights N(0, 2/fan_in), biases U(-0.1, 0.1), in the reference state_dict order.  The reference's own
init (std=0.01, rtpose_vgg.py:200-206) would make every output ~1e-10 and any accuracy statement vacuous."""
import numpy as np
import torch

from .distributed import tensor_shapes


def he_state_arrays(seed=1234):
    g = torch.Generator().manual_seed(seed)
    out = []
    for shape in tensor_shapes():
        if len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            out.append((torch.randn(shape, generator=g) * float(np.sqrt(2.0 / fan_in))).numpy())
        else:
            out.append((torch.rand(shape, generator=g) * 0.2 - 0.1).numpy())
    return out
