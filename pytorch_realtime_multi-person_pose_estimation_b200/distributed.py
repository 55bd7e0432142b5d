"""Multi-GPU plumbing: one process per GPU (torchrun), frames sharded on the batch dimension, ONE broadcast of the
packed weights from rank 0 (NCCL over NVLink on GPUs, gloo in CPU tests), no steady-state collective.
Replaces the per-forward parameter broadcast + scatter/gather of torch.nn.DataParallel
(/root/reference/demo/picture_demo.py:47)."""
import numpy as np
import torch
import torch.distributed as dist

from . import _native as nat

_SHAPES = None


def tensor_shapes():
    """Shapes of the 184 state_dict tensors (known on every rank, from the library's layer table)."""
    global _SHAPES
    if _SHAPES is None:
        import ctypes
        dims = (ctypes.c_long * 4)()
        shapes = []
        for i in range(nat.NUM_TENSORS):
            nd = nat.lib().b200pose_net_tensor_shape(i, dims)
            shapes.append(tuple(int(dims[k]) for k in range(nd)))
        _SHAPES = shapes
    return _SHAPES


def shard_range(n_items, rank, world):
    """Contiguous, balanced [lo, hi) slice of n_items frames for `rank`."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_state_arrays(arrays, device="cuda", src=0):
    """Rank `src` passes the 184 float32 arrays, the others None; returns the arrays on every rank.
    One flat fp32 blob (209 MB) in a single broadcast."""
    shapes = tensor_shapes()
    total = sum(int(np.prod(s)) for s in shapes)
    if dist.get_rank() == src:
        flat = torch.from_numpy(np.concatenate([np.asarray(a, np.float32).reshape(-1) for a in arrays]))
        assert flat.numel() == total
        flat = flat.to(device)
    else:
        flat = torch.empty(total, dtype=torch.float32, device=device)
    dist.broadcast(flat, src=src)
    host = flat.cpu().numpy()
    out, off = [], 0
    for s in shapes:
        k = int(np.prod(s))
        out.append(host[off:off + k].reshape(s))
        off += k
    return out


def max_over_ranks(value, device="cuda"):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
