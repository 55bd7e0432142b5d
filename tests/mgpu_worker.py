"""torchrun worker of tests/test_gpu.py::test_two_rank_sharded_batch_equals_single_gpu: rank r takes its shard_range
slice of a fixed 64-frame batch, receives the weights by the one NCCL broadcast, and writes its per-frame person rows."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _b200_alias  # noqa: E402


def main(out_dir):
    _b200_alias.load_package()
    engine = importlib.import_module(_b200_alias.PKG + ".engine")
    dist_mod = importlib.import_module(_b200_alias.PKG + ".distributed")
    syn = importlib.import_module(_b200_alias.PKG + ".synthetic")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    arrays = dist_mod.broadcast_state_arrays(syn.he_state_arrays(1234) if rank == 0 else None, device=dev)
    frames = np.random.RandomState(77).randint(0, 256, (64, 184, 184, 3)).astype(np.uint8)
    lo, hi = dist_mod.shard_range(len(frames), rank, world)
    pe = engine.PoseEngine(arrays, local, mode="bf16", batch_cap=32, peak_cap=1024, human_cap=1024)
    pe.infer_batch(frames[lo:hi])
    rows = pe.fetch_arrays()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **{str(lo + i): r for i, r in enumerate(rows)})
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
