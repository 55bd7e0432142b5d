#!/usr/bin/env python
"""Phase timing of the post-processing kernels on real random-net maps (batch 32)."""
import ctypes, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _b200_alias
_b200_alias.load_package()
import torch
engine = importlib.import_module(_b200_alias.PKG + ".engine")
nat = importlib.import_module(_b200_alias.PKG + "._native")
syn = importlib.import_module(_b200_alias.PKG + ".synthetic")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
eng = engine.PoseEngine(syn.he_state_arrays(1234), 0, "bf16", batch_cap=B)
x = (torch.rand((B, 3, 368, 368), generator=torch.Generator().manual_seed(1)) - 0.5).cuda()
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    eng.infer_async(x.data_ptr(), True, B, 368, 368, 0.1, st)
torch.cuda.synchronize()
dbg = (ctypes.c_ulonglong * 16)()
nat.lib().b200pose_post_debug(eng.post._h, dbg, 16, 1)
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
e[0].record(); eng.infer_async(x.data_ptr(), True, B, 368, 368, 0.1, st); e[1].record(); torch.cuda.synchronize()
nat.lib().b200pose_post_debug(eng.post._h, dbg, 16, 0)
print("step %.3f ms; candidates per batch %d, largest limb %d" % (e[0].elapsed_time(e[1]), dbg[4], dbg[3]))
res = eng.fetch()
print("humans per image:", [len(r) for r in res][:8], "status", eng.post.status(0))
