#!/usr/bin/env python
"""ncu target for the Merkle kernels: one 8-ary 2^24-leaf subtree-root pass (merkle_level_kernel, coop_merkle_level_kernel),
one 4-ary 2^22-leaf full build, one 2^18-proof 4-ary batch verification.  Usage under ncu:
  ncu --set full --clock-control none -k regex:merkle_ -o gpurun_out/prof_merkle python tools/merkle_ncu_target.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cuzk_b200 import api, lib as cl  # noqa: E402

api.initialize(0)
L = cl.get_lib()
dev = torch.device("cuda", 0)
n = 1 << 24
leaves = torch.empty((n, 4), dtype=torch.int64, device=dev)
L.check(L.cuzk_synth_u64_leaves(leaves.data_ptr(), n, 4, 0, None), "synth")
root = torch.empty((1, 4), dtype=torch.int64, device=dev)
L.check(L.cuzk_merkle_subtree_roots(leaves.data_ptr(), n, 8, 8, 1, root.data_ptr(), 0, None), "subtree_roots")
m = 1 << 22
t = api.CudaNaryMerkleTree(leaves[:m], arity=4)
q = 1 << 18
idx = torch.arange(q, dtype=torch.int64, device=dev) * 16
pb = t.generate_batch_proofs(idx)
res = t.verify_batch_proofs(pb, leaves[idx].contiguous())
torch.cuda.synchronize()
print("ok", bool(res.all()))
