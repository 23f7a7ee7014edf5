#!/usr/bin/env python
"""Time the 8-ary subtree-root pass with per-level launches, forced fusion and the cost model, for shard sizes 2^23..2^26."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cuzk_b200 import api, lib as cl
api.initialize(0)
L = cl.get_lib(); dev = torch.device("cuda", 0)
out = {}
for lg, height, count in ((23, 7, 4), (24, 8, 1), (25, 8, 2), (26, 9, 1)):
    n = 1 << lg
    leaves = torch.empty((n, 4), dtype=torch.int64, device=dev)
    L.cuzk_synth_u64_leaves(leaves.data_ptr(), n, 4, 0, None)
    roots = torch.empty((count, 4), dtype=torch.int64, device=dev)
    row = {}
    for name, mode in (("per_level", 0), ("fused", 1)):
        L.cuzk_debug_set_fuse(mode)
        f = lambda: L.check(L.cuzk_merkle_subtree_roots(leaves.data_ptr(), n, 8, height, count, roots.data_ptr(), 0, None), "sr")
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): f()
        e1.record(); torch.cuda.synchronize()
        row[name] = e0.elapsed_time(e1) / 3
    out[f"2^{lg}"] = row
    del leaves
L.cuzk_debug_set_fuse(0)
print(json.dumps(out, indent=1))
