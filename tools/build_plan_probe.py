"""Sweep of the full-tree build plan (groups x streams x in-group cooperative cap) on one GPU.
Prints one JSON line per plan: build times (ms, median of 7) of the BASELINE trees, roots checked against the first plan."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from cuzk_b200 import api
from cuzk_b200.lib import get_lib

lib = get_lib()
api.initialize(0)
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(7)
NOCAP = (1 << 64) - 1
trees = {"binary_50k": (50000, 2), "quaternary_2p20": (1 << 20, 4), "octary_2p23": (1 << 23, 8), "octary_2p21": (1 << 21, 8), "binary_2p20": (1 << 20, 2)}
leaves = {}
for name, (n, a) in trees.items():
    x = torch.randint(0, 2**62, (n, 4), dtype=torch.int64, generator=g)
    x[:, 3] &= (1 << 60) - 1
    leaves[name] = x.to(dev)


def build_ms(name, reps=int(__import__("os").environ.get("REPS", "7"))):
    n, a = trees[name]
    total = lib.cuzk_merkle_total_nodes(n, a)
    out = torch.empty((total, 4), dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    ts = []
    for i in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        lib.check(lib.cuzk_merkle_build(leaves[name].data_ptr(), n, a, out.data_ptr(), 0, st), "build")
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out[-1].cpu().numpy().tobytes().hex()


plans = [(1, 1, NOCAP), (4, 4, NOCAP), (4, 4, 256), (8, 8, NOCAP), (8, 8, 256), (8, 8, 1184), (16, 8, 256), (16, 16, NOCAP), (16, 16, 256),
         (16, 16, 1184), (16, 16, 0), (32, 16, 256), (64, 16, 256), (8, 4, 256), (16, 4, 256)]
if len(sys.argv) > 1:
    plans = [tuple(int(v) if int(v) >= 0 else NOCAP for v in a.split(",")) for a in sys.argv[1:]]
roots = {}
for groups, streams, cap in plans:
    lib.cuzk_debug_set_build_plan(groups, streams, cap)
    row = {"groups": groups, "streams": streams, "group_coop_max": None if cap == NOCAP else cap}
    for name in trees:
        ms, root = build_ms(name)
        roots.setdefault(name, root)
        assert roots[name] == root, (name, groups, streams, cap)
        row[name] = round(ms, 3)
    print(json.dumps(row), flush=True)
