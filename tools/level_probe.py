"""Times single Merkle levels and sponges on the cooperative kernels (us per launch, median of 9), to see how a level's
duration depends on its arity, on the share of real nodes and on the layout."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cuzk_b200 import api
from cuzk_b200.lib import get_lib

lib = get_lib()
api.initialize(0)
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(3)
x = torch.randint(0, 2**62, (1 << 17, 4), dtype=torch.int64, generator=g)
x[:, 3] &= (1 << 60) - 1
x = x.to(dev)
out = torch.empty((1 << 16, 4), dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=9):
    ts = []
    for i in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1) * 1000)
    return round(float(np.median(ts)), 1)


def level(n, arity, count):
    return timed(lambda: lib.check(lib.cuzk_merkle_subtree_roots(x.data_ptr(), n, arity, 1, count, out.data_ptr(), 0, st), "level"))


def sponge(width, n):
    return timed(lambda: lib.check(lib.cuzk_poseidon_sponge(x.data_ptr(), width, 3, out.data_ptr(), n, 0, st), "sponge"))


def pairs(n):
    return timed(lambda: lib.check(lib.cuzk_poseidon_hash_pairs(x.data_ptr(), x[n:].data_ptr(), out.data_ptr(), n, 0, st), "pairs"))


modes = {"default": (2368, 4736), "wide16": (1 << 20, 1 << 20), "narrow8": (0, 1 << 20), "one_thread": (0, 0)}
for name, (wide, cmax) in modes.items():
    lib.cuzk_debug_set_coop_wide_max(wide)
    lib.cuzk_debug_set_coop_max(cmax)
    row = {"mode": name}
    for arity, count, n in ((2, 4096, 8192), (2, 4096, 6250), (2, 4096, 8191), (2, 3125, 6250), (2, 2048, 4096), (4, 4096, 16384), (4, 4096, 12000),
                            (4, 2048, 8192), (4, 1024, 4096), (8, 2048, 16384), (8, 1024, 8192), (8, 4096, 32768)):
        row[f"level_a{arity}_c{count}_n{n}"] = level(n, arity, count)
    for width, n in ((2, 4096), (4, 4096), (8, 4096), (4, 2048), (8, 2048), (4, 1024), (8, 1024), (3, 4096)):
        row[f"sponge_w{width}_n{n}"] = sponge(width, n)
    for n in (1024, 2048, 4096):
        row[f"pairs_n{n}"] = pairs(n)
    print(json.dumps(row), flush=True)
