"""The reference harness shape (245 synchronous host calls of <= 4096 pair hashes, poseidon_cuda_benchmarks.cpp:63-117) with the
direct small-call path on and off, pinned and pageable caller memory: M hashes/s and us per call."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cuzk_b200 import api
from cuzk_b200.lib import get_lib

L = get_lib()
api.initialize(0)
total, batch = 1_000_000, 4096
rng = np.random.default_rng(5)
l = rng.integers(0, 2**62, (total, 4), dtype=np.uint64)
r = rng.integers(0, 2**62, (total, 4), dtype=np.uint64)
out = np.zeros_like(l)
pl, pr = torch.from_numpy(l.view(np.int64)).pin_memory(), torch.from_numpy(r.view(np.int64)).pin_memory()
po = torch.empty_like(pl).pin_memory()


def run(lp, rp, op):
    t0 = time.perf_counter()
    calls = 0
    for at in range(0, total, batch):
        m = min(batch, total - at)
        L.check(L.cuzk_poseidon_hash_pairs(lp + 32 * at, rp + 32 * at, op + 32 * at, m, 1, None), "pairs")
        calls += 1
    dt = time.perf_counter() - t0
    return {"mhash_per_s": round(total / dt / 1e6, 2), "us_per_call": round(dt / calls * 1e6, 1)}


for batch in (4096, 1024, 8192):
    for direct in (1 << 20, 0):
        L.cuzk_debug_set_direct_max(direct)
        row = {"batch": batch, "direct_max": direct}
        for name, ptrs in (("pinned", (pl.data_ptr(), pr.data_ptr(), po.data_ptr())), ("pageable", (l.ctypes.data, r.ctypes.data, out.ctypes.data))):
            run(*ptrs)
            row[name] = run(*ptrs)
        print(json.dumps(row), flush=True)
assert (out == po.numpy().view(np.uint64)).all()
