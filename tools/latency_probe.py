#!/usr/bin/env python
"""Latency-bound cases for kernel variants: 4096 pair hashes per launch, 50K-leaf binary tree build, 5K proof verify.
usage: latency_probe.py lib1.so [lib2.so ...]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cuzk_b200 import lib as cl

def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for path in sys.argv[1:]:
    L = cl.Lib(path)
    L.check(L.cuzk_init(0), "init")
    row = {}
    for n in (1024, 4096, 16384):
        l = torch.empty((n, 4), dtype=torch.int64, device="cuda"); r = torch.empty_like(l); o = torch.empty_like(l)
        L.cuzk_synth_elements(l.data_ptr(), n, 1, 0, 1, None); L.cuzk_synth_elements(r.data_ptr(), n, 2, 0, 1, None)
        row[f"pairs{n}_us"] = 1e3 * timed(lambda: L.cuzk_poseidon_hash_pairs(l.data_ptr(), r.data_ptr(), o.data_ptr(), n, 0, None), 50)
    n = 50_000
    leaves = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    L.cuzk_synth_u64_leaves(leaves.data_ptr(), n, 3, 0, None)
    tot = L.cuzk_merkle_total_nodes(n, 2)
    lv = torch.empty((tot, 4), dtype=torch.int64, device="cuda")
    row["build50k_ms"] = timed(lambda: L.cuzk_merkle_build(leaves.data_ptr(), n, 2, lv.data_ptr(), 0, None), 10)
    print(f"{os.path.basename(path):32s} " + "  ".join(f"{k}={v:8.3f}" for k, v in row.items()))
    L.cuzk_shutdown()
