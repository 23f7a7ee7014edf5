#!/usr/bin/env python
"""Latency-bound cases: pair-hash launches of 256 .. 32768 units on the cooperative and on the one-thread kernels (where do
they cross?), the 50K-leaf binary build, 5K-proof verification, the 2^20-leaf 4-ary build.  usage: latency_probe.py [lib.so]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cuzk_b200 import lib as cl


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


L = cl.Lib(sys.argv[1]) if len(sys.argv) > 1 else cl.get_lib()
L.check(L.cuzk_init(0), "init")
default_max = L.cuzk_debug_set_coop_max(0)
L.cuzk_debug_set_coop_max(default_max)
default_wide = L.cuzk_debug_set_coop_wide_max(0)
L.cuzk_debug_set_coop_wide_max(default_wide)
out = {"coop_max_default": default_max, "coop_wide_max_default": default_wide, "pairs_us": {}}
nmax = 32768
l = torch.empty((nmax, 4), dtype=torch.int64, device="cuda")
r = torch.empty_like(l)
o = torch.empty_like(l)
L.cuzk_synth_elements(l.data_ptr(), nmax, 1, 0, 1, None)
L.cuzk_synth_elements(r.data_ptr(), nmax, 2, 0, 1, None)
for n in (256, 592, 1184, 1776, 2368, 3552, 4096, 4736, 5920, 7104, 8192, 12288, 16384, 32768):
    row = {}
    for name, cm, wm in (("wide16", 1 << 30, 1 << 30), ("narrow8", 1 << 30, 0), ("one_thread", 0, 0)):
        L.cuzk_debug_set_coop_max(cm)
        L.cuzk_debug_set_coop_wide_max(wm)
        row[name] = round(1e3 * timed(lambda: L.cuzk_poseidon_hash_pairs(l.data_ptr(), r.data_ptr(), o.data_ptr(), n, 0, None), 30), 1)
    out["pairs_us"][n] = row
L.cuzk_debug_set_coop_max(default_max)
L.cuzk_debug_set_coop_wide_max(default_wide)
for label, n, arity in (("build_50k_binary_ms", 50_000, 2), ("build_2p20_4ary_ms", 1 << 20, 4), ("build_2p23_8ary_ms", 1 << 23, 8)):
    leaves = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    L.cuzk_synth_u64_leaves(leaves.data_ptr(), n, 3, 0, None)
    tot = L.cuzk_merkle_total_nodes(n, arity)
    lv = torch.empty((tot, 4), dtype=torch.int64, device="cuda")
    row = {}
    for name, cm in (("default", default_max), ("one_thread", 0), ("coop_12k", 12288)):
        L.cuzk_debug_set_coop_max(cm)
        row[name] = round(timed(lambda: L.cuzk_merkle_build(leaves.data_ptr(), n, arity, lv.data_ptr(), 0, None), 10), 3)
    out[label] = row
    if n == 50_000:
        q, nlv = 5000, L.cuzk_merkle_num_levels(n, arity) - 1
        idx = (torch.arange(q, dtype=torch.int64, device="cuda") * 7) % n
        sib = torch.empty((q, nlv, arity - 1, 4), dtype=torch.int64, device="cuda")
        pos = torch.empty((q, nlv), dtype=torch.int32, device="cuda")
        res = torch.empty(q, dtype=torch.uint8, device="cuda")
        L.check(L.cuzk_merkle_prove_batch(lv.data_ptr(), n, arity, idx.data_ptr(), q, sib.data_ptr(), pos.data_ptr(), 0, None), "prove")
        vals = leaves[idx].contiguous()
        row = {}
        for name, cm in (("default", default_max), ("one_thread", 0), ("coop_12k", 12288)):
            L.cuzk_debug_set_coop_max(cm)
            row[name] = round(timed(lambda: L.cuzk_merkle_verify_batch(vals.data_ptr(), sib.data_ptr(), pos.data_ptr(), nlv, arity, lv[-1:].data_ptr(),
                                                                       res.data_ptr(), q, 0, None), 10), 3)
            assert bool(res.all())
        out["verify_5k_binary_ms"] = row
L.cuzk_debug_set_coop_max(default_max)
print(json.dumps(out))
L.cuzk_shutdown()
