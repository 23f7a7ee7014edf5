#!/usr/bin/env python
"""GPU probe: where the time of the 8-ary 2^k-leaf build goes (per-level timings)."""
import os, sys, json, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cuzk_b200 import api, lib as cl
api.initialize(0)
L = cl.get_lib(); dev = torch.device("cuda", 0)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 26
n = 1 << k
leaves = torch.empty((n, 4), dtype=torch.int64, device=dev)
L.cuzk_synth_u64_leaves(leaves.data_ptr(), n, 4, 0, None)
def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
cur, m = leaves, n
for lvl in range(1, 10):
    cnt = max(1, -(-m // 8))
    # pad to full groups
    cnt_slots = cnt
    out = torch.empty((cnt_slots, 4), dtype=torch.int64, device=dev)
    ms = timed(lambda: L.check(L.cuzk_merkle_subtree_roots(cur.data_ptr(), m, 8, 1, cnt_slots, out.data_ptr(), 0, None), "sr"))
    print(f"level {lvl}: {m} -> {cnt_slots} nodes  {ms:.3f} ms  {cnt_slots*4/ms/1e3:.1f} Mperm/s")
    cur, m = out, cnt_slots
    if m == 1: break
root = torch.empty((1, 4), dtype=torch.int64, device=dev)
span_h = 0; p = 1
while p < n: p *= 8; span_h += 1
ms = timed(lambda: L.check(L.cuzk_merkle_subtree_roots(leaves.data_ptr(), n, 8, span_h, 1, root.data_ptr(), 0, None), "sr"))
print(f"whole tree via subtree_roots(height={span_h}): {ms:.2f} ms")
tot = api.total_nodes(n, 8)
lv = torch.empty((tot, 4), dtype=torch.int64, device=dev)
ms = timed(lambda: L.check(L.cuzk_merkle_build(leaves.data_ptr(), n, 8, lv.data_ptr(), 0, None), "b"))
print(f"cuzk_merkle_build all levels: {ms:.2f} ms; roots equal: {bool((lv[-1]==root[0]).all())}")
