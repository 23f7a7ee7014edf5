#!/usr/bin/env python
"""One-off soak: N seeded pair hashes on the GPU, EVERY output compared with the reference CPU implementation on all host
threads (oracle/_ref when present, else the oracle port).  Prints one JSON line.  usage: soak_compare.py [log2_n=24]"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cuzk_b200 import api, lib as cl
from oracle_lib import Oracle, Ref, have_ref, synth_elements
api.initialize(0); L = cl.get_lib()
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 24)
path = sys.argv[2] if len(sys.argv) > 2 else "default"
if path != "default":
    L.cuzk_debug_set_coop_max(1 << 40)
    L.cuzk_debug_set_coop_wide_max((1 << 40) if path == "wide16" else 0)
impl = Ref() if have_ref() else Oracle()
threads = len(os.sched_getaffinity(0))
before = L.cuzk_debug_fallback_count()
bad = 0
step = 1 << 21
t0 = time.time()
for start in range(0, n, step):
    m = min(step, n - start)
    dl = torch.empty((m, 4), dtype=torch.int64, device="cuda"); dr = torch.empty_like(dl); out = torch.empty_like(dl)
    L.cuzk_synth_elements(dl.data_ptr(), m, 101, start, 1, None); L.cuzk_synth_elements(dr.data_ptr(), m, 102, start, 1, None)
    L.check(L.cuzk_poseidon_hash_pairs(dl.data_ptr(), dr.data_ptr(), out.data_ptr(), m, 0, None), "pairs")
    got = out.cpu().numpy().view(np.uint64)
    want = impl.hash_pairs_mt(synth_elements(101, m, start), synth_elements(102, m, start), threads)
    bad += int((got != want).any(axis=1).sum())
print(json.dumps({"path": path, "pair_hashes_compared": n, "mismatches": bad, "exact_fallbacks_taken": int(L.cuzk_debug_fallback_count() - before),
                  "cpu_impl": "reference" if have_ref() else "port", "cpu_threads": threads, "seconds": round(time.time() - t0, 1)}))
