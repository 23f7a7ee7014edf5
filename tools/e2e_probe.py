#!/usr/bin/env python
"""End-to-end pair hashing from pinned host memory for one or more library builds.  usage: e2e_probe.py [lib.so ...]"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cuzk_b200 import lib as cl
n = 1_000_000
for path in (sys.argv[1:] or [cl.LIB_PATH]):
    L = cl.Lib(path); L.check(L.cuzk_init(0), "init")
    dl = torch.empty((n, 4), dtype=torch.int64, device='cuda'); dr = torch.empty_like(dl)
    L.cuzk_synth_elements(dl.data_ptr(), n, 1, 0, 1, None); L.cuzk_synth_elements(dr.data_ptr(), n, 2, 0, 1, None)
    hl = dl.cpu().pin_memory(); hr = dr.cpu().pin_memory(); ho = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    for _ in range(3): L.cuzk_poseidon_hash_pairs(hl.data_ptr(), hr.data_ptr(), ho.data_ptr(), n, 1, None)
    t0 = time.perf_counter()
    for _ in range(20): L.cuzk_poseidon_hash_pairs(hl.data_ptr(), hr.data_ptr(), ho.data_ptr(), n, 1, None)
    dt = (time.perf_counter() - t0) / 20
    print(f"{os.path.basename(path):28s} chunk {os.environ.get('CUZK_CHUNK_PERCENT', '100'):>4s}%  {dt * 1e3:.3f} ms  {n / dt / 1e6:.1f} M/s")
    L.cuzk_shutdown()
