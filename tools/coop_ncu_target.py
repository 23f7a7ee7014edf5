#!/usr/bin/env python
"""ncu target for the cooperative kernels: 1024 pair hashes (one warp per SM sub-partition), launched a few times.
  ncu --set full --clock-control none --import-source on -k regex:coop_hash_pairs -s 2 -c 1 -o gpurun_out/prof_coop python tools/coop_ncu_target.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cuzk_b200 import lib as cl  # noqa: E402

L = cl.Lib(sys.argv[2]) if len(sys.argv) > 2 else cl.get_lib()
L.check(L.cuzk_init(0), "init")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
if len(sys.argv) > 3:
    L.cuzk_debug_set_coop_max(1 << 30)
    L.cuzk_debug_set_coop_wide_max(int(sys.argv[3]))
l = torch.empty((n, 4), dtype=torch.int64, device="cuda")
r = torch.empty_like(l)
o = torch.empty_like(l)
L.cuzk_synth_elements(l.data_ptr(), n, 1, 0, 1, None)
L.cuzk_synth_elements(r.data_ptr(), n, 2, 0, 1, None)
for _ in range(4):
    L.check(L.cuzk_poseidon_hash_pairs(l.data_ptr(), r.data_ptr(), o.data_ptr(), n, 0, None), "hash")
torch.cuda.synchronize()
print("ok", int(o[0, 0]))
