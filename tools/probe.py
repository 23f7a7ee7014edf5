#!/usr/bin/env python
"""GPU probe used while tuning (not part of the product): pipe microbenchmarks, sustained pair-hash
throughput at several sizes with clock sampling, Merkle level-kernel throughput."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cuzk_b200 import api, lib as cl  # noqa: E402

api.initialize(0)
L = cl.get_lib()
dev = torch.device("cuda", 0)
out = {}

names = ["imad_wide", "imad_lo", "imad_hi", "imad_wide_x_chain", "iadd3_x_chain", "wide+1add", "wide+2add", "wide+3add", "sel", "dfma"]
mb = {}
for v, name in enumerate(names):
    d = C.c_double()
    L.check(L.cuzk_imad_peak(v, 4000, C.byref(d)), "peak")
    mb[name] = {"per_s": d.value, "per_clk_per_sm_at_1965": d.value / 148 / 1.965e9}
out["microbench"] = mb
print(json.dumps(mb, indent=1))


def clocks():
    r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True)
    return r.stdout.strip()


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    # sample clocks while the work is in flight
    c = clocks()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, c


res = []
for n in (1 << 16, 1 << 18, 1_000_000, 1 << 22, 1 << 24):
    l = torch.empty((n, 4), dtype=torch.int64, device=dev)
    r = torch.empty((n, 4), dtype=torch.int64, device=dev)
    o = torch.empty((n, 4), dtype=torch.int64, device=dev)
    L.cuzk_synth_elements(l.data_ptr(), n, 1, 0, 1, None)
    L.cuzk_synth_elements(r.data_ptr(), n, 2, 0, 1, None)
    reps = max(2, min(50, (1 << 25) // n))
    ms, c = timed(lambda: L.cuzk_poseidon_hash_pairs(l.data_ptr(), r.data_ptr(), o.data_ptr(), n, 0, None), reps)
    res.append({"pairs": n, "ms": ms, "Mhash_s": n / ms / 1e3, "clocks": c, "reps": reps})
    print(res[-1])
    del l, r, o
out["pair_hash_sweep"] = res

# sponge width 8 (Merkle node, arity 8) throughput
for width in (2, 4, 8):
    n = 1 << 21
    x = torch.empty((n * width, 4), dtype=torch.int64, device=dev)
    o = torch.empty((n, 4), dtype=torch.int64, device=dev)
    L.cuzk_synth_elements(x.data_ptr(), n * width, 5, 0, 1, None)
    ms, c = timed(lambda: L.cuzk_poseidon_sponge(x.data_ptr(), width, 3, o.data_ptr(), n, 0, None), 3)
    perms = n * ((width + 1) // 2)
    print({"sponge_width": width, "n": n, "ms": ms, "Mperm_s": perms / ms / 1e3, "clocks": c})
    out[f"sponge_w{width}"] = {"n": n, "ms": ms, "Mperm_s": perms / ms / 1e3, "clocks": c}
    del x, o

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
