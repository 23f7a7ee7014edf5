#!/usr/bin/env python
"""Tuning helper for the cooperative kernels: `build NAME=-DFLAGS ...` compiles library variants into build/variants/,
`run lib.so ...` times pair-hash launches of a few sizes on both layouts of each (one JSON line per library)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cuzk_b200 import lib as cl

if sys.argv[1] == "build":
    for spec in sys.argv[2:]:
        name, _, flags = spec.partition("=")
        out = os.path.join(ROOT, "build", "variants", f"coop_{name}.so")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        cu = [s for s in cl._sources() if s.endswith(".cu")]
        r = subprocess.run(["nvcc"] + cl.NVCC_FLAGS + flags.split() + ["-Xptxas", "-v", "-o", out] + cu, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        lines = r.stderr.splitlines()
        for i, l in enumerate(lines):
            if "coop_hash_pairs_kernel" in l and "Compiling" in l:
                print(name, "|", "Wide16" if "Wide16" in l else "Narrow8", "|", lines[i + 2].strip())
    sys.exit(0)

import torch


def timed(fn, reps=30):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for path in sys.argv[2:]:
    L = cl.Lib(path)
    L.check(L.cuzk_init(0), "init")
    nmax = 8192
    l = torch.empty((nmax, 4), dtype=torch.int64, device="cuda")
    r = torch.empty_like(l)
    o = torch.empty_like(l)
    L.cuzk_synth_elements(l.data_ptr(), nmax, 1, 0, 1, None)
    L.cuzk_synth_elements(r.data_ptr(), nmax, 2, 0, 1, None)
    row = {"lib": os.path.basename(path)}
    for name, wm in (("wide16", 1 << 30), ("narrow8", 0)):
        L.cuzk_debug_set_coop_max(1 << 30)
        L.cuzk_debug_set_coop_wide_max(wm)
        row[name] = {n: round(1e3 * timed(lambda: L.cuzk_poseidon_hash_pairs(l.data_ptr(), r.data_ptr(), o.data_ptr(), n, 0, None)), 1)
                     for n in (592, 1184, 2368, 4096, 6144, 8192)}
    # how often a unit falls back to the exact path: 2^20 pair hashes per layout
    big = 1 << 20
    bl = torch.empty((big, 4), dtype=torch.int64, device="cuda")
    br = torch.empty_like(bl)
    bo = torch.empty_like(bl)
    L.cuzk_synth_elements(bl.data_ptr(), big, 11, 0, 1, None)
    L.cuzk_synth_elements(br.data_ptr(), big, 12, 0, 1, None)
    row["fallbacks_per_2p20"] = {}
    for name, wm in (("wide16", 1 << 30), ("narrow8", 0), ("one_thread", None)):
        L.cuzk_debug_set_coop_max(0 if wm is None else 1 << 30)
        L.cuzk_debug_set_coop_wide_max(wm or 0)
        torch.cuda.synchronize()
        before = L.cuzk_debug_fallback_count()
        L.check(L.cuzk_poseidon_hash_pairs(bl.data_ptr(), br.data_ptr(), bo.data_ptr(), big, 0, None), "pairs")
        torch.cuda.synchronize()
        row["fallbacks_per_2p20"][name] = L.cuzk_debug_fallback_count() - before
        row.setdefault("digest", {})[name] = int(bo.sum().item()) & 0xFFFFFFFF
    print(json.dumps(row))
    L.cuzk_shutdown()
