#!/usr/bin/env python
"""Tuning helper: build kernel variants with extra -D flags into scratch .so files and time the pair-hash kernel
(1M pairs, device resident), checking a sample against the oracle.  usage: quick_bench.py "NAME=-DFOO -DBAR=2" ..."""
import ctypes as C, os, subprocess, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cuzk_b200 import lib as cl
from oracle_lib import Oracle, synth_elements
if sys.argv[1] == "build":
    for spec in sys.argv[2:]:
        name, _, flags = spec.partition("=")
        out = os.path.join(ROOT, "build", "variants", f"variant_{name}.so")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        cu = [s for s in cl._sources() if s.endswith(".cu")]
        cmd = ["nvcc"] + cl.NVCC_FLAGS + flags.split() + ["-Xptxas", "-v", "-o", out] + cu
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        lines = r.stderr.splitlines()
        for i, l in enumerate(lines):
            if "hash_pairs_kernel" in l and "Compiling" in l:
                print(name, "|", lines[i + 1].strip(), "|", lines[i + 2].strip())
    sys.exit(0)
oracle = Oracle()
n = 1_000_000
res = {}
for path in sys.argv[1:]:
    L = cl.Lib(path)
    L.check(L.cuzk_init(0), "init")
    l = torch.empty((n, 4), dtype=torch.int64, device="cuda"); r = torch.empty_like(l); o = torch.empty_like(l)
    L.cuzk_synth_elements(l.data_ptr(), n, 1, 0, 1, None); L.cuzk_synth_elements(r.data_ptr(), n, 2, 0, 1, None)
    for _ in range(3): L.cuzk_poseidon_hash_pairs(l.data_ptr(), r.data_ptr(), o.data_ptr(), n, 0, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): L.cuzk_poseidon_hash_pairs(l.data_ptr(), r.data_ptr(), o.data_ptr(), n, 0, None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    got = o[:256].cpu().numpy().view(np.uint64)
    ok = bool((got == oracle.hash_pairs(synth_elements(1, 256), synth_elements(2, 256))).all())
    print(f"{os.path.basename(path):40s} {ms:8.3f} ms  {n / ms / 1e3:8.1f} Mhash/s  ok={ok}")
    L.cuzk_shutdown()
