#!/usr/bin/env python
"""Stress of the host-buffer pipeline: random batch sizes (so chunk boundaries fall everywhere), pageable and pinned host
memory, every host entry point compared with the device-pointer path of the same library.  usage: host_path_stress.py [seconds=60]"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cuzk_b200 import api, lib as cl
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(time.time()); rng = np.random.default_rng(seed)
api.initialize(0); L = cl.get_lib()
h = api.CudaPoseidonHash(); F = api.CudaFieldArithmetic
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
host = lambda t: t.cpu().numpy().view(np.uint64)
def pinned(a):
    t = torch.from_numpy(a.view(np.int64)).pin_memory()
    return t.numpy().view(np.uint64)          # numpy view of pinned memory: the library sees a registered host pointer
t0 = time.time(); rounds = 0; units = 0
while time.time() - t0 < budget:
    rounds += 1
    n = int(rng.choice([rng.integers(1, 5000), rng.integers(50_000, 400_000), rng.integers(400_000, 1_500_000)]))
    a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64); b = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    a[:, 3] >>= np.uint64(rng.integers(0, 5)); b[:, 3] >>= np.uint64(rng.integers(0, 5))
    if rng.integers(0, 2):
        a, b = pinned(a), pinned(b)
    da, db = dev(a), dev(b)
    assert (h.batch_hash_pairs(a, b) == host(h.batch_hash_pairs(da, db))).all(), (seed, rounds, n, "pairs")
    assert (h.batch_hash_single(a) == host(h.batch_hash_single(da))).all(), (seed, rounds, n, "single")
    assert (F.batch_multiply(a, b) == host(F.batch_multiply(da, db))).all(), (seed, rounds, n, "mul")
    assert (F.batch_subtract(a, b) == host(F.batch_subtract(da, db))).all(), (seed, rounds, n, "sub")
    m = n // 3
    if m:
        st = a[: 3 * m].copy()
        want = host(h.batch_permutation(dev(st))).reshape(-1, 3, 4)
        assert (h.batch_permutation(st) == want).all(), (seed, rounds, n, "permutation")
    w = int(rng.integers(1, 9)); k = n // w
    if k:
        assert (h.batch_sponge(a[: k * w], w, 3) == host(h.batch_sponge(da[: k * w], w, 3))).all(), (seed, rounds, n, "sponge")
    arity = int(rng.integers(2, 9)); nl = min(n, 300_000)
    t_h = api.CudaNaryMerkleTree(a[:nl], arity=arity); t_d = api.CudaNaryMerkleTree(da[:nl], arity=arity)
    assert (t_h.levels == host(t_d.levels)).all(), (seed, rounds, n, "merkle")
    units += n
print(json.dumps({"seed": seed, "seconds": round(time.time() - t0, 1), "rounds": rounds, "elements": units, "mismatches": 0}))
