#!/usr/bin/env python
"""Randomised differential run against the oracle: fresh seeds every run, all entry points, for a wall-clock budget.
usage: fuzz.py [seconds=60] [seed]     prints one JSON line; exits non-zero on the first mismatch"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cuzk_b200 import api, lib as cl
from oracle_lib import Oracle
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else int(time.time())
rng = np.random.default_rng(seed)
api.initialize(0); L = cl.get_lib(); oracle = Oracle()
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
host = lambda t: t.cpu().numpy().view(np.uint64)
def elems(n):
    kind = rng.integers(0, 4)
    a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    if kind == 0: a[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)          # canonical
    elif kind == 1: a[:, 1:] = 0                                      # 64-bit values
    elif kind == 2: a[:, 3] |= np.uint64(0xF000000000000000)          # far above p
    return a                                                          # kind 3: any 256-bit value
h = api.CudaPoseidonHash(); F = api.CudaFieldArithmetic
counts = {"field": 0, "hash": 0, "sponge": 0, "tree_nodes": 0, "proofs": 0, "updates": 0}
t0 = time.time(); rounds = 0
while time.time() - t0 < budget:
    rounds += 1
    n = int(rng.integers(1, 3000))
    a, b = elems(n), elems(n)
    on_dev = bool(rng.integers(0, 2))
    w = (lambda x: dev(x)) if on_dev else (lambda x: x)
    g = (lambda x: host(x)) if on_dev else (lambda x: x)
    for name, fn in (("add", F.batch_add), ("sub", F.batch_subtract), ("mul", F.batch_multiply)):
        assert (g(fn(w(a), w(b))) == oracle.batch_fr(name, a, b)).all(), (seed, rounds, name)
    assert (g(F.batch_power5(w(a))) == oracle.batch_fr("pow5", a)).all(), (seed, rounds, "pow5")
    counts["field"] += 4 * n
    assert (g(h.batch_hash_pairs(w(a), w(b))) == oracle.hash_pairs(a, b)).all(), (seed, rounds, "pairs")
    assert (g(h.batch_hash_single(w(a))) == oracle.hash_single(a)).all(), (seed, rounds, "single")
    counts["hash"] += 2 * n
    width = int(rng.integers(1, 12)); m = max(1, n // width // 4)
    x = elems(m * width)
    assert (g(h.batch_sponge(w(x), width, 3)) == oracle.sponge(x, width, 3)).all(), (seed, rounds, "sponge")
    counts["sponge"] += m
    arity = int(rng.integers(2, 9)); nl = int(rng.integers(1, 1500))
    leaves = elems(nl)
    want = oracle.merkle_build(leaves, arity)
    t = api.DeviceMerkleTree(w(leaves), arity=arity)
    got = t.get_tree_levels()
    assert len(got) == len(want) and all((x1 == x2).all() for x1, x2 in zip(got, want)), (seed, rounds, "build", arity, nl)
    counts["tree_nodes"] += sum(x.shape[0] for x in want)
    if nl > 1:
        idx = np.unique(rng.integers(0, nl, size=min(nl, 64))).astype(np.uint64)
        pb = t.generate_batch_proofs(dev(idx.astype(np.int64)) if on_dev else idx)
        lv = leaves[idx.astype(np.int64)].copy()
        flip = rng.integers(0, 2, size=idx.size).astype(bool)
        lv[flip, int(rng.integers(0, 4))] ^= np.uint64(1 << int(rng.integers(0, 60)))
        res = t.verify_batch_proofs(pb, w(lv))
        res = res.cpu().numpy() if on_dev else res
        assert (res.astype(bool) == ~flip).all(), (seed, rounds, "verify")
        counts["proofs"] += idx.size
        upd = np.unique(rng.integers(0, nl, size=min(nl, 16))).astype(np.uint64)
        newv = elems(upd.size)
        t.update_leaves(dev(upd.astype(np.int64)) if on_dev else upd, w(newv))
        leaves[upd.astype(np.int64)] = newv
        assert (t.get_root_hash() == oracle.merkle_build(leaves, arity)[-1][0]).all(), (seed, rounds, "update")
        counts["updates"] += upd.size
    t.close()
print(json.dumps({"seed": seed, "seconds": round(time.time() - t0, 1), "rounds": rounds, "checked": counts, "mismatches": 0,
                  "exact_fallbacks_taken": int(L.cuzk_debug_fallback_count())}))
