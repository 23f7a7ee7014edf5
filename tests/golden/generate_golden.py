#!/usr/bin/env python
"""Generate tests/golden/*.json from the UNMODIFIED reference CPU implementation.

Runs only where /root/reference exists (the build container): oracle/Makefile compiles the reference's
own sources into oracle/_ref/libcuzk_ref.so, this script calls it through tests/oracle_lib.Ref and
writes the outputs as hex strings.  The fixtures are committed; the GPU box never needs the reference.

    python tests/golden/generate_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import K_INT, P_INT, Ref, hexes, ints_to_array  # noqa: E402

ref = Ref()
rng = np.random.default_rng(20261018)


def rnd(n, canonical):
    a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    if canonical:
        a[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    return a


EDGE = [0, 1, 2, 5, P_INT - 1, P_INT, P_INT + 1, 2 * P_INT, 4 * P_INT + 7, 5 * P_INT, 5 * P_INT + 1, 2**256 - 1, 2**128 - 1,
        2**255, K_INT, 2**256 - 2**64, 2**64 - 1, 2**64, (1 << 192) + (5 << 64), 1 + ((2**64 - 1) << 64), 2**224 - 1, 2**253]


def dump(name, obj):
    path = os.path.join(HERE, name)
    with open(path, "w") as f:
        json.dump(obj, f, indent=0, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


# ---- field ops ----
edge = ints_to_array(EDGE)
a = np.concatenate([edge, edge, rnd(48, True), rnd(48, False)])
b = np.concatenate([edge[::-1], np.roll(edge, 7, axis=0), rnd(48, True), rnd(48, False)])
fr = {"a": hexes(a), "b": hexes(b)}
for op in ("add", "sub", "mul", "sqr", "pow5"):
    fr[op] = hexes(ref.batch_fr(op, a, b))
prod = rng.integers(0, 2**64, size=(48, 8), dtype=np.uint64)
prod[:4] = np.uint64(2**64 - 1)
prod[4:12, 4:] = 0
prod[12:24, 5:] = 0
prod[12:24, 4] &= np.uint64(7)
prod[24:28, 6:] = 0
fr["reduce_512_in"] = ["%0128x" % sum(int(prod[i, j]) << (64 * j) for j in range(8)) for i in range(prod.shape[0])]
fr["reduce_512_out"] = hexes(ref.reduce_512(prod))
dump("fr_ops.json", fr)

# ---- poseidon ----
pos = {"round_constants": hexes(ref.round_constants()), "mds": hexes(ref.mds())}
states = np.concatenate([ints_to_array([1, 2, 3]), ints_to_array([0, 0, 0]), ints_to_array([P_INT - 1, 2**256 - 1, P_INT]),
                         rnd(3 * 8, True), rnd(3 * 5, False)]).reshape(-1, 3, 4)
pos["perm_in"] = hexes(states.reshape(-1, 4))
pos["perm_out"] = hexes(ref.permutation(states).reshape(-1, 4))
x = np.concatenate([ints_to_array([0, 1, 42, P_INT - 1, P_INT, 2**256 - 1]), rnd(20, True), rnd(6, False)])
y = np.concatenate([ints_to_array([0, 2, 20, 2**256 - 1, 1, P_INT + 5]), rnd(20, True), rnd(6, False)])
pos["single_in"] = hexes(x)
pos["single_out"] = hexes(ref.hash_single(x))
pos["pair_l"], pos["pair_r"] = hexes(x), hexes(y)
pos["pair_out"] = hexes(ref.hash_pairs(x, y))
pos["sponge"] = []
for width in range(0, 9):
    for ds in (3, 5):
        z = np.concatenate([rnd(4 * width, True), rnd(2 * width, False)]) if width else np.zeros((0, 4), dtype=np.uint64)
        pos["sponge"].append({"width": width, "ds": ds, "in": hexes(z), "out": hexes(ref.sponge(z, width, ds))})
z = ints_to_array([1, 2, 3, 4])
pos["hash_multiple_1234"] = hexes(ref.sponge(z, 4, 3))[0]
pos["empty_hash"] = {str(ar): hexes(ref.empty_hash(ar))[0] for ar in range(2, 9)}
dump("poseidon.json", pos)

# ---- merkle ----
mk = {"trees": []}
for arity in range(2, 9):
    for n in (1, 2, 3, 5, 8, 9, 16, 27, 64, 100):
        leaves = ref.generate_test_leaves(n, 42)
        t = ref.tree(leaves, arity)
        ent = {"arity": arity, "n": n, "seed": 42, "root": hexes(t.root())[0], "height": t.height(), "proofs": []}
        for idx in sorted({0, n // 2, n - 1}):
            sib, p = t.prove(idx)
            ent["proofs"].append({"index": idx, "positions": [int(v) for v in p], "siblings": hexes(sib.reshape(-1, 4))})
        mk["trees"].append(ent)
# full-width (not 64-bit) leaves too
for arity, n in ((2, 13), (3, 10), (4, 17), (8, 65)):
    leaves = rnd(n, False)
    t = ref.tree(leaves, arity)
    mk["trees"].append({"arity": arity, "n": n, "leaves": hexes(leaves), "root": hexes(t.root())[0], "height": t.height(), "proofs": []})
mk["leaves_seed42_first8"] = hexes(ref.generate_test_leaves(8, 42))
mk["empty_root"] = {str(ar): hexes(ref.tree(np.zeros((0, 4), dtype=np.uint64), ar).root())[0] for ar in range(2, 9)}
mk["height_table"] = [[n, ar, ref.tree_height_float(n, ar)] for ar in range(2, 9)
                      for n in (0, 1, 2, 3, 4, 7, 8, 9, 16, 64, 125, 216, 1000, 4096, 50000, 2**20, 2**21, 2**26, 8**7, 8**9)]
dump("merkle.json", mk)
