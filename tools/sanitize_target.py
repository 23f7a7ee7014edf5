#!/usr/bin/env python
"""Small run of every kernel through the C ABI, each result checked against the oracle (seconds on a GPU).  Written as a
compute-sanitizer target (`compute-sanitizer --tool memcheck python tools/sanitize_target.py`); on the round-1 GPU pool
compute-sanitizer is closed, so it only ran plain there."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cuzk_b200 import api, lib as cl  # noqa: E402
from oracle_lib import Oracle, synth_elements, synth_u64_leaves  # noqa: E402

api.initialize(0)
L = cl.get_lib()
oracle = Oracle()
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()
host = lambda t: t.cpu().numpy().view(np.uint64)

n = 2048
l, r = synth_elements(1, n), synth_elements(2, n)
h = api.CudaPoseidonHash()
assert (host(h.batch_hash_pairs(dev(l), dev(r))) == oracle.hash_pairs(l, r)).all()
assert (h.batch_hash_pairs(l, r) == oracle.hash_pairs(l, r)).all()          # host-buffer pipeline
assert (host(h.batch_hash_single(dev(l[:1000]))) == oracle.hash_single(l[:1000])).all()
st = synth_elements(3, 3 * 300, canonical=False)
assert (host(h.batch_permutation(dev(st))).reshape(-1, 3, 4) == oracle.permutation(st.reshape(-1, 3, 4))).all()
for w in (1, 3, 8):
    x = synth_elements(4, 200 * w)
    assert (host(h.batch_sponge(dev(x), w, 3)) == oracle.sponge(x, w, 3)).all()
for op, name in ((api.CudaFieldArithmetic.batch_add, "add"), (api.CudaFieldArithmetic.batch_subtract, "sub"), (api.CudaFieldArithmetic.batch_multiply, "mul")):
    assert (host(op(dev(l), dev(r))) == oracle.batch_fr(name, l, r)).all()
assert (host(api.CudaFieldArithmetic.batch_power5(dev(l))) == oracle.batch_fr("pow5", l)).all()
for coop_max in (0, 1 << 20):
    L.cuzk_debug_set_coop_max(coop_max)
    for arity, m in ((2, 300), (4, 300), (8, 700)):
        leaves = synth_u64_leaves(5, m)
        want = oracle.merkle_build(leaves, arity)
        t = api.CudaNaryMerkleTree(dev(leaves), arity=arity)
        assert all((host(g) == w).all() for g, w in zip(t.get_tree_levels(), want))
        idx = torch.arange(0, m, 3, dtype=torch.int64, device="cuda")
        pb = t.generate_batch_proofs(idx)
        assert bool(t.verify_batch_proofs(pb, dev(leaves)[idx].contiguous()).all())
L.cuzk_debug_set_coop_max(2368)
trees = api.build_batch_trees(dev(synth_u64_leaves(6, 5 * 64)).reshape(5, 64, 4), arity=4)
assert (host(trees[3].get_tree_levels()[-1])[0] == oracle.merkle_build(synth_u64_leaves(6, 5 * 64)[192:256], 4)[-1][0]).all()
# grouped subtree-root pass (several real subtrees, height >= 3) + top levels
leaves = synth_u64_leaves(7, 50)
roots = torch.empty((8, 4), dtype=torch.int64, device="cuda")
L.check(L.cuzk_merkle_subtree_roots(dev(leaves).data_ptr(), 50, 2, 3, 8, roots.data_ptr(), 0, None), "subtree_roots")
root = torch.empty((1, 4), dtype=torch.int64, device="cuda")
L.check(L.cuzk_merkle_top_root(roots.data_ptr(), 8, 2, root.data_ptr(), 0, None), "top_root")
assert (host(root)[0] == oracle.merkle_build(leaves, 2)[-1][0]).all()
dt = api.DeviceMerkleTree(leaves, arity=2)
dt.update_leaves(np.array([0, 49], dtype=np.uint64), synth_elements(8, 2))
leaves2 = leaves.copy()
leaves2[[0, 49]] = synth_elements(8, 2)
assert (dt.get_root_hash() == oracle.merkle_build(leaves2, 2)[-1][0]).all()
dt.close()
torch.cuda.synchronize()
print("sanitize target ok")
