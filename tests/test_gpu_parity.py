"""GPU parity tests (run with -m gpu on a B200): every result produced through the C ABI is compared
bit-for-bit with the oracle on the same seeded inputs and with the committed golden fixtures."""
import numpy as np
import pytest

from conftest import load_golden
from helpers import h2a, rnd, to_dev, to_host
from oracle_lib import K_INT, P_INT, hexes, ints_to_array, synth_elements, synth_u64_leaves
from test_oracle import mt19937_64_leaves

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["default", "cooperative"])
def kernel_path(request):
    """Every test of this module runs twice: with the launchers' own choice between the one-thread-per-unit and the
    cooperative (sixteen lanes per permutation) kernels, and with every launch of up to 2^20 units forced onto the
    cooperative kernels.  Results must not depend on the path."""
    from cuzk_b200 import lib

    L = lib.get_lib()
    old = L.cuzk_debug_set_coop_max(1 << 20) if request.param == "cooperative" else None
    yield request.param
    if old is not None:
        L.cuzk_debug_set_coop_max(old)

EDGE = ints_to_array([0, 1, 2, P_INT - 1, P_INT, P_INT + 1, 2 * P_INT, 5 * P_INT, 5 * P_INT + 1, 2**256 - 1, 2**128 - 1, 2**255, K_INT,
                      2**256 - 2**64, 2**64 - 1, (1 << 192) + (5 << 64), 1 + ((2**64 - 1) << 64), 2**224 - 1])


def _ops(api):
    F = api.CudaFieldArithmetic
    return {"add": F.batch_add, "sub": F.batch_subtract, "mul": F.batch_multiply, "sqr": lambda a, b: F.batch_square(a),
            "pow5": lambda a, b: F.batch_power5(a)}


def test_native_library_is_loaded(gpu):
    from cuzk_b200 import lib

    before = lib.get_lib().cuzk_launch_count()
    gpu.CudaFieldArithmetic.batch_add(EDGE, EDGE)
    assert lib.get_lib().cuzk_launch_count() > before
    assert lib.LIB_PATH in open("/proc/self/maps").read()


def test_constants_match_oracle(gpu, oracle):
    rc, mds = gpu.CudaPoseidonHash.constants()
    assert (rc == oracle.round_constants()).all()
    assert (mds == oracle.mds()).all()


@pytest.mark.parametrize("where", ["host", "device"])
def test_field_ops_golden(gpu, where):
    g = load_golden("fr_ops.json")
    a, b = h2a(g["a"]), h2a(g["b"])
    for op, fn in _ops(gpu).items():
        got = fn(a, b) if where == "host" else to_host(fn(to_dev(a), to_dev(b)))
        assert hexes(got) == g[op], op


@pytest.mark.parametrize("canonical", [True, False])
def test_field_ops_random_vs_oracle(gpu, oracle, canonical):
    rng = np.random.default_rng(100 + canonical)
    n = 200_000
    a = np.concatenate([rnd(rng, n, canonical), EDGE, EDGE[::-1]])
    b = np.concatenate([rnd(rng, n, canonical), EDGE[::-1], EDGE])
    da, db = to_dev(a), to_dev(b)
    for op, fn in _ops(gpu).items():
        m = n if op in ("add", "sub") else (n if op != "pow5" else 60_000)
        sel = np.r_[0:m, n : a.shape[0]]
        got = to_host(fn(da, db))[sel]
        want = oracle.batch_fr(op, a[sel], b[sel])
        bad = np.nonzero((got != want).any(axis=1))[0]
        assert bad.size == 0, (op, canonical, bad[:5], hexes(a[sel][bad[:1]]), hexes(b[sel][bad[:1]]))


def test_field_small_operands(gpu, oracle):
    """mh == 0 / high == 0 paths of reduce_512 (small operands, as in the first rounds of hash_single(0))."""
    vals = [0, 1, 2, 3, 5, 2**32 - 1, 2**32, 2**64 - 1, 2**64, 2**96 + 12345, 2**127, 2**128 - 1, 2**129, 2**160 + 7, 2**192 - 1]
    a = ints_to_array([x for x in vals for _ in vals])
    b = ints_to_array([y for _ in vals for y in vals])
    for op, fn in _ops(gpu).items():
        assert (fn(a, b) == oracle.batch_fr(op, a, b)).all(), op


def test_poseidon_golden(gpu):
    g = load_golden("poseidon.json")
    h = gpu.CudaPoseidonHash()
    assert hexes(h.batch_hash_single(h2a(g["single_in"]))) == g["single_out"]
    assert hexes(h.batch_hash_pairs(h2a(g["pair_l"]), h2a(g["pair_r"]))) == g["pair_out"]
    st = h2a(g["perm_in"]).reshape(-1, 3, 4).copy()
    assert hexes(h.batch_permutation(st).reshape(-1, 4)) == g["perm_out"]
    for case in g["sponge"]:
        got = h.batch_sponge(h2a(case["in"]), case["width"], case["ds"])
        assert hexes(got) == case["out"], (case["width"], case["ds"])
    for a, want in g["empty_hash"].items():
        assert hexes(gpu.empty_hash(int(a)))[0] == want


def test_poseidon_random_vs_oracle(gpu, oracle):
    rng = np.random.default_rng(5)
    h = gpu.CudaPoseidonHash()
    n = 3000
    x = np.concatenate([rnd(rng, n, False), EDGE])
    y = np.concatenate([rnd(rng, n, True), EDGE[::-1]])
    assert (to_host(h.batch_hash_single(to_dev(x))) == oracle.hash_single(x)).all()
    assert (to_host(h.batch_hash_pairs(to_dev(x), to_dev(y))) == oracle.hash_pairs(x, y)).all()
    assert (h.batch_hash_pairs(x, y) == oracle.hash_pairs(x, y)).all()  # host-pointer path
    st = rnd(rng, 3 * 1000, False).reshape(-1, 3, 4)
    assert (to_host(h.batch_permutation(to_dev(st.reshape(-1, 4)))).reshape(-1, 3, 4) == oracle.permutation(st)).all()
    for width in (1, 2, 3, 4, 5, 7, 8):
        z = rnd(rng, 200 * width, False)
        assert (to_host(h.batch_sponge(to_dev(z), width, 3)) == oracle.sponge(z, width, 3)).all(), width


def test_reference_test_inputs(gpu, oracle):
    """The inputs verify_cuda_implementations_match uses (poseidon_cuda_benchmarks.cpp:137-259) and the
    sequential inputs of PoseidonCUDATest (test_poseidon_cuda.cpp:90-110)."""
    h = gpu.CudaPoseidonHash()
    i = np.arange(100, dtype=np.uint64)
    single = np.stack([i + 1, 2 * i + 1, 3 * i + 1, 4 * i + 1], axis=1).astype(np.uint64)
    right = np.stack([5 * i + 1, 6 * i + 1, 7 * i + 1, 8 * i + 1], axis=1).astype(np.uint64)
    assert (h.batch_hash_single(single) == oracle.hash_single(single)).all()
    assert (h.batch_hash_pairs(single, right) == oracle.hash_pairs(single, right)).all()
    seq = np.zeros((10000, 4), dtype=np.uint64)
    seq[:, 0] = np.arange(10000)
    got = h.batch_hash_single(seq)
    pick = np.r_[0:50, 5000:5050, 9950:10000]
    assert (got[pick] == oracle.hash_single(seq[pick])).all()


def test_empty_and_mismatched_batches(gpu):
    h = gpu.CudaPoseidonHash()
    z = np.zeros((0, 4), dtype=np.uint64)
    assert h.batch_hash_single(z).shape == (0, 4)
    assert h.batch_hash_pairs(z, z).shape == (0, 4)
    with pytest.raises(Exception):
        h.batch_hash_pairs(np.zeros((3, 4), dtype=np.uint64), np.zeros((2, 4), dtype=np.uint64))
    with pytest.raises(Exception):
        gpu.CudaFieldArithmetic.batch_add(np.zeros((3, 4), dtype=np.uint64), np.zeros((2, 4), dtype=np.uint64))
    with pytest.raises(ValueError):
        gpu.CudaNaryMerkleTree(arity=1)
    with pytest.raises(ValueError):
        gpu.CudaNaryMerkleTree(arity=10)


def test_pair_hash_synthetic_stream_and_determinism(gpu, oracle):
    """Config 1 inputs (seeded splitmix64 stream): device generator == host generator, GPU == oracle on a sample,
    and two runs agree bit-for-bit at the full 1M size."""
    import torch

    from cuzk_b200 import lib

    L = lib.get_lib()
    n = 1_000_000
    dl = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    dr = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    L.check(L.cuzk_synth_elements(dl.data_ptr(), n, 1, 0, 1, None), "synth")
    L.check(L.cuzk_synth_elements(dr.data_ptr(), n, 2, 0, 1, None), "synth")
    assert (to_host(dl[:1000]) == synth_elements(1, 1000)).all()
    assert (to_host(dr[-1000:]) == synth_elements(2, 1000, start=n - 1000)).all()
    h = gpu.CudaPoseidonHash()
    o1 = h.batch_hash_pairs(dl, dr)
    o2 = h.batch_hash_pairs(dl, dr)
    assert torch.equal(o1, o2)
    pick = torch.cat([torch.arange(0, 2000), torch.arange(n - 2000, n), torch.randint(0, n, (2000,))]).cuda()
    want = oracle.hash_pairs(to_host(dl[pick]), to_host(dr[pick]))
    assert (to_host(o1[pick]) == want).all()
    # canonical outputs
    top = to_host(o1)[:, 3]
    assert (top <= np.uint64(0x30644E72E131A029)).all()


# ---------------------------------------------------------------- Merkle ----
def test_merkle_golden(gpu):
    g = load_golden("merkle.json")
    for ent in g["trees"]:
        leaves = h2a(ent["leaves"]) if "leaves" in ent else mt19937_64_leaves(ent["n"], 42)
        t = gpu.CudaNaryMerkleTree(leaves, arity=ent["arity"])
        assert hexes(t.get_root_hash())[0] == ent["root"], (ent["arity"], ent["n"])
        assert t.get_tree_height() == ent["height"]
        if ent["proofs"] and ent["n"] > 1:
            idx = [p["index"] for p in ent["proofs"]]
            pb = t.generate_batch_proofs(idx)
            for k, p in enumerate(ent["proofs"]):
                assert [int(v) for v in pb.positions[k]] == p["positions"]
                assert hexes(pb.siblings[k].reshape(-1, 4)) == p["siblings"]
    for a, want in g["empty_root"].items():
        assert hexes(gpu.CudaNaryMerkleTree(arity=int(a)).get_root_hash())[0] == want


@pytest.mark.parametrize("arity", [2, 3, 4, 5, 6, 7, 8])
def test_merkle_levels_proofs_verify_vs_oracle(gpu, oracle, arity):
    rng = np.random.default_rng(arity)
    for n in (1, 2, arity, arity + 1, arity**2, arity**2 + 1, 100, 257, 1000):
        leaves = rnd(rng, n, False) if n % 2 else synth_u64_leaves(9, n)
        want = oracle.merkle_build(leaves, arity)
        for dev in (False, True):
            t = gpu.CudaNaryMerkleTree(to_dev(leaves) if dev else leaves, arity=arity)
            levels = [to_host(x) if dev else x for x in t.get_tree_levels()]
            assert len(levels) == len(want)
            for l, (gl, wl) in enumerate(zip(levels, want)):
                assert (gl == wl).all(), (arity, n, dev, l)
            if n == 1:
                continue
            idx = np.unique(np.r_[0, n - 1, rng.integers(0, n, size=min(n, 40))])
            pb = t.generate_batch_proofs(to_dev(idx.astype(np.int64)) if dev else idx)
            sib = to_host(pb.siblings) if dev else pb.siblings
            pos = pb.positions.cpu().numpy().view(np.uint32) if dev else pb.positions
            for k, i in enumerate(idx):
                so, po = oracle.merkle_prove(want, n, arity, int(i))
                assert (sib[k] == so).all() and (pos[k] == po.astype(np.uint32)).all()
            lv = leaves[idx]
            res = t.verify_batch_proofs(pb, to_dev(lv) if dev else lv)
            res = res.cpu().numpy() if dev else res
            assert res.all()
            # corrupted leaf / sibling / position / root must fail exactly like the oracle
            bad_leaf = lv.copy()
            bad_leaf[0, 0] ^= np.uint64(1)
            r2 = t.verify_batch_proofs(pb, to_dev(bad_leaf) if dev else bad_leaf)
            r2 = r2.cpu().numpy() if dev else r2
            assert r2[0] == 0 and r2[1:].all()
            if not dev:
                sib2 = pb.siblings.copy()
                sib2[-1, -1, 0, 3] ^= np.uint64(1 << 40)
                pos2 = pb.positions.copy()
                pos2[0, 0] = (pos2[0, 0] + 1) % arity
                pb2 = gpu.MerkleProofBatch(sib2, pos2, idx, arity)
                r3 = t.verify_batch_proofs(pb2, lv)
                want3 = [oracle.merkle_verify(lv[k], sib2[k], pos2[k].astype(np.uint64), arity, want[-1][0]) for k in range(len(idx))]
                assert [bool(v) for v in r3] == want3
                pos3 = pb.positions.copy()
                pos3[0, -1] = arity  # out-of-range position -> false (merkle_tree.cpp:228-230)
                r4 = t.verify_batch_proofs(gpu.MerkleProofBatch(pb.siblings, pos3, idx, arity), lv)
                assert r4[0] == 0 and r4[1:].all()
            # out-of-range index -> sentinel positions (reference: std::nullopt)
            if not dev:
                pbo = t.generate_batch_proofs(np.array([n, 0], dtype=np.uint64))
                assert (pbo.positions[0] == 0xFFFFFFFF).all() and (pbo.positions[1] != 0xFFFFFFFF).all()


def test_config2_binary_50k_leaves_5k_proofs(gpu, oracle):
    """BASELINE config: binary tree over 50 000 leaves (generate_test_leaves seed 0) + 5 000 proof verifications."""
    n, arity = 50_000, 2
    leaves = mt19937_64_leaves(n, 0)
    t = gpu.CudaNaryMerkleTree(to_dev(leaves), arity=arity)
    levels = [to_host(x) for x in t.get_tree_levels()]
    assert len(levels) == 17 and levels[0].shape[0] == 65536
    # level 1 in full and every upper level against the oracle's hashing of OUR lower level (chain of custody)
    for l in range(1, len(levels)):
        m = levels[l].shape[0]
        sel = np.arange(m) if m <= 4096 else np.unique(np.r_[0:1024, m - 1024 : m, np.random.default_rng(l).integers(0, m, 2048)])
        kids = levels[l - 1].reshape(m, arity, 4)[sel].reshape(-1, 4)
        assert (levels[l][sel] == oracle.sponge(kids, arity, 3)).all(), l
    idx = (np.arange(5000, dtype=np.uint64) * 10) % n
    pb = t.generate_batch_proofs(to_dev(idx.astype(np.int64)))
    res = t.verify_batch_proofs(pb, to_dev(leaves[idx]))
    assert bool(res.all())
    k = 17
    so, po = oracle.merkle_prove(levels, n, arity, int(idx[k]))
    assert (to_host(pb.siblings[k]) == so).all()
    assert oracle.merkle_verify(leaves[idx[k]], so, po, arity, levels[-1][0])


def test_subtree_roots_and_top_equal_full_build(gpu, oracle):
    """The multi-GPU decomposition on one GPU: per-shard subtree roots + top levels == full build root."""
    import torch

    from cuzk_b200 import lib

    L = lib.get_lib()
    for arity, n, height in ((8, 8**3, 2), (8, 300, 2), (4, 4**4, 2), (4, 200, 3), (2, 1000, 5), (3, 81, 2), (5, 30, 1)):
        leaves = synth_u64_leaves(4, n)
        want_root = oracle.merkle_build(leaves, arity)[-1][0]
        padded = gpu.padded_leaves(n, arity)
        span = arity**height
        count = padded // span
        d = to_dev(leaves)
        roots = torch.empty((count, 4), dtype=torch.int64, device="cuda")
        L.check(L.cuzk_merkle_subtree_roots(d.data_ptr(), n, arity, height, count, roots.data_ptr(), 0, None), "subtree_roots")
        root = torch.empty((1, 4), dtype=torch.int64, device="cuda")
        L.check(L.cuzk_merkle_top_root(roots.data_ptr(), count, arity, root.data_ptr(), 0, None), "top_root")
        assert (to_host(root)[0] == want_root).all(), (arity, n, height)
        # all-padding subtrees carry the padding constant of that level
        real = -(-n // span)
        if real < count:
            assert (to_host(roots[real:]) == gpu.padding_root(arity, height)).all()


def test_forest_build_equals_per_tree_builds(gpu, oracle):
    """cuzk_merkle_build_batch: many equal-sized trees in one forest pass == the oracle's tree, tree by tree."""
    for arity, n, T in ((2, 32, 10), (4, 100, 7), (8, 2048, 5), (3, 1, 4), (8, 4096, 40)):
        leaves = synth_u64_leaves(11, n * T).reshape(T, n, 4)
        for dev in (False, True):
            trees = gpu.build_batch_trees(to_dev(leaves.reshape(-1, 4)).reshape(T, n, 4) if dev else leaves, arity=arity)
            assert len(trees) == T
            for t in (0, T // 2, T - 1):
                want = oracle.merkle_build(leaves[t], arity)
                got = [to_host(x) if dev else x for x in trees[t].get_tree_levels()]
                assert len(got) == len(want)
                for gl, wl in zip(got, want):
                    assert (gl == wl).all(), (arity, n, T, t, dev)


def test_mds_layer_fast_path_and_fallback(gpu, oracle):
    """The fast MDS layer (linear form + wrap count + exact fallback) against the oracle's term-by-term layer, on
    states crafted to sit on every decision boundary, and against the GPU's own exact path."""
    import torch

    from cuzk_b200 import lib
    from oracle_lib import oracle_mds_layer
    from test_oracle import craft_mds_states

    L = lib.get_lib()
    rng = np.random.default_rng(99)
    st = np.concatenate([craft_mds_states(rng), rnd(rng, 3 * 200_000, True).reshape(-1, 3, 4)])
    want = oracle_mds_layer(oracle, st)
    for mode in (0, 1):
        d = to_dev(st.reshape(-1, 4))
        L.check(L.cuzk_debug_mds_layer(d.data_ptr(), st.shape[0], mode, None), "debug_mds")
        got = to_host(d).reshape(-1, 3, 4)
        bad = np.nonzero((got != want).any(axis=(1, 2)))[0]
        assert bad.size == 0, (mode, bad[:4], hexes(st[bad[0]]) if bad.size else None)


# ------------------------------------------------------- BASELINE full sizes, size-independent properties ----
def _level_views(flat, n, arity):
    out, off, p = [], 0, 1
    while p < n:
        p *= arity
    while True:
        out.append(flat[off : off + p])
        off += p
        if p == 1:
            return out
        p //= arity


def _check_sampled_nodes(levels, arity, oracle, rng, per_level=48):
    """chain of custody: sampled nodes of every level == oracle hash_multiple of OUR children one level below"""
    for l in range(1, len(levels)):
        m = levels[l].shape[0]
        sel = np.unique(np.r_[0, m - 1, rng.integers(0, m, size=min(m, per_level))])
        kids = to_host(levels[l - 1].reshape(m, arity, 4)[sel.tolist()].reshape(-1, 4))
        got = to_host(levels[l][sel.tolist()])
        assert (got == oracle.sponge(kids, arity, 3)).all(), l


def test_config3_quaternary_2p20_full_batch_verify(gpu, oracle):
    """BASELINE configs[2]: 4-ary tree over 2^20 leaves with full-batch proof verification (device resident)."""
    import torch

    from cuzk_b200 import lib

    L = lib.get_lib()
    n, arity = 1 << 20, 4
    leaves = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    L.check(L.cuzk_synth_u64_leaves(leaves.data_ptr(), n, 3, 0, None), "synth")
    assert (to_host(leaves[:64]) == synth_u64_leaves(3, 64)).all()
    t = gpu.CudaNaryMerkleTree(leaves, arity=arity)
    levels = t.get_tree_levels()
    assert len(levels) == 11 and levels[0].shape[0] == n
    rng = np.random.default_rng(3)
    _check_sampled_nodes(levels, arity, oracle, rng)
    idx = torch.arange(n, dtype=torch.int64, device="cuda")
    pb = t.generate_batch_proofs(idx)
    assert tuple(pb.siblings.shape) == (n, 10, 3, 4)
    res = t.verify_batch_proofs(pb, leaves)
    assert bool(res.all()), "a valid proof of the full batch was rejected"
    # every proof path of a sample equals the oracle's walk over OUR level arrays
    host_levels = None
    for q in rng.integers(0, n, size=8).tolist():
        at = q
        for l in range(10):
            grp = to_host(levels[l][at - at % arity : at - at % arity + arity])
            assert int(pb.positions[q, l]) == at % arity
            assert (to_host(pb.siblings[q, l]) == np.delete(grp, at % arity, axis=0)).all()
            at //= arity
    # corrupt a known subset: exactly those proofs fail
    bad = np.unique(rng.integers(0, n, size=1000))
    lv = leaves.clone()
    lv[bad.tolist(), 0] ^= 1
    res2 = t.verify_batch_proofs(pb, lv).cpu().numpy()
    want = np.ones(n, dtype=np.uint8)
    want[bad] = 0
    assert (res2 == want).all()
    del pb, lv


def test_config4_octary_2p26_sharded_equals_full_build(gpu, oracle):
    """BASELINE configs[3]: 8-ary tree over 2^26 leaves.  Full build (every level kept) and the sharded subtree-root
    decomposition for world sizes 1/2/4/8 (run shard by shard on this GPU) must give one and the same root, and sampled
    nodes must equal the oracle's hash of their children."""
    import torch

    from cuzk_b200 import lib
    from cuzk_b200.distributed import plan_merkle_shards

    L = lib.get_lib()
    n, arity = 1 << 26, 8
    leaves = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    L.check(L.cuzk_synth_u64_leaves(leaves.data_ptr(), n, 4, 0, None), "synth")
    t = gpu.CudaNaryMerkleTree(leaves, arity=arity)
    levels = t.get_tree_levels()
    assert len(levels) == 10 and levels[0].shape[0] == 1 << 27     # padded to 8^9
    assert t.get_tree_height() == 10                               # float formula agrees here (2^26 is not a power of 8)
    _check_sampled_nodes(levels, arity, oracle, np.random.default_rng(4), per_level=24)
    # padding: the upper half of every level is the level's padding constant
    for l in (0, 3, 8):
        assert (to_host(levels[l][-1:]) == gpu.padding_root(arity, l)).all()
    root = to_host(levels[-1])[0]
    del t, levels
    torch.cuda.empty_cache()
    for world in (1, 2, 4, 8):
        plan = plan_merkle_shards(n, arity, world)
        nodes = torch.empty((plan.total_subtrees, 4), dtype=torch.int64, device="cuda")
        for rank in range(world):
            lo, hi = plan.rank_subtrees(rank)
            l0, l1 = plan.rank_leaves(rank)
            L.check(L.cuzk_merkle_subtree_roots(leaves[l0:l1].data_ptr(), l1 - l0, arity, plan.height, hi - lo,
                                                nodes[lo:hi].data_ptr(), 0, None), "subtree_roots")
        if plan.total_subtrees > plan.real_subtrees:
            nodes[plan.real_subtrees :] = to_dev(gpu.padding_root(arity, plan.height).reshape(1, 4))
        out = torch.empty((1, 4), dtype=torch.int64, device="cuda")
        L.check(L.cuzk_merkle_top_root(nodes.data_ptr(), plan.total_subtrees, arity, out.data_ptr(), 0, None), "top_root")
        assert (to_host(out)[0] == root).all(), world


def test_exact_fallback_path_is_bit_exact(gpu, oracle, kernel_path):
    """The production kernels evaluate every unit on a fast path whose top-word comparisons can be undecided (2^-32 each);
    such units are recomputed on the exact path.  libcuzk_b200_widen.so is the same source built with the flags widened
    to near misses, so a large share of the units takes the fallback: every result must still equal the oracle."""
    import os

    import torch

    from cuzk_b200 import lib

    path = lib.DEBUG_LIB_PATH
    assert os.path.exists(path), "libcuzk_b200_widen.so missing: run __graft_entry__.build()"
    D = lib.Lib(path)
    D.check(D.cuzk_init(0), "init(widen)")
    if kernel_path == "cooperative":   # the widened flags of the cooperative kernels: every launch below goes through them
        D.cuzk_debug_set_coop_max(1 << 20)
    try:
        before = D.cuzk_debug_fallback_count()
        n = 20_000
        l, r = synth_elements(21, n), synth_elements(22, n)
        dl, dr = to_dev(l), to_dev(r)
        out = torch.empty_like(dl)
        D.check(D.cuzk_poseidon_hash_pairs(dl.data_ptr(), dr.data_ptr(), out.data_ptr(), n, 0, None), "pairs")
        assert (to_host(out) == oracle.hash_pairs(l, r)).all()
        D.check(D.cuzk_poseidon_hash_single(dl.data_ptr(), out.data_ptr(), n, 0, None), "single")
        assert (to_host(out) == oracle.hash_single(l)).all()
        rng = np.random.default_rng(5)
        st = rnd(rng, 3 * 3000, False)
        ds = to_dev(st)
        D.check(D.cuzk_poseidon_permutation(ds.data_ptr(), 3000, 0, None), "permutation")
        assert (to_host(ds).reshape(-1, 3, 4) == oracle.permutation(st.reshape(-1, 3, 4))).all()
        for arity, m in ((2, 3000), (8, 5000)):
            leaves = synth_u64_leaves(23, m) if arity == 8 else rnd(rng, m, True)
            want = oracle.merkle_build(leaves, arity)
            tot = sum(x.shape[0] for x in want)
            lv = torch.empty((tot, 4), dtype=torch.int64, device="cuda")
            D.check(D.cuzk_merkle_build(to_dev(leaves).data_ptr(), m, arity, lv.data_ptr(), 0, None), "build")
            assert (to_host(lv) == np.concatenate(want)).all()
            L = len(want) - 1
            idx = torch.arange(0, m, 7, dtype=torch.int64, device="cuda")
            q = idx.numel()
            sib = torch.empty((q, L, arity - 1, 4), dtype=torch.int64, device="cuda")
            pos = torch.empty((q, L), dtype=torch.int32, device="cuda")
            D.check(D.cuzk_merkle_prove_batch(lv.data_ptr(), m, arity, idx.data_ptr(), q, sib.data_ptr(), pos.data_ptr(), 0, None), "prove")
            res = torch.empty(q, dtype=torch.uint8, device="cuda")
            vals = to_dev(leaves)[idx].contiguous()
            D.check(D.cuzk_merkle_verify_batch(vals.data_ptr(), sib.data_ptr(), pos.data_ptr(), L, arity, lv[-1:].data_ptr(), res.data_ptr(), q, 0, None), "verify")
            assert bool(res.all())
        torch.cuda.synchronize()
        taken = D.cuzk_debug_fallback_count() - before
        assert taken > 1000, f"the widened build should take the exact path often, took it {taken} times"
    finally:
        D.cuzk_shutdown()


@pytest.mark.parametrize("arity", [2, 4, 8])
def test_device_tree_handle_and_incremental_update(gpu, oracle, arity):
    """cuzk_tree_*: device-resident levels, proofs and verification against them, batched leaf updates by path re-hash
    == a fresh oracle build over the updated leaves (the reference's update_leaf rebuilds the whole tree)."""
    rng = np.random.default_rng(100 + arity)
    n = 1000
    leaves = rnd(rng, n, True)
    for dev in (False, True):
        t = gpu.DeviceMerkleTree(to_dev(leaves) if dev else leaves, arity=arity)
        want = oracle.merkle_build(leaves, arity)
        assert (t.get_root_hash() == want[-1][0]).all()
        got = t.get_tree_levels()
        assert len(got) == len(want) and all((g == w).all() for g, w in zip(got, want))
        idx = np.unique(rng.integers(0, n, size=64)).astype(np.uint64)
        pb = t.generate_batch_proofs(to_dev(idx.astype(np.int64)) if dev else idx)
        sib = to_host(pb.siblings) if dev else pb.siblings
        for k, i in enumerate(idx):
            so, _ = oracle.merkle_prove(want, n, arity, int(i))
            assert (sib[k] == so).all()
        res = t.verify_batch_proofs(pb, to_dev(leaves[idx.astype(np.int64)]) if dev else leaves[idx.astype(np.int64)])
        assert bool((res.cpu().numpy() if dev else res).all())
        # batched update: 37 distinct leaves (first, last, a straddling group) get new values
        upd = np.unique(np.r_[0, n - 1, n - 2, rng.integers(0, n, size=34)]).astype(np.uint64)
        newv = rnd(rng, upd.size, False)          # non-canonical replacement values are absorbed like any leaf
        t.update_leaves(to_dev(upd.astype(np.int64)) if dev else upd, to_dev(newv) if dev else newv)
        leaves2 = leaves.copy()
        leaves2[upd.astype(np.int64)] = newv
        want2 = oracle.merkle_build(leaves2, arity)
        got2 = t.get_tree_levels()
        assert all((g == w).all() for g, w in zip(got2, want2)), (arity, dev)
        # the old proofs of untouched leaves whose path was not affected may now fail; fresh proofs must verify
        pb2 = t.generate_batch_proofs(to_dev(idx.astype(np.int64)) if dev else idx)
        lv2 = leaves2[idx.astype(np.int64)]
        res2 = t.verify_batch_proofs(pb2, to_dev(lv2) if dev else lv2)
        assert bool((res2.cpu().numpy() if dev else res2).all())
        if not dev:
            with pytest.raises(Exception):
                t.update_leaves(np.array([n], dtype=np.uint64), newv[:1])
        t.close()


@pytest.mark.parametrize("coop_max", [0, 64, 1 << 30])
def test_builds_agree_with_oracle_whatever_the_kernel_mix(gpu, oracle, coop_max):
    """Levels and subtree roots equal the oracle's for every arity, including ragged and single-group trees, whether the
    levels run on the one-thread kernels only (0), switch to the cooperative kernels for levels of <= 64 nodes, or run on
    the cooperative kernels throughout."""
    import torch

    from cuzk_b200 import lib

    L = lib.get_lib()
    old = L.cuzk_debug_set_coop_max(coop_max)
    try:
        rng = np.random.default_rng(coop_max % 7)
        for arity in range(2, 9):
            for n in (arity, arity * arity, arity**3 - 1, arity**3 + 1, 700):
                leaves = rnd(rng, n, n % 2 == 0)
                want = oracle.merkle_build(leaves, arity)
                t = gpu.CudaNaryMerkleTree(to_dev(leaves), arity=arity)
                got = [to_host(x) for x in t.get_tree_levels()]
                assert len(got) == len(want) and all((g == w).all() for g, w in zip(got, want)), (coop_max, arity, n)
                # roots only (no middle level stored), whole tree as one subtree
                root = torch.empty((1, 4), dtype=torch.int64, device="cuda")
                L.check(L.cuzk_merkle_subtree_roots(to_dev(leaves).data_ptr(), n, arity, len(want) - 1, 1, root.data_ptr(), 0, None), "roots")
                assert (to_host(root)[0] == want[-1][0]).all(), (coop_max, arity, n)
    finally:
        L.cuzk_debug_set_coop_max(old)


def test_large_tree_built_in_groups_equals_oracle(gpu, oracle):
    """Trees of >= 2^17 leaves hash their lower levels as groups of subtrees on internal streams (the narrow upper levels of
    one group hide behind the wide levels of the next): every stored level must still equal the oracle's, full and ragged,
    whatever the plan (cuzk_debug_set_build_plan: groups, streams, cooperative cap inside groups)."""
    from cuzk_b200.lib import get_lib

    L = get_lib()
    nocap = (1 << 64) - 1
    try:
        for arity, n in ((2, (1 << 17) + 12345), (4, 1 << 18), (8, 300_000)):
            leaves = synth_u64_leaves(40 + arity, n)
            want = oracle.merkle_build(leaves, arity)
            # the default plan, level by level on one stream, few / many groups, with and without cooperative kernels inside groups
            for plan in (None, (1, 1, nocap), (4, 4, nocap), (16, 16, 0), (5, 3, 256)):
                if plan is not None:
                    L.cuzk_debug_set_build_plan(*plan)
                t = gpu.DeviceMerkleTree(to_dev(leaves), arity=arity)
                got = t.get_tree_levels()
                assert len(got) == len(want) and all((g == w).all() for g, w in zip(got, want)), (arity, n, plan)
                t.close()
    finally:
        L.cuzk_debug_set_build_plan(8, 8, 1184)   # the defaults (host_path.cuh)


def test_update_semantics_duplicates_and_out_of_range(gpu, oracle):
    """cuzk_tree_update_leaves equals a serial loop of update_leaf calls (merkle_tree.cpp:294-301): the LAST value of a repeated
    index wins; an index >= leaf_count is refused on the host path (the reference throws) and skipped + counted on the
    asynchronous device path, without touching memory outside the leaf level."""
    import ctypes as C

    import torch

    from cuzk_b200 import lib

    L = lib.get_lib()
    rng = np.random.default_rng(77)
    n, arity = 1000, 4
    leaves = rnd(rng, n, True)
    t = gpu.DeviceMerkleTree(to_dev(leaves), arity=arity)
    idx = np.array([5, 17, 5, 999, 17, 5, 400], dtype=np.uint64)
    vals = rnd(rng, idx.size, True)
    expect = leaves.copy()
    for i, v in zip(idx, vals):
        expect[int(i)] = v
    for where in ("device", "host"):
        t2 = gpu.DeviceMerkleTree(to_dev(leaves), arity=arity)
        if where == "device":
            t2.update_leaves(torch.from_numpy(idx.view(np.int64)).cuda(), to_dev(vals))
        else:
            t2.update_leaves(idx, vals)
        got, want = t2.get_tree_levels(), oracle.merkle_build(expect, arity)
        assert all((g == w).all() for g, w in zip(got, want)), where
        t2.close()
    # out of range on the device path: skipped, counted, everything else applied
    bad_idx = np.array([3, n, 2**40, 7, n + 5], dtype=np.uint64)
    bad_vals = rnd(rng, bad_idx.size, True)
    t.update_leaves(torch.from_numpy(bad_idx.view(np.int64)).cuda(), to_dev(bad_vals))
    cnt = C.c_uint64()
    L.check(L.cuzk_tree_oob_count(t._h, C.byref(cnt)), "oob")
    assert cnt.value == 3
    expect2 = leaves.copy()
    expect2[3], expect2[7] = bad_vals[0], bad_vals[3]
    assert all((g == w).all() for g, w in zip(t.get_tree_levels(), oracle.merkle_build(expect2, arity)))
    with pytest.raises(Exception):
        t.update_leaves(bad_idx, bad_vals)   # host path: refused as a whole
    assert all((g == w).all() for g, w in zip(t.get_tree_levels(), oracle.merkle_build(expect2, arity)))
    t.close()


@pytest.mark.parametrize("arity,n", [(2, 5000), (8, 70_000), (4, 1), (3, 2)])
def test_sharded_tree_handle_on_one_device(gpu, oracle, arity, n):
    """cuzk_mg_* with a single shard (no NCCL involved): root, every proof path and the verdicts equal the oracle's for the
    whole tree; out-of-range queries get the sentinel."""
    leaves = synth_u64_leaves(60 + arity, n)
    want = oracle.merkle_build(leaves, arity)
    mg = gpu.MultiGpu.local(1)
    try:
        t = mg.build_tree(leaves, arity=arity)
        assert (t.get_root_hash() == want[-1][0]).all()
        assert t.num_levels == len(want)
        rng = np.random.default_rng(n)
        idx = np.unique(np.concatenate([rng.integers(0, n, 300), [0, n - 1]])).astype(np.uint64)
        pb = t.generate_batch_proofs(np.concatenate([idx, [n, 2**50]]).astype(np.uint64))
        L = len(want) - 1
        if L:
            assert (pb.positions[-2:] == 0xFFFFFFFF).all()
            for q, i in enumerate(idx):
                sib, pos = oracle.merkle_prove(want, n, arity, int(i))
                assert (pb.siblings[q] == sib).all() and (pb.positions[q] == pos).all(), (arity, n, int(i))
        good = gpu.MerkleProofBatch(pb.siblings[: idx.size], pb.positions[: idx.size], idx, arity)
        vals = leaves[idx.astype(np.int64)].copy()
        assert t.verify_batch_proofs(good, vals).all()
        if idx.size > 1:
            vals[1, 0] ^= 1
            res = t.verify_batch_proofs(good, vals)
            assert res[0] == 1 and res[1] == 0
        l, r = synth_elements(1, 3000), synth_elements(2, 3000)
        assert (mg.batch_hash_pairs(l, r) == oracle.hash_pairs(l, r)).all()
        t.close()
    finally:
        mg.close()


def test_sharded_tree_over_all_devices_nccl(gpu, oracle):
    """One process driving every GPU of the box through cuzk_mg_init_local: the subtree roots cross NVLink in one NCCL
    all-gather issued by the library; root == single-device full build == oracle, proofs served from the shards verify."""
    import torch

    ngpus = min(torch.cuda.device_count(), 8)
    if ngpus < 2:
        pytest.skip("needs at least two GPUs")
    for arity, n in ((8, 300_000), (2, (1 << 16) + 3), (5, 40)):
        leaves = synth_u64_leaves(80 + arity, n)
        want = oracle.merkle_build(leaves, arity)
        mg = gpu.MultiGpu.local(ngpus)
        try:
            t = mg.build_tree(leaves, arity=arity)
            assert (t.get_root_hash() == want[-1][0]).all(), (arity, n)
            single = gpu.DeviceMerkleTree(to_dev(leaves), arity=arity)
            assert (single.get_root_hash() == t.get_root_hash()).all()
            single.close()
            rng = np.random.default_rng(arity)
            idx = np.unique(np.concatenate([rng.integers(0, n, 12_000 if n > 1000 else 30), [0, n - 1]])).astype(np.uint64)
            pb = t.generate_batch_proofs(idx)
            for q in range(0, idx.size, max(1, idx.size // 50)):
                sib, pos = oracle.merkle_prove(want, n, arity, int(idx[q]))
                assert (pb.siblings[q] == sib).all() and (pb.positions[q] == pos).all(), (arity, n, int(idx[q]))
            vals = leaves[idx.astype(np.int64)].copy()
            assert t.verify_batch_proofs(pb, vals).all()
            vals[0, 0] ^= 1
            assert t.verify_batch_proofs(pb, vals)[0] == 0
            l, r = synth_elements(5, 10_000), synth_elements(6, 10_000)
            assert (mg.batch_hash_pairs(l, r) == oracle.hash_pairs(l, r)).all()
            t.close()
        finally:
            mg.close()


def test_fast_path_flags_are_sound(gpu, oracle):
    """The fast-path reduce / multiply / square / power5 may leave a comparison undecided, and must then say so: whenever the
    flag is clear the result equals the reference operation.  Checked on values crafted around every multiple of p (top words
    equal, one above, one below; low words all-ones / zero / random) and on random operands; the crafted reduce inputs must
    also actually raise the flag in the undecidable cases (otherwise this test would be vacuous)."""
    import torch

    from cuzk_b200 import lib
    from oracle_lib import P_INT, ints_to_array

    L = lib.get_lib()
    rng = np.random.default_rng(77)
    W = 1 << 256
    vals = []
    for m in range(0, 6):
        mp = m * P_INT
        top = mp >> 224
        for dt in (-1, 0, 1):
            t = top + dt
            if t < 0 or t >= (1 << 32):
                continue
            for low in (0, (1 << 224) - 1, mp & ((1 << 224) - 1), (mp & ((1 << 224) - 1)) - 1, (mp & ((1 << 224) - 1)) + 1,
                        int(rng.integers(0, 2**62)) << 160):
                vals.append(((t << 224) | (low & ((1 << 224) - 1))) % W)
        for d in (-2, -1, 0, 1, 2):
            if 0 <= mp + d < W:
                vals.append(mp + d)
    vals += [0, 1, W - 1, W - 2]
    crafted = ints_to_array(vals)

    def run(op, a, b=None):
        da, db = to_dev(a), (to_dev(b) if b is not None else None)
        out = torch.empty_like(da)
        fl = torch.empty(a.shape[0], dtype=torch.int32, device="cuda")
        L.check(L.cuzk_debug_fast_ops(op, da.data_ptr(), db.data_ptr() if db is not None else None, out.data_ptr(), fl.data_ptr(),
                                      a.shape[0], None), "debug_fast_ops")
        return to_host(out), fl.cpu().numpy()

    # reduce: oracle add(x, 0) = reduce(x)
    x = np.concatenate([crafted, rnd(rng, 200_000, False)])
    got, fl = run(0, x)
    want = oracle.batch_fr("add", x, np.zeros_like(x))
    ok = (got == want).all(axis=1)
    assert (ok | (fl != 0)).all(), "a fast reduce was wrong without raising its flag"
    assert fl[: crafted.shape[0]].any() and (~ok[: crafted.shape[0]]).any(), "crafted inputs did not reach the undecidable case"
    assert fl[crafted.shape[0] :].mean() < 1e-3
    # multiply / square / power5 on random canonical, random 256-bit, small, and crafted operands
    pools = [rnd(rng, 150_000, True), rnd(rng, 50_000, False), synth_u64_leaves(1, 20_000), crafted]
    a = np.concatenate(pools)
    b = np.concatenate([rnd(rng, 150_000, True), rnd(rng, 50_000, False), synth_u64_leaves(2, 20_000), crafted[::-1].copy()])
    for op, name, args in ((1, "mul", (a, b)), (2, "sqr", (a,)), (3, "pow5", (a,))):
        got, fl = run(op, *args)
        want = oracle.batch_fr(name, *args)
        ok = (got == want).all(axis=1)
        assert (ok | (fl != 0)).all(), f"fast {name} wrong without a flag at {np.nonzero(~ok & (fl == 0))[0][:5]}"
        assert fl.mean() < 1e-2, name


def test_million_pair_hashes_compared_in_full(gpu, oracle):
    """BASELINE configs[0] size: 1,000,000 pair hashes (the bench's seeded stream), EVERY output compared with the CPU
    implementation running on all host threads -- not a sample, so that a rare fast-path slip could not hide."""
    import os

    import torch

    from cuzk_b200 import lib

    L = lib.get_lib()
    n = 1_000_000
    dl = torch.empty((n, 4), dtype=torch.int64, device="cuda")
    dr = torch.empty_like(dl)
    out = torch.empty_like(dl)
    L.check(L.cuzk_synth_elements(dl.data_ptr(), n, 1, 0, 1, None), "synth")
    L.check(L.cuzk_synth_elements(dr.data_ptr(), n, 2, 0, 1, None), "synth")
    before = L.cuzk_debug_fallback_count()
    L.check(L.cuzk_poseidon_hash_pairs(dl.data_ptr(), dr.data_ptr(), out.data_ptr(), n, 0, None), "pairs")
    got = to_host(out)
    taken = L.cuzk_debug_fallback_count() - before
    l, r = synth_elements(1, n), synth_elements(2, n)
    assert (to_host(dl[:1000]) == l[:1000]).all()
    threads = max(1, len(os.sched_getaffinity(0)))
    want = oracle.hash_pairs_mt(l, r, threads)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, (bad[:5], taken)
    assert taken < 1000, f"{taken} of 1e6 hashes took the exact fallback: the fast path should decide almost always"


def test_large_host_batches_take_the_staged_path(gpu, oracle):
    """Host-buffer calls on pageable memory (numpy here, std::vector in the C++ layer) of more than 4 MiB go through the pinned
    bounce buffers and the copy pool, several chunks deep, in place for the permutation: same results as the device path."""
    rng = np.random.default_rng(314)
    n = 700_000                                   # 2 inputs + 1 output of 22 MB each: 2 chunks of the field-op pipeline
    a, b = rnd(rng, n, False), rnd(rng, n, False)
    ops = _ops(gpu)
    for name in ("add", "sub"):
        got = ops[name](a, b)
        assert (got == oracle.batch_fr(name, a, b)).all(), name
    h = gpu.CudaPoseidonHash()
    m = 45_000                                    # 4.3 MB of states, permuted in place through the staged pipeline
    st = rnd(rng, 3 * m, False)
    want = oracle.permutation(st.reshape(-1, 3, 4))
    got = h.batch_permutation(st.copy())
    assert (got == want).all()
    k = 400_000                                   # pair hashes from pageable memory: 4 chunks; spot-check against the oracle
    l, r = synth_elements(41, k), synth_elements(42, k)
    out = h.batch_hash_pairs(l, r)
    sel = np.unique(np.r_[0, k - 1, 113_663, 113_664, 227_327, 227_328, rng.integers(0, k, 500)])
    assert (out[sel] == oracle.hash_pairs(l[sel], r[sel])).all()
    dev = to_host(h.batch_hash_pairs(to_dev(l), to_dev(r)))
    assert (out == dev).all()


def test_gpu_against_the_compiled_reference_itself(gpu, ref):
    """The other tests compare with the plain-C port (which the CPU suite pins to the reference); this one compares the GPU
    with the reference's own CPU classes, compiled unmodified into oracle/_ref/libcuzk_ref.so, directly: PoseidonHash pair
    hashes and sponges on arbitrary 256-bit inputs, the permutation, and a NaryMerkleTree (root, proofs, verification)."""
    rng = np.random.default_rng(2718)
    h = gpu.CudaPoseidonHash()
    n = 3000
    l, r = rnd(rng, n, False), rnd(rng, n, False)
    l[:EDGE.shape[0]] = EDGE
    assert (to_host(h.batch_hash_pairs(to_dev(l), to_dev(r))) == ref.hash_pairs(l, r)).all()
    for width in (1, 3, 4, 7, 8):
        x = rnd(rng, 200 * width, False)
        assert (to_host(h.batch_sponge(to_dev(x), width, 3)) == ref.sponge(x, width, 3)).all(), width
    st = rnd(rng, 3 * 500, False).reshape(-1, 3, 4)
    assert (to_host(h.batch_permutation(to_dev(st.copy()))) == ref.permutation(st)).all()
    for arity, nleaves in ((2, 1000), (4, 777), (8, 4097)):
        leaves = synth_u64_leaves(90 + arity, nleaves)
        cpu = ref.tree(leaves, arity)
        t = gpu.CudaNaryMerkleTree(to_dev(leaves), arity=arity)
        assert (t.get_root_hash().reshape(-1) == cpu.root()).all(), arity
        idx = np.unique(np.r_[0, nleaves - 1, rng.integers(0, nleaves, 40)])
        pb = t.generate_batch_proofs(idx)
        sib, pos = to_host(pb.siblings), pb.positions.cpu().numpy().view(np.uint32)
        for k, i in enumerate(idx):
            rs, rp = cpu.prove(int(i))
            assert (sib[k] == rs).all() and (pos[k] == rp).all(), (arity, i)
            assert cpu.verify(leaves[i], sib[k], pos[k].astype(np.uint64), cpu.root())      # the reference accepts the GPU's proof
        assert bool(t.verify_batch_proofs(pb, to_dev(leaves[idx])).cpu().numpy().all())     # and the GPU the reference's (same bytes)


@pytest.mark.parametrize("direct", [True, False])
def test_small_host_calls_direct_on_pinned_memory_and_staged_agree(gpu, oracle, direct):
    """Host-buffer calls of at most 1 MiB run their kernel directly on pinned host memory (the caller's own when it is pinned,
    a bounce copy of pageable memory); cuzk_debug_set_direct_max(0) forces the staged copy path.  Same results either way, in
    place for the permutation, for pinned and pageable callers, for the cooperative and the one-thread kernels."""
    import torch

    from cuzk_b200.lib import get_lib

    L = get_lib()
    rng = np.random.default_rng(77)
    prev = L.cuzk_debug_set_direct_max((1 << 20) if direct else 0)
    try:
        h = gpu.CudaPoseidonHash()
        for n in (1, 37, 4096, 9000):
            l, r = rnd(rng, n, False), rnd(rng, n, False)
            want = oracle.hash_pairs(l, r)
            assert (h.batch_hash_pairs(l, r) == want).all(), ("pageable", n)          # numpy: pageable
            pl, pr = torch.from_numpy(l.view(np.int64)).pin_memory(), torch.from_numpy(r.view(np.int64)).pin_memory()
            po = torch.empty_like(pl).pin_memory()
            L.check(L.cuzk_poseidon_hash_pairs(pl.data_ptr(), pr.data_ptr(), po.data_ptr(), n, 1, None), "pairs")   # CUZK_MEM_HOST
            assert (po.numpy().view(np.uint64) == want).all(), ("pinned", n)
            # pinned inputs, pageable output
            out = np.zeros((n, 4), dtype=np.uint64)
            L.check(L.cuzk_poseidon_hash_pairs(pl.data_ptr(), pr.data_ptr(), out.ctypes.data, n, 1, None), "pairs")
            assert (out == want).all(), ("mixed", n)
        m = 1500
        st = rnd(rng, 3 * m, False)
        want = oracle.permutation(st.reshape(-1, 3, 4))
        assert (h.batch_permutation(st.copy()) == want).all()
        ps = torch.from_numpy(st.view(np.int64).copy()).pin_memory()
        L.check(L.cuzk_poseidon_permutation(ps.data_ptr(), m, 1, None), "permutation")
        assert (ps.numpy().view(np.uint64).reshape(-1, 3, 4) == want).all()
        x = rnd(rng, 5 * 700, False)
        assert (h.batch_sponge(x, 5, 3) == oracle.sponge(x, 5, 3)).all()
    finally:
        L.cuzk_debug_set_direct_max(prev)


@pytest.mark.parametrize("arity", [2, 3, 8])
def test_device_tree_append_equals_fresh_build(gpu, oracle, arity):
    """cuzk_tree_append_leaves (NaryMerkleTree::insert_leaf, batched): appending within the padded level (path re-hash),
    exactly filling it, and overflowing it (device rebuild) all give the oracle's tree over the concatenated leaves."""
    rng = np.random.default_rng(200 + arity)
    leaves = rnd(rng, arity**3 - 5, True)
    for dev in (False, True):
        t = gpu.DeviceMerkleTree(to_dev(leaves) if dev else leaves, arity=arity)
        cur = leaves
        for extra in (1, 3, 1, arity**3, 7):          # within, up to the brim, over it (twice), within the larger tree
            add = rnd(rng, extra, extra % 2 == 0)
            t.append_leaves(to_dev(add) if dev else add)
            cur = np.concatenate([cur, add])
            want = oracle.merkle_build(cur, arity)
            got = t.get_tree_levels()
            assert t.leaf_count == cur.shape[0]
            assert len(got) == len(want) and all((g == w).all() for g, w in zip(got, want)), (arity, dev, extra)
        idx = np.array([0, cur.shape[0] - 1], dtype=np.uint64)
        pb = t.generate_batch_proofs(to_dev(idx.astype(np.int64)) if dev else idx)
        res = t.verify_batch_proofs(pb, to_dev(cur[idx.astype(np.int64)]) if dev else cur[idx.astype(np.int64)])
        assert bool((res.cpu().numpy() if dev else res).all())
        t.close()


def test_long_sponges(gpu, oracle):
    """hash_multiple over long inputs (the reference's sponge has no length limit): widths 65, 200 and 1001, host and device."""
    h = gpu.CudaPoseidonHash()
    rng = np.random.default_rng(65)
    for width, n in ((65, 40), (200, 12), (1001, 3)):
        x = rnd(rng, width * n, False)
        want = oracle.sponge(x, width, 3)
        assert (h.batch_sponge(x, width, 3) == want).all(), width
        assert (to_host(h.batch_sponge(to_dev(x), width, 3)) == want).all(), width
