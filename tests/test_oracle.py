"""CPU tests: pin the oracle (oracle/cuzk_oracle.c) against the golden vectors generated from the
reference CPU implementation (tests/golden/, made by generate_golden.py), against the compiled reference
itself when oracle/_ref is present, and against the known answers recorded in SURVEY.md Appendix B."""
import numpy as np
import pytest

from conftest import load_golden
from helpers import h2a, rnd
from oracle_lib import K_INT, P_INT, array_to_ints, hexes, ints_to_array, synth_elements


# ---- Appendix B known answers (generated from the reference CPU code during the survey) ----
def test_appendix_b_known_answers(oracle):
    perm = hexes(oracle.permutation(ints_to_array([1, 2, 3])).reshape(-1, 4))
    assert perm == [
        "07b845866686a60a43f75f0cd778887cc9c304376fcd0b3de6964e45b9630501",
        "0ef091199adbccb5a4f16d125495a5088efad30e7157b84e7429c087d234c932",
        "157a12c9c56ae74429660dfb6aebdf9148e6afb977080be9c424ccb07472ae04",
    ]
    single = {1: "1cfb848a89ee158e7c30178b8fdd4259c93f3ba4c72a3991e2bc6c08444117b1",
              42: "066e59aed12901e110f7d8459d3c2fa7705b3ce5a5eb1c7593e7e1465f85dafb",
              P_INT - 1: "29f0e823803b1b68fa2f390c0c40f342ef7a6b056eff31fd088dc1ce373710fc",
              2**256 - 1: "153a3a960c0ace0674ca2e37add196bb1a2e725b47231deffaf3374a2f224a21"}
    for x, want in single.items():
        assert hexes(oracle.hash_single(ints_to_array([x])))[0] == want
    pairs = {(0, 0): "165f8b7df1ea95746c405cda67aee6c254867d1a66af04a9bec5693b4093c38d",
             (1, 2): "1e65b90b79908fd07065fe718f4bb09be0424411567cb80d34b9a501bdf51a75",
             (10, 20): "2dd359f92d31c747e06c02b360a9f5c761777b285edcf09724efef5cbd51d9ba"}
    for (l, r), want in pairs.items():
        assert hexes(oracle.hash_pairs(ints_to_array([l]), ints_to_array([r])))[0] == want
    assert hexes(oracle.sponge(ints_to_array([1, 2, 3, 4]), 4, 3))[0] == "2c12b96d3926e4862876ae9ca67cddad85313fa6fa5f266fb7ab683826a6a497"
    assert hexes(oracle.sponge(np.zeros((0, 4), dtype=np.uint64), 0, 3)) == []
    empty = {2: "194324f01efa21d2dcdd7453800fde166a852e2906e0e6de5de6921eeb77feec",
             3: "1c7842d7703c243a99d6e6ca4033851791b5ae206220fc8c9bcdde10e5befbdd",
             4: "1c7842d7703c243a99d6e6ca4033851791b5ae206220fc8c9bcdde10e5befbdd",
             5: "23e5a7fd958be849ffe194cb94e567d674239bffe6f7076708685d63fd40ccb3",
             6: "23e5a7fd958be849ffe194cb94e567d674239bffe6f7076708685d63fd40ccb3",
             7: "2ca165c9c68473c20eb293f63de5986e10a90fb68f6e54bd7932e5166048445d",
             8: "2ca165c9c68473c20eb293f63de5986e10a90fb68f6e54bd7932e5166048445d"}
    for a, want in empty.items():
        assert hexes(oracle.empty_hash(a))[0] == want


def test_round_constants_fit_64_bits(oracle):
    rc = oracle.round_constants()
    assert (rc[:, 1:] == 0).all()
    assert hexes(rc[[0, 1, 2, 191]]) == ["%064x" % v for v in (0x0123456789ABCDEF, 0x02468AD89ABCDEFF, 0x0369D049ABCDF00F, 0xDA7414C3456788DF)]
    assert [int(v) for v in oracle.mds()[:, 0]] == [7, 23, 8, 26, 5, 4, 15, 20, 9]


# ---- golden fixtures ----
def test_golden_field_ops(oracle):
    g = load_golden("fr_ops.json")
    a, b = h2a(g["a"]), h2a(g["b"])
    for op in ("add", "sub", "mul", "sqr", "pow5"):
        assert hexes(oracle.batch_fr(op, a, b)) == g[op], op
    prod = np.array([[(int(h, 16) >> (64 * j)) & (2**64 - 1) for j in range(8)] for h in g["reduce_512_in"]], dtype=np.uint64)
    assert hexes(oracle.reduce_512(prod)) == g["reduce_512_out"]


def test_golden_poseidon(oracle):
    g = load_golden("poseidon.json")
    assert hexes(oracle.round_constants()) == g["round_constants"]
    assert hexes(oracle.mds()) == g["mds"]
    assert hexes(oracle.permutation(h2a(g["perm_in"]).reshape(-1, 3, 4)).reshape(-1, 4)) == g["perm_out"]
    assert hexes(oracle.hash_single(h2a(g["single_in"]))) == g["single_out"]
    assert hexes(oracle.hash_pairs(h2a(g["pair_l"]), h2a(g["pair_r"]))) == g["pair_out"]
    for case in g["sponge"]:
        assert hexes(oracle.sponge(h2a(case["in"]), case["width"], case["ds"])) == case["out"], case["width"]
    for a, want in g["empty_hash"].items():
        assert hexes(oracle.empty_hash(int(a)))[0] == want


def _golden_leaves(ent, ref_leaves_seed42):
    if "leaves" in ent:
        return h2a(ent["leaves"])
    return ref_leaves_seed42(ent["n"])


def mt19937_64_leaves(n, seed=42):
    """MerkleUtils::generate_test_leaves (merkle_tree.cpp:448-460): FieldElement(std::mt19937_64(seed)())."""
    # numpy's MT19937 is the 32-bit generator; implement the 64-bit variant directly.
    NN, MM = 312, 156
    UM, LM = 0xFFFFFFFF80000000, 0x7FFFFFFF
    mask = (1 << 64) - 1
    mt = [0] * NN
    mt[0] = seed & mask
    for i in range(1, NN):
        mt[i] = (6364136223846793005 * (mt[i - 1] ^ (mt[i - 1] >> 62)) + i) & mask
    out = np.zeros((n, 4), dtype=np.uint64)
    idx = NN
    for k in range(n):
        if idx >= NN:
            for i in range(NN):
                x = (mt[i] & UM) | (mt[(i + 1) % NN] & LM)
                mt[i] = mt[(i + MM) % NN] ^ (x >> 1) ^ (0xB5026F5AA96619E9 if x & 1 else 0)
            idx = 0
        x = mt[idx]
        idx += 1
        x ^= (x >> 29) & 0x5555555555555555
        x ^= (x << 17) & 0x71D67FFFEDA60000
        x ^= (x << 37) & 0xFFF7EEE000000000
        x ^= x >> 43
        out[k, 0] = np.uint64(x & mask)
    return out


def test_mt19937_64_matches_reference_leaves():
    g = load_golden("merkle.json")
    assert hexes(mt19937_64_leaves(8, 42)) == g["leaves_seed42_first8"]


def test_golden_merkle(oracle):
    g = load_golden("merkle.json")
    for ent in g["trees"]:
        leaves = _golden_leaves(ent, lambda n: mt19937_64_leaves(n, 42))
        a, n = ent["arity"], ent["n"]
        levels = oracle.merkle_build(leaves, a)
        assert hexes(levels[-1])[0] == ent["root"], (a, n)
        assert oracle.tree_height_float(n, a) == ent["height"]
        for pr in ent["proofs"]:
            sib, pos = oracle.merkle_prove(levels, n, a, pr["index"])
            assert [int(v) for v in pos] == pr["positions"]
            assert hexes(sib.reshape(-1, 4)) == pr["siblings"]
            assert oracle.merkle_verify(leaves[pr["index"]], sib, pos, a, levels[-1][0])
    for a, want in g["empty_root"].items():
        assert hexes(oracle.merkle_root(np.zeros((0, 4), dtype=np.uint64), int(a)))[0] == want
    for n, a, h in g["height_table"]:
        if n >= 1:
            assert oracle.tree_height_float(n, a) == h, (n, a)


def test_float_height_overcounts_at_exact_powers(oracle):
    """the getter's float formula is off by one at some exact powers, so levels must
    be sized with the integer loop."""
    assert oracle.num_levels(8**7, 8) == 8
    assert oracle.tree_height_float(8**7, 8) in (8, 9)
    assert oracle.num_levels(125, 5) == 4
    assert oracle.tree_height_float(125, 5) == 5  # reference quirk
    assert oracle.padded_size(2**26, 8) == 2**27 and oracle.num_levels(2**26, 8) == 10


# ---- oracle vs the compiled reference (present in the build container and shipped to the GPU box) ----
def test_oracle_vs_reference_field_ops(oracle, ref):
    rng = np.random.default_rng(7)
    edge = ints_to_array([0, 1, P_INT - 1, P_INT, P_INT + 1, 2**256 - 1, 2**128 - 1, K_INT, 5 * P_INT, 5 * P_INT + 1, 2**64 - 1])
    for canonical in (True, False):
        a = np.concatenate([rnd(rng, 4000, canonical), edge, edge[::-1]])
        b = np.concatenate([rnd(rng, 4000, canonical), edge[::-1], edge])
        for op in ("add", "sub", "mul", "sqr", "pow5"):
            assert (oracle.batch_fr(op, a, b) == ref.batch_fr(op, a, b)).all(), (op, canonical)


def test_oracle_vs_reference_hashes(oracle, ref):
    rng = np.random.default_rng(8)
    x, y = rnd(rng, 300, False), rnd(rng, 300, True)
    assert (oracle.hash_single(x) == ref.hash_single(x)).all()
    assert (oracle.hash_pairs(x, y) == ref.hash_pairs(x, y)).all()
    st = rnd(rng, 300, False).reshape(-1, 3, 4)
    assert (oracle.permutation(st) == ref.permutation(st)).all()
    for w in range(0, 9):
        z = rnd(rng, 16 * w, False) if w else np.zeros((0, 4), dtype=np.uint64)
        assert (oracle.sponge(z, w, 3) == ref.sponge(z, w, 3)).all()


def test_oracle_vs_reference_merkle(oracle, ref):
    for arity in range(2, 9):
        for n in (1, 2, 6, 7, 31, 50):
            leaves = ref.generate_test_leaves(n, 7)
            t = ref.tree(leaves, arity)
            levels = oracle.merkle_build(leaves, arity)
            assert (levels[-1][0] == t.root()).all()
            for i in range(n):
                so, po = oracle.merkle_prove(levels, n, arity, i)
                sr, pr = t.prove(i)
                assert (so == sr).all() and (po == pr).all()
                assert t.verify(leaves[i], so, po, levels[-1][0])
            assert oracle.merkle_prove(levels, n, arity, n) is None and t.prove(n) is None


def test_multiply_is_not_modular(oracle):
    """The property that makes Montgomery/Barrett unusable: the reference multiply differs from a*b mod p."""
    a = synth_elements(11, 64)
    b = synth_elements(12, 64)
    got = array_to_ints(oracle.batch_fr("mul", a, b))
    true = [(x * y) % P_INT for x, y in zip(array_to_ints(a), array_to_ints(b))]
    assert sum(g != t for g, t in zip(got, true)) > 32
    # ... and equals the Appendix A formula
    W = 1 << 256

    def red(x):
        return x % P_INT

    def model(x, y):
        P = x * y
        low, high = P % W, P // W
        Mh = high * K_INT
        ml, mh = Mh % W, Mh // W
        t = (ml + (mh * K_INT) % W) % W
        hc = red(t) if mh else t
        return red((low + hc) % W)

    assert got == [model(x, y) for x, y in zip(array_to_ints(a), array_to_ints(b))]


def craft_mds_states(rng):
    """Canonical states that hit every special case of the GPU's fast MDS layer (cuzk_b200/csrc/poseidon.cuh):
    wrap thresholds of every (constant, h), ambiguous carries into word 7, tiny and maximal values."""
    from oracle_lib import K_INT, P_INT

    W = 1 << 256
    vals = [0, 1, 2, P_INT - 1, P_INT - 2, P_INT // 2]
    for c in (7, 23, 8, 26, 5, 4, 15, 20, 9):
        for h in range(1, 5):
            T = (h + 1) * W - h * K_INT  # c*s >= T  <=>  the term wraps
            for d in range(-3, 4):
                vals.append((T + d * c + c - 1) // c)          # around the exact threshold
                vals.append(((h + 1) * W + d * c) // c)         # around the next multiple of W
            top = ((h << 32) | (0xFFFFFFFF - ((h * K_INT) >> 224))) << 224   # low_7 == threshold word, lower words vary
            for lowbits in (0, 1, (1 << 224) - 1, (1 << 223), int(rng.integers(0, 2**62)) << 160):
                vals.append((top + lowbits) // c)
        a2 = (c & -c).bit_length() - 1                                      # c = 2^a2 * odd
        inv = pow(c >> a2, -1, 1 << 32)
        for tgt in (0xFFFFFFE0, 0xFFFFFFF5, 0xFFFFFFFF, 0xFFFFFFDF):       # lo32(c * s_6) at the ambiguity limit
            s6 = (((tgt >> a2) * inv) & 0xFFFFFFFF) % (1 << (32 - a2))     # c * s6 = tgt rounded down to 2^a2 (mod 2^32)
            assert (c * s6) & 0xFFFFFFFF == (tgt >> a2) << a2
            for s7 in (0, 0x10000000, 0x30644E72, 0x2FFFFFFF):
                for low in (0, (1 << 192) - 1, int(rng.integers(0, 2**62)) << 128):
                    vals.append((s7 << 224) | (s6 << 192) | low)
    vals = [v for v in vals if 0 <= v < P_INT]
    picks = rng.integers(0, len(vals), size=(6000, 3))
    states = np.stack([ints_to_array([vals[i] for i in picks[:, j]]) for j in range(3)], axis=1)
    rand = rnd(rng, 3 * 3000, True).reshape(-1, 3, 4)
    rand[::3, 1] = ints_to_array([vals[i] for i in rng.integers(0, len(vals), size=len(rand[::3]))])
    return np.concatenate([states, rand])


def test_oracle_mds_layer_vs_reference_ops(oracle, ref):
    """The oracle's single MDS layer equals the composition of the reference's multiply/add (poseidon.cpp:148-167)."""
    from oracle_lib import oracle_mds_layer

    rng = np.random.default_rng(77)
    st = craft_mds_states(rng)[:3000]
    got = oracle_mds_layer(oracle, st)
    mds = ref.mds()
    for i in range(3):
        acc = np.zeros((st.shape[0], 4), dtype=np.uint64)
        for j in range(3):
            term = ref.batch_fr("mul", np.tile(mds[3 * i + j], (st.shape[0], 1)), st[:, j])
            acc = ref.batch_fr("add", acc, term)
        assert (got[:, i] == acc).all(), i
