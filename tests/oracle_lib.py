"""ctypes doors into the parity oracle (oracle/libcuzk_oracle.so) and, when present, the
compiled reference CPU implementation (oracle/_ref/libcuzk_ref.so).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module.  Nothing under cuzk_b200/ does.

Elements are numpy uint64 arrays of shape (..., 4): little-endian limbs, the memory layout
of the reference's `struct FieldElement` (src/poseidon/field_arithmetic.hpp:11-14).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libcuzk_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libcuzk_ref.so")

P_INT = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
K_INT = 0x0E0A77C19A07DF2F666EA36F7879462E36FC76959F60CD29AC96341C4FFFFFFB

_u64p = C.POINTER(C.c_uint64)


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


def to_limbs(x: int) -> np.ndarray:
    return np.array([(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def from_limbs(a) -> int:
    a = np.asarray(a, dtype=np.uint64).reshape(-1)
    return sum(int(a[i]) << (64 * i) for i in range(4))


def ints_to_array(xs) -> np.ndarray:
    out = np.zeros((len(xs), 4), dtype=np.uint64)
    for i, x in enumerate(xs):
        out[i] = to_limbs(x)
    return out


def array_to_ints(a) -> list[int]:
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    return [from_limbs(r) for r in a]


def hexes(a) -> list[str]:
    return ["%064x" % v for v in array_to_ints(a)]


def build_oracle() -> None:
    """(Re)build the oracle libraries when sources are newer or libraries are missing."""
    need = not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
        os.path.join(ORACLE_DIR, "cuzk_oracle.c")
    )
    have_ref_src = os.path.isdir("/root/reference/src")
    need_ref = have_ref_src and (
        not os.path.exists(REF_SO)
        or os.path.getmtime(REF_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "ref_shim.cpp"))
    )
    if need:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle"], stdout=subprocess.DEVNULL)
    if need_ref:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL)


class _Lib:
    """Common surface of the oracle ('cuzk_oracle_') and the compiled reference ('cuzk_ref_')."""

    OPS = {"add": 0, "sub": 1, "mul": 2, "sqr": 3, "pow5": 4}

    def __init__(self, path: str, prefix: str):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.path = path

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    # ---- Fr ----
    def batch_fr(self, op: str, a: np.ndarray, b: np.ndarray | None = None) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        if b is None:
            b = a
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
        r = np.empty_like(a)
        f = self._f("batch_fr")
        f.argtypes = [C.c_int, _u64p, _u64p, _u64p, C.c_size_t]
        f.restype = None
        f(self.OPS[op], _p(a), _p(b), _p(r), a.shape[0])
        return r

    def reduce_512(self, prod: np.ndarray) -> np.ndarray:
        prod = np.ascontiguousarray(prod, dtype=np.uint64).reshape(-1, 8)
        out = np.empty((prod.shape[0], 4), dtype=np.uint64)
        f = self._f("fr_reduce_512" if self.prefix == "cuzk_oracle_" else "reduce_512")
        f.argtypes = [_u64p, _u64p]
        f.restype = None
        for i in range(prod.shape[0]):
            f(_p(prod[i]), _p(out[i]))
        return out

    # ---- Poseidon ----
    def round_constants(self) -> np.ndarray:
        out = np.empty((192, 4), dtype=np.uint64)
        f = self._f("round_constants")
        f.argtypes = [_u64p]
        f.restype = None
        f(_p(out))
        return out

    def mds(self) -> np.ndarray:
        out = np.empty((9, 4), dtype=np.uint64)
        f = self._f("mds")
        f.argtypes = [_u64p]
        f.restype = None
        f(_p(out))
        return out

    def permutation(self, states: np.ndarray) -> np.ndarray:
        s = np.array(states, dtype=np.uint64).reshape(-1, 3, 4).copy()
        f = self._f("batch_permutation")
        f.argtypes = [_u64p, C.c_size_t]
        f.restype = None
        f(_p(s), s.shape[0])
        return s

    def hash_single(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 4)
        out = np.empty_like(x)
        f = self._f("batch_hash_single")
        f.argtypes = [_u64p, _u64p, C.c_size_t]
        f.restype = None
        f(_p(x), _p(out), x.shape[0])
        return out

    def hash_pairs(self, l: np.ndarray, r: np.ndarray) -> np.ndarray:
        l = np.ascontiguousarray(l, dtype=np.uint64).reshape(-1, 4)
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(-1, 4)
        assert l.shape == r.shape
        out = np.empty_like(l)
        f = self._f("batch_hash_pairs")
        f.argtypes = [_u64p, _u64p, _u64p, C.c_size_t]
        f.restype = None
        f(_p(l), _p(r), _p(out), l.shape[0])
        return out

    def hash_pairs_mt(self, l: np.ndarray, r: np.ndarray, threads: int) -> np.ndarray:
        l = np.ascontiguousarray(l, dtype=np.uint64).reshape(-1, 4)
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(-1, 4)
        out = np.empty_like(l)
        f = self._f("batch_hash_pairs_mt")
        f.argtypes = [_u64p, _u64p, _u64p, C.c_size_t, C.c_int]
        f.restype = None
        f(_p(l), _p(r), _p(out), l.shape[0], threads)
        return out

    def sponge(self, x: np.ndarray, width: int, ds: int) -> np.ndarray:
        """x: (n*width, 4) -> (n, 4); hash i absorbs x[i*width:(i+1)*width] with domain separator ds."""
        x = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 4)
        n = x.shape[0] // width if width else 0
        out = np.empty((n, 4), dtype=np.uint64)
        f = self._f("batch_sponge")
        f.argtypes = [_u64p, C.c_size_t, C.c_uint64, _u64p, C.c_size_t]
        f.restype = None
        f(_p(x), width, ds, _p(out), n)
        return out

    def empty_hash(self, arity: int) -> np.ndarray:
        out = np.empty(4, dtype=np.uint64)
        f = self._f("empty_hash")
        f.argtypes = [C.c_size_t, _u64p]
        f.restype = None
        f(arity, _p(out))
        return out


class Oracle(_Lib):
    def __init__(self):
        build_oracle()
        super().__init__(ORACLE_SO, "cuzk_oracle_")
        L = self.lib
        for name in ("padded_size", "num_levels", "total_nodes", "tree_height_float"):
            getattr(L, "cuzk_oracle_" + name).argtypes = [C.c_size_t, C.c_size_t]
            getattr(L, "cuzk_oracle_" + name).restype = C.c_size_t

    def padded_size(self, n, a):
        return self.lib.cuzk_oracle_padded_size(n, a)

    def num_levels(self, n, a):
        return self.lib.cuzk_oracle_num_levels(n, a)

    def total_nodes(self, n, a):
        return self.lib.cuzk_oracle_total_nodes(n, a)

    def tree_height_float(self, n, a):
        return self.lib.cuzk_oracle_tree_height_float(n, a)

    def merkle_build(self, leaves: np.ndarray, arity: int) -> list[np.ndarray]:
        """All levels of the padded tree: [padded leaves, ..., root]."""
        leaves = np.ascontiguousarray(leaves, dtype=np.uint64).reshape(-1, 4)
        n = leaves.shape[0]
        assert n >= 1
        tot = self.total_nodes(n, arity)
        flat = np.empty((tot, 4), dtype=np.uint64)
        f = self.lib.cuzk_oracle_merkle_build
        f.argtypes = [_u64p, C.c_size_t, C.c_size_t, _u64p]
        f.restype = None
        f(_p(leaves), n, arity, _p(flat))
        out, off, p = [], 0, self.padded_size(n, arity)
        while True:
            out.append(flat[off : off + p])
            off += p
            if p == 1:
                break
            p //= arity
        return out

    def merkle_root(self, leaves: np.ndarray, arity: int) -> np.ndarray:
        leaves = np.ascontiguousarray(leaves, dtype=np.uint64).reshape(-1, 4)
        root = np.empty(4, dtype=np.uint64)
        f = self.lib.cuzk_oracle_merkle_root
        f.argtypes = [_u64p, C.c_size_t, C.c_size_t, _u64p]
        f.restype = None
        f(_p(leaves), leaves.shape[0], arity, _p(root))
        return root

    def merkle_prove(self, levels: list[np.ndarray], n: int, arity: int, index: int):
        flat = np.ascontiguousarray(np.concatenate(levels, axis=0))
        nl = len(levels) - 1
        sib = np.zeros((max(nl, 1), arity - 1, 4), dtype=np.uint64)
        pos = np.zeros(max(nl, 1), dtype=np.uint64)
        f = self.lib.cuzk_oracle_merkle_prove
        f.argtypes = [_u64p, C.c_size_t, C.c_size_t, C.c_size_t, _u64p, _u64p]
        f.restype = C.c_long
        got = f(_p(flat), n, arity, index, _p(sib), _p(pos))
        if got < 0:
            return None
        return sib[:got], pos[:got]

    def merkle_verify(self, leaf, sib, pos, arity: int, root) -> bool:
        leaf = np.ascontiguousarray(leaf, dtype=np.uint64).reshape(4)
        sib = np.ascontiguousarray(sib, dtype=np.uint64)
        pos = np.ascontiguousarray(pos, dtype=np.uint64)
        root = np.ascontiguousarray(root, dtype=np.uint64).reshape(4)
        f = self.lib.cuzk_oracle_merkle_verify
        f.argtypes = [_u64p, _u64p, _u64p, C.c_size_t, C.c_size_t, _u64p]
        f.restype = C.c_int
        return bool(f(_p(leaf), _p(sib), _p(pos), pos.shape[0], arity, _p(root)))


class Ref(_Lib):
    """The reference's own CPU code, compiled unmodified (oracle/Makefile target `ref`)."""

    def __init__(self):
        build_oracle()
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        super().__init__(REF_SO, "cuzk_ref_")
        L = self.lib
        L.cuzk_ref_tree_new.argtypes = [_u64p, C.c_size_t, C.c_size_t]
        L.cuzk_ref_tree_new.restype = C.c_void_p
        L.cuzk_ref_tree_free.argtypes = [C.c_void_p]
        L.cuzk_ref_tree_root.argtypes = [C.c_void_p, _u64p]
        L.cuzk_ref_tree_get_height.argtypes = [C.c_void_p]
        L.cuzk_ref_tree_get_height.restype = C.c_size_t
        L.cuzk_ref_tree_prove.argtypes = [C.c_void_p, C.c_size_t, _u64p, _u64p]
        L.cuzk_ref_tree_prove.restype = C.c_long
        L.cuzk_ref_tree_verify.argtypes = [C.c_void_p, _u64p, _u64p, _u64p, C.c_size_t, _u64p]
        L.cuzk_ref_tree_verify.restype = C.c_int
        L.cuzk_ref_tree_height.argtypes = [C.c_size_t, C.c_size_t]
        L.cuzk_ref_tree_height.restype = C.c_size_t
        L.cuzk_ref_generate_test_leaves.argtypes = [C.c_size_t, C.c_uint64, _u64p]
        L.cuzk_ref_benchmark_poseidon_pairs.argtypes = [C.c_size_t]
        L.cuzk_ref_benchmark_poseidon_pairs.restype = C.c_double
        L.cuzk_ref_tree_build_ms.argtypes = [_u64p, C.c_size_t, C.c_size_t, _u64p]
        L.cuzk_ref_tree_build_ms.restype = C.c_double

    def tree_height_float(self, n, a):
        return self.lib.cuzk_ref_tree_height(n, a)

    def generate_test_leaves(self, count: int, seed: int = 0) -> np.ndarray:
        out = np.empty((count, 4), dtype=np.uint64)
        self.lib.cuzk_ref_generate_test_leaves(count, seed, _p(out))
        return out

    def benchmark_poseidon_pairs(self, n: int) -> float:
        return self.lib.cuzk_ref_benchmark_poseidon_pairs(n)

    def tree_build_ms(self, leaves: np.ndarray, arity: int):
        leaves = np.ascontiguousarray(leaves, dtype=np.uint64).reshape(-1, 4)
        root = np.empty(4, dtype=np.uint64)
        ms = self.lib.cuzk_ref_tree_build_ms(_p(leaves), leaves.shape[0], arity, _p(root))
        return ms, root

    class Tree:
        def __init__(self, ref: "Ref", leaves: np.ndarray, arity: int):
            self.ref, self.arity = ref, arity
            leaves = np.ascontiguousarray(leaves, dtype=np.uint64).reshape(-1, 4)
            self.n = leaves.shape[0]
            self.h = ref.lib.cuzk_ref_tree_new(_p(leaves) if self.n else None, self.n, arity)

        def __del__(self):
            if getattr(self, "h", None):
                self.ref.lib.cuzk_ref_tree_free(self.h)
                self.h = None

        def root(self) -> np.ndarray:
            out = np.empty(4, dtype=np.uint64)
            self.ref.lib.cuzk_ref_tree_root(self.h, _p(out))
            return out

        def height(self) -> int:
            return self.ref.lib.cuzk_ref_tree_get_height(self.h)

        def prove(self, index: int, max_levels: int = 64):
            sib = np.zeros((max_levels, self.arity - 1, 4), dtype=np.uint64)
            pos = np.zeros(max_levels, dtype=np.uint64)
            got = self.ref.lib.cuzk_ref_tree_prove(self.h, index, _p(sib), _p(pos))
            if got < 0:
                return None
            return sib[:got].copy(), pos[:got].copy()

        def verify(self, leaf, sib, pos, root) -> bool:
            leaf = np.ascontiguousarray(leaf, dtype=np.uint64).reshape(4)
            sib = np.ascontiguousarray(sib, dtype=np.uint64)
            pos = np.ascontiguousarray(pos, dtype=np.uint64)
            root = np.ascontiguousarray(root, dtype=np.uint64).reshape(4)
            return bool(self.ref.lib.cuzk_ref_tree_verify(self.h, _p(leaf), _p(sib), _p(pos), pos.shape[0], _p(root)))

    def tree(self, leaves, arity) -> "Ref.Tree":
        return Ref.Tree(self, leaves, arity)


def have_ref() -> bool:
    try:
        build_oracle()
    except Exception:
        pass
    return os.path.exists(REF_SO)


# ---- seeded synthetic inputs: counter-based, reproducible on host and device ----
def splitmix64(seed: int, idx: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser of (seed * 0x9E3779B97F4A7C15 + idx + 1) * golden; vectorised, uint64."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) * np.uint64(0xD1342543DE82EF95) + (idx.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_elements(seed: int, n: int, start: int = 0, canonical: bool = True) -> np.ndarray:
    """n field elements: limb j of element i = splitmix64(seed, 4*(start+i)+j); top limb masked to 60 bits
    (< 2^252 < p) when canonical."""
    idx = np.arange(4 * start, 4 * (start + n), dtype=np.uint64)
    a = splitmix64(seed, idx).reshape(n, 4)
    if canonical:
        a[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    return a


def synth_u64_leaves(seed: int, n: int, start: int = 0) -> np.ndarray:
    """n leaves FieldElement(splitmix64(seed, start+i)) -- 64-bit leaves like the reference's test data."""
    a = np.zeros((n, 4), dtype=np.uint64)
    a[:, 0] = splitmix64(seed, np.arange(start, start + n, dtype=np.uint64))
    return a


def oracle_mds_layer(oracle: "Oracle", states: np.ndarray) -> np.ndarray:
    """One MDS layer (apply_mds_matrix) on (n, 3, 4) states via the oracle."""
    s = np.array(states, dtype=np.uint64).reshape(-1, 3, 4).copy()
    f = oracle.lib.cuzk_oracle_batch_mds_layer
    f.argtypes = [_u64p, C.c_size_t]
    f.restype = None
    f(_p(s), s.shape[0])
    return s
