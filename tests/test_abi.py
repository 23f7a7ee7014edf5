"""CPU tests of the C-ABI boundary: the library builds/loads, exports every symbol include/cuzk_b200.h
declares, the pure-host geometry helpers agree with the oracle, and compute entry points fail loudly
(no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from cuzk_b200 import lib as cl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    cl.build_library()
    return cl.get_lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cuzk_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cuzk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(L):
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(L.dll, name), f"{name} declared in include/cuzk_b200.h but not exported"
    assert set(names) == set(cl.EXPORTED_SYMBOLS), "ctypes table and header disagree"


def test_library_is_sm100a_only(L):
    out = os.popen(f"cuobjdump -lelf {cl.LIB_PATH} 2>/dev/null").read()
    assert "sm_100a" in out


def test_geometry_matches_oracle(L, oracle):
    for arity in range(2, 9):
        for n in list(range(0, 70)) + [100, 125, 216, 1000, 4096, 50000, 2**20, 2**21, 8**7, 2**26]:
            if n:
                assert L.cuzk_merkle_padded_leaves(n, arity) == oracle.padded_size(n, arity)
                assert L.cuzk_merkle_tree_height(n, arity) == oracle.tree_height_float(n, arity)
            assert L.cuzk_merkle_num_levels(n, arity) == oracle.num_levels(n, arity)
            assert L.cuzk_merkle_total_nodes(n, arity) == oracle.total_nodes(n, arity)


def test_no_cpu_fallback(L):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present; the failure path is for GPU-less hosts")
    assert L.cuzk_device_count() == 0
    assert L.cuzk_init(0) != 0
    assert b"no CUDA device" in L.dll.cuzk_last_error()
    a = np.zeros((4, 4), dtype=np.uint64)
    out = np.zeros((4, 4), dtype=np.uint64)
    rc = L.cuzk_poseidon_hash_pairs(a.ctypes.data, a.ctypes.data, out.ctypes.data, 4, cl.MEM_HOST, None)
    assert rc != 0 and b"not initialised" in L.dll.cuzk_last_error()
    with pytest.raises(cl.CuzkError):
        L.check(rc, "cuzk_poseidon_hash_pairs")


def test_product_does_not_reference_oracle():
    """The product path must never import, link or call anything under oracle/."""
    for base, _, files in os.walk(os.path.join(ROOT, "cuzk_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(base, f)).read()
                assert "oracle_lib" not in text and "cuzk_oracle" not in text and "libcuzk_ref" not in text, f
    out = os.popen(f"ldd {cl.LIB_PATH}").read()
    assert "oracle" not in out and "cuzk_ref" not in out


def test_library_shard_plan_equals_the_python_plan(L):
    """cuzk_mg_shard_leaves (pure arithmetic, no GPU) deals the leaves of a sharded tree exactly like
    distributed.plan_merkle_shards, which the gloo tests check against the single-tree oracle: contiguous, in rank order,
    covering [0, n), cut at subtree boundaries."""
    from cuzk_b200.distributed import plan_merkle_shards

    rng = np.random.default_rng(12)
    cases = [(1, 2, 1), (1, 8, 8), (7, 2, 3), (4096, 8, 8), (4097, 8, 8), (1 << 20, 4, 4), (1 << 26, 8, 8), (50_000, 2, 2), (9, 3, 5)]
    cases += [(int(rng.integers(1, 1 << 22)), int(rng.integers(2, 9)), int(rng.integers(1, 9))) for _ in range(200)]
    for n, arity, world in cases:
        plan = plan_merkle_shards(n, arity, world)
        at = 0
        for rank in range(world):
            first, count = ctypes.c_size_t(), ctypes.c_size_t()
            L.check(L.cuzk_mg_shard_leaves(n, arity, world, rank, ctypes.byref(first), ctypes.byref(count)), "shard_leaves")
            lo, hi = plan.rank_leaves(rank)
            assert (first.value, count.value) == (lo, hi - lo), (n, arity, world, rank)
            assert count.value == 0 or first.value == at, (n, arity, world, rank)
            assert first.value % plan.span == 0 or count.value == 0
            at += count.value
        assert at == n, (n, arity, world)
    assert L.cuzk_mg_shard_leaves(0, 2, 1, 0, None, None) != 0 and L.cuzk_mg_shard_leaves(5, 9, 1, 0, None, None) != 0
