"""ctypes binding of include/cuzk_b200.h -- exactly the stub a reference-side maintainer would write
(INTEGRATION.md shows the C++ one).  Pointers are raw integers: ``tensor.data_ptr()`` for device
memory (torch is only the allocator / stream provider) or numpy buffers for host memory.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libcuzk_b200.so")

FR_ADD, FR_SUB, FR_MUL, FR_SQR, FR_POW5 = 0, 1, 2, 3, 4
MEM_DEVICE, MEM_HOST = 0, 1

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-O3",
    "-std=c++17",
    "-shared",
    "-Xcompiler",
    "-fPIC",
]


class CuzkError(RuntimeError):
    pass


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def library_is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = _sources() + [os.path.join(ROOT, "include", "cuzk_b200.h")]
    return any(os.path.getmtime(s) > t for s in deps)


def _compile(out_path: str, extra: list[str], verbose: bool = False) -> None:
    nvcc = os.environ.get("NVCC", "nvcc")
    cu = [s for s in _sources() if s.endswith(".cu")]
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out_path] + cu
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise CuzkError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile libcuzk_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    if not force and not library_is_stale():
        return LIB_PATH
    _compile(LIB_PATH, [], verbose)
    return LIB_PATH


# Test-only build of the same sources in which the fast path's "undecided comparison" flags also fire on near misses
# (top 12 bits equal instead of all 32), so that a large share of the units takes the exact fallback path.  The
# results must not change; tests/test_gpu_parity.py runs the parity checks against this library too.
DEBUG_LIB_PATH = os.path.join(PKG_DIR, "libcuzk_b200_widen.so")


def build_debug_library(force: bool = False) -> str:
    deps = _sources() + [os.path.join(ROOT, "include", "cuzk_b200.h")]
    fresh = os.path.exists(DEBUG_LIB_PATH) and all(os.path.getmtime(s) <= os.path.getmtime(DEBUG_LIB_PATH) for s in deps)
    if force or not fresh:
        _compile(DEBUG_LIB_PATH, ["-DCUZK_UNC_WIDEN=20"])
    return DEBUG_LIB_PATH


_SIGS = {
    "cuzk_init": (C.c_int, [C.c_int]),
    "cuzk_shutdown": (C.c_int, []),
    "cuzk_is_initialized": (C.c_int, []),
    "cuzk_is_initialized_on": (C.c_int, [C.c_int]),
    "cuzk_device_count": (C.c_int, []),
    "cuzk_device_info": (C.c_int, [C.c_int, C.c_void_p]),
    "cuzk_last_error": (C.c_char_p, []),
    "cuzk_version": (C.c_char_p, []),
    "cuzk_launch_count": (C.c_uint64, []),
    "cuzk_debug_set_coop_max": (C.c_size_t, [C.c_size_t]),
    "cuzk_debug_set_coop_wide_max": (C.c_size_t, [C.c_size_t]),
    "cuzk_debug_set_build_plan": (None, [C.c_int, C.c_int, C.c_size_t]),
    "cuzk_debug_set_direct_max": (C.c_size_t, [C.c_size_t]),
    "cuzk_fr_batch": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "cuzk_poseidon_hash_single": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "cuzk_poseidon_hash_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "cuzk_poseidon_permutation": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "cuzk_poseidon_sponge": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "cuzk_poseidon_constants": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cuzk_merkle_padded_leaves": (C.c_size_t, [C.c_size_t, C.c_uint]),
    "cuzk_merkle_num_levels": (C.c_size_t, [C.c_size_t, C.c_uint]),
    "cuzk_merkle_total_nodes": (C.c_size_t, [C.c_size_t, C.c_uint]),
    "cuzk_merkle_tree_height": (C.c_size_t, [C.c_size_t, C.c_uint]),
    "cuzk_merkle_empty_hash": (C.c_int, [C.c_uint, C.c_void_p]),
    "cuzk_merkle_build": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.c_int, C.c_void_p]),
    "cuzk_merkle_build_batch": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint, C.c_void_p, C.c_int, C.c_void_p]),
    "cuzk_merkle_subtree_roots": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint, C.c_uint, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]),
    "cuzk_merkle_top_root": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.c_int, C.c_void_p]),
    "cuzk_merkle_padding_root": (C.c_int, [C.c_uint, C.c_uint, C.c_void_p]),
    "cuzk_merkle_prove_batch": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "cuzk_merkle_verify_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "cuzk_tree_build": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "cuzk_tree_free": (C.c_int, [C.c_void_p]),
    "cuzk_tree_leaf_count": (C.c_size_t, [C.c_void_p]),
    "cuzk_tree_num_levels": (C.c_size_t, [C.c_void_p]),
    "cuzk_tree_total_nodes": (C.c_size_t, [C.c_void_p]),
    "cuzk_tree_arity": (C.c_uint, [C.c_void_p]),
    "cuzk_tree_device": (C.c_int, [C.c_void_p]),
    "cuzk_tree_oob_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "cuzk_tree_device_levels": (C.c_void_p, [C.c_void_p]),
    "cuzk_tree_root": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "cuzk_tree_levels": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "cuzk_tree_level": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]),
    "cuzk_tree_prove_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "cuzk_tree_verify_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "cuzk_tree_update_leaves": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "cuzk_tree_append_leaves": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "cuzk_mg_unique_id": (C.c_int, [C.c_void_p]),
    "cuzk_mg_init_local": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "cuzk_mg_init_rank": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "cuzk_mg_free": (C.c_int, [C.c_void_p]),
    "cuzk_mg_world": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cuzk_mg_device": (C.c_int, [C.c_void_p, C.c_int]),
    "cuzk_mg_stream": (C.c_void_p, [C.c_void_p, C.c_int]),
    "cuzk_mg_nccl_version": (C.c_int, []),
    "cuzk_mg_shard_leaves": (C.c_int, [C.c_size_t, C.c_uint, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "cuzk_mg_tree_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_int, C.POINTER(C.c_void_p)]),
    "cuzk_mg_tree_free": (C.c_int, [C.c_void_p]),
    "cuzk_mg_tree_root": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cuzk_mg_tree_num_levels": (C.c_size_t, [C.c_void_p]),
    "cuzk_mg_tree_leaf_count": (C.c_size_t, [C.c_void_p]),
    "cuzk_mg_tree_subtree_height": (C.c_size_t, [C.c_void_p]),
    "cuzk_mg_tree_shard_levels": (C.c_void_p, [C.c_void_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "cuzk_mg_tree_prove_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "cuzk_mg_tree_verify_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "cuzk_mg_poseidon_hash_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "cuzk_synth_elements": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]),
    "cuzk_synth_u64_leaves": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_void_p]),
    "cuzk_debug_fallback_count": (C.c_uint64, []),
    "cuzk_debug_fast_ops": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "cuzk_debug_mds_layer": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "cuzk_imad_peak": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double)]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)


class Lib:
    """Loaded libcuzk_b200.so with typed entry points; every call raises CuzkError on a non-zero status."""

    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise CuzkError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(cuzk_b200 has no CPU fallback)"
            )
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            fn = getattr(self.dll, name)
            fn.restype = res
            fn.argtypes = args

    def check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise CuzkError(f"{what} failed ({rc}): {self.dll.cuzk_last_error().decode()}")

    def __getattr__(self, name):
        return getattr(self.dll, name)


_lib = None
_lock = threading.Lock()


def get_lib() -> Lib:
    global _lib
    with _lock:
        if _lib is None:
            _lib = Lib()
        return _lib
