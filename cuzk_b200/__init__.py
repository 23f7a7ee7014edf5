"""cuzk_b200 -- B200-native (sm_100a) implementation of the davencyw/cuZK hot path:
batched Poseidon over BN254 Fr, n-ary Merkle build, batch proof generation and verification.

This Python package is host-side plumbing only (device memory via torch, torch.distributed for the
multi-GPU root gather).  All arithmetic happens in hand-written CUDA kernels inside
``libcuzk_b200.so`` behind the C ABI declared in ``include/cuzk_b200.h``; there is no CPU fallback and
importing :mod:`cuzk_b200.lib` fails loudly when the library has not been built.
"""
from .lib import (  # noqa: F401
    CuzkError,
    FR_ADD,
    FR_MUL,
    FR_POW5,
    FR_SQR,
    FR_SUB,
    LIB_PATH,
    Lib,
    build_library,
    get_lib,
)

__all__ = [
    "CuzkError",
    "Lib",
    "get_lib",
    "build_library",
    "LIB_PATH",
    "FR_ADD",
    "FR_SUB",
    "FR_MUL",
    "FR_SQR",
    "FR_POW5",
]
