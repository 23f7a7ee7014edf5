"""Shared helpers for the parity tests."""
import numpy as np

from oracle_lib import ints_to_array


def h2a(hex_list):
    """list of 64-digit hex strings -> (n, 4) uint64"""
    return ints_to_array([int(h, 16) for h in hex_list]) if len(hex_list) else np.zeros((0, 4), dtype=np.uint64)


def to_dev(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def to_host(t):
    return t.cpu().numpy().view(np.uint64)


def rnd(rng, n, canonical):
    a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    if canonical:
        a[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    return a
