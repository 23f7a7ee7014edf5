"""CPU tests of the multi-GPU host logic (cuzk_b200/distributed.py): shard plans, and the subtree-root
all-gather + top levels under torch.distributed/gloo with world_size 2 and 3, the oracle standing in for the
GPU kernels (injected backend).  The real NCCL path is the same code with CudaOps."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cuzk_b200.distributed import plan_merkle_shards, shard_slice, sharded_merkle_root
from oracle_lib import Oracle, synth_u64_leaves


def test_plan_config4_matches_survey():
    p = plan_merkle_shards(2**26, 8, 8)
    assert (p.padded, p.height, p.span, p.total_subtrees, p.real_subtrees, p.per_rank) == (2**27, 7, 2**21, 64, 32, 4)
    assert p.rank_leaves(0) == (0, 2**23) and p.rank_leaves(7) == (7 * 2**23, 2**26)
    p = plan_merkle_shards(2**26, 8, 2)
    assert (p.height, p.real_subtrees, p.per_rank) == (8, 4, 2)
    p = plan_merkle_shards(2**26, 8, 1)
    assert (p.height, p.real_subtrees, p.total_subtrees) == (9, 1, 1)


@pytest.mark.parametrize("arity", [2, 3, 4, 8])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_plans_cover_all_leaves_once(arity, world):
    for n in (1, 2, 7, 64, 100, 1000, 4097, 2**16):
        p = plan_merkle_shards(n, arity, world)
        assert p.span == arity**p.height and p.padded % p.span == 0
        covered = 0
        for r in range(world):
            lo, hi = p.rank_leaves(r)
            assert lo == covered or lo == hi == n
            covered = max(covered, hi)
            s0, s1 = p.rank_subtrees(r)
            assert s1 - s0 <= p.per_rank
        assert covered == n
        assert p.per_rank * world >= p.real_subtrees


def test_shard_slice():
    for n in (0, 1, 10, 1_000_003):
        for w in (1, 2, 8):
            parts = [shard_slice(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))


class OracleOps:
    """CPU stand-in for CudaOps: same contract, hashing done by the oracle."""

    def __init__(self):
        self.o = Oracle()

    def subtree_roots(self, leaves, n_local, arity, height, count):
        out = np.zeros((count, 4), dtype=np.uint64)
        lv = leaves.numpy().view(np.uint64)[:n_local]
        span = arity**height
        for i in range(count):
            part = lv[i * span : (i + 1) * span]
            if part.shape[0] == 0:
                out[i] = self.padding_root(arity, height).numpy().view(np.uint64).reshape(-1)
                continue
            # subtree root with virtual padding: build over `span` slots
            full = np.concatenate([part, np.tile(self.o.empty_hash(arity), (span - part.shape[0], 1))]) if part.shape[0] < span else part
            cur = full
            while cur.shape[0] > 1:
                cur = self.o.sponge(cur, arity, 3)
            out[i] = cur[0]
        return torch.from_numpy(out.view(np.int64))

    def top_root(self, nodes, arity):
        cur = nodes.numpy().view(np.uint64)
        while cur.shape[0] > 1:
            cur = self.o.sponge(np.ascontiguousarray(cur), arity, 3)
        return torch.from_numpy(cur.view(np.int64).reshape(1, 4).copy())

    def padding_root(self, arity, height):
        e = self.o.empty_hash(arity)
        for _ in range(height):
            e = self.o.sponge(np.tile(e, (arity, 1)), arity, 3)[0]
        return torch.from_numpy(e.view(np.int64).reshape(1, 4).copy())

    def empty(self, k):
        return torch.empty((k, 4), dtype=torch.int64)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, arity, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = plan_merkle_shards(n, arity, world)
        l0, l1 = plan.rank_leaves(rank)
        local = torch.from_numpy(synth_u64_leaves(4, l1 - l0, start=l0).view(np.int64)) if l1 > l0 else torch.empty((0, 4), dtype=torch.int64)
        root = sharded_merkle_root(local, plan, rank, OracleOps())
        q.put((rank, root.numpy().view(np.uint64).reshape(-1).tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,arity", [(2, 300, 8), (2, 64, 4), (3, 100, 2), (2, 5, 3)])
def test_sharded_root_equals_single_tree_gloo(world, n, arity):
    want = Oracle().merkle_root(synth_u64_leaves(4, n), arity).tolist()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, arity, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, root in got:
        assert root == want, (rank, world, n, arity)


def _verify_worker(rank, world, port, nproofs, bad_at, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cuzk_b200.distributed import shard_slice, sharded_all_valid

        lo, hi = shard_slice(nproofs, rank, world)
        verdicts = np.ones(hi - lo, dtype=np.uint8)          # stands in for this rank's merkle_verify_kernel results
        if bad_at is not None and lo <= bad_at < hi:
            verdicts[bad_at - lo] = 0
        q.put((rank, sharded_all_valid(verdicts)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nproofs,bad_at,want", [(100, None, True), (100, 73, False), (1, None, True), (0, None, False), (3, 0, False)])
def test_sharded_verify_and_reduce_gloo(nproofs, bad_at, want):
    """batch verification shards the proofs; the global answer is the AND over ranks (False for an empty batch)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_verify_worker, args=(r, world, port, nproofs, bad_at, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [g for _, g in got] == [want] * world
