"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for the plumbing.

The reference is single-GPU (no NCCL, no streams); this layer is new.

On GPUs the sharded tree lives in the LIBRARY (cuzk_mg_*, csrc/multi_gpu.cuh): every rank keeps all levels of its subtrees in
HBM and serves proofs from them, the library itself issues the one NCCL all-gather of subtree roots, and this module is a
thin caller (``native_sharded_tree``): torch.distributed only carries the 128-byte NCCL id at start-up.  The pure
torch.distributed formulation below (``sharded_merkle_root`` with an injected ``ops`` backend) is kept because it runs under
gloo on CPU with the oracle as hasher, which is how the sharding logic is tested without GPUs.

* Batch hashing shards by contiguous slices -- no collective.
* Merkle build: the padded tree is cut at level ``m`` into subtrees of ``arity**m`` leaves.  The subtrees that
  contain real leaves are dealt to the ranks in contiguous blocks, each rank reduces its block to subtree
  roots on its GPU (cuzk_merkle_subtree_roots; nothing but the roots leaves the GPU), ONE all-gather of
  32-byte roots crosses NVLink, and every rank hashes the few top levels (cuzk_merkle_top_root).
  All-padding subtrees are the level-``m`` padding constant and are never hashed or sent.
* Batch verify shards proofs; the result is an AND (all-reduce MIN of one byte).

The arithmetic backend is injected (``ops``) so the sharding logic runs under gloo on CPU in the tests with
the oracle as hasher, and on GPUs with the C ABI (``CudaOps``).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


@dataclass(frozen=True)
class ShardPlan:
    n: int                # real leaves
    arity: int
    world: int
    padded: int           # arity**L >= n   (integer loop; never the float height formula)
    height: int           # m: subtree height
    span: int             # arity**m leaves per subtree
    total_subtrees: int   # padded // span
    real_subtrees: int    # subtrees containing at least one real leaf
    per_rank: int         # subtree slots per rank in the gather buffer = ceil(real_subtrees / world)

    def rank_subtrees(self, rank: int) -> tuple[int, int]:
        """[first, last) subtree index owned by `rank` (may be empty for trailing ranks)."""
        lo = min(rank * self.per_rank, self.real_subtrees)
        hi = min(lo + self.per_rank, self.real_subtrees)
        return lo, hi

    def rank_leaves(self, rank: int) -> tuple[int, int]:
        """[first, last) REAL leaf index owned by `rank`."""
        lo, hi = self.rank_subtrees(rank)
        return min(lo * self.span, self.n), min(hi * self.span, self.n)


def plan_merkle_shards(n: int, arity: int, world: int, min_subtrees_per_rank: int = 1) -> ShardPlan:
    if n < 1 or not (2 <= arity <= 8) or world < 1:
        raise ValueError("bad shard plan arguments")
    padded, levels = 1, 0
    while padded < n:
        padded *= arity
        levels += 1
    # tallest subtrees that still give every rank at least `min_subtrees_per_rank` real subtrees
    height, span = 0, 1
    while height < levels:
        nxt = span * arity
        if -(-n // nxt) < world * min_subtrees_per_rank:
            break
        height, span = height + 1, nxt
    real = -(-n // span)
    return ShardPlan(n, arity, world, padded, height, span, padded // span, real, -(-real // world))


class CudaOps:
    """GPU backend over the C ABI; arrays are int64 CUDA tensors of shape (k, 4)."""

    def __init__(self, device):
        from . import api
        from .lib import get_lib

        self.api, self.lib, self.device = api, get_lib(), torch.device(device)
        api.initialize(self.device.index)

    def subtree_roots(self, leaves, n_local: int, arity: int, height: int, count: int):
        out = torch.empty((count, 4), dtype=torch.int64, device=self.device)
        if count:
            st = torch.cuda.current_stream(self.device).cuda_stream
            self.lib.check(
                self.lib.cuzk_merkle_subtree_roots(leaves.data_ptr() if n_local else None, n_local, arity, height, count,
                                                   out.data_ptr(), 0, st),
                "cuzk_merkle_subtree_roots",
            )
        return out

    def top_root(self, nodes, arity: int):
        out = torch.empty((1, 4), dtype=torch.int64, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        self.lib.check(self.lib.cuzk_merkle_top_root(nodes.data_ptr(), nodes.shape[0], arity, out.data_ptr(), 0, st), "cuzk_merkle_top_root")
        return out

    def padding_root(self, arity: int, height: int):
        return torch.from_numpy(self.api.padding_root(arity, height).view(np.int64)).to(self.device).reshape(1, 4)

    def empty(self, k: int):
        return torch.empty((k, 4), dtype=torch.int64, device=self.device)


def sharded_merkle_root(local_leaves, plan: ShardPlan, rank: int, ops, group=None):
    """Root of the n-leaf tree whose real leaves are dealt to ranks by ``plan``; ``local_leaves`` holds this
    rank's ``plan.rank_leaves(rank)`` slice.  Returns a (1, 4) tensor, identical on every rank."""
    lo, hi = plan.rank_subtrees(rank)
    l0, l1 = plan.rank_leaves(rank)
    mine = ops.subtree_roots(local_leaves, l1 - l0, plan.arity, plan.height, hi - lo)
    # fixed-size slot per rank so one all_gather_into_tensor moves everything (32 B x per_rank per rank)
    send = ops.empty(plan.per_rank)
    send.zero_()
    if hi > lo:
        send[: hi - lo] = mine
    if plan.world > 1:
        gathered = ops.empty(plan.per_rank * plan.world)
        dist.all_gather_into_tensor(gathered, send, group=group)
    else:
        gathered = send
    nodes = ops.empty(plan.total_subtrees)
    nodes[: plan.real_subtrees] = gathered[: plan.real_subtrees]   # rank blocks are contiguous and in order
    if plan.total_subtrees > plan.real_subtrees:
        nodes[plan.real_subtrees :] = ops.padding_root(plan.arity, plan.height)
    return ops.top_root(nodes, plan.arity)


def shard_slice(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice of n independent units (hashes, proofs) for `rank`."""
    return n * rank // world, n * (rank + 1) // world


def sharded_all_valid(local_results, group=None) -> bool:
    """AND of per-proof verdicts that were computed shard by shard (``shard_slice`` of the proof batch per rank): one
    all-reduce(MIN) of a single byte; an empty global batch is False, like CudaNaryMerkleTree::verify_batch_proofs
    (merkle_tree_cuda.cu:343).  ``local_results`` is this rank's uint8 verdict tensor (may be empty)."""
    t = local_results
    if not isinstance(t, torch.Tensor):
        t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.uint8))
    flags = torch.stack([(t.min() if t.numel() else torch.ones((), dtype=torch.uint8, device=t.device)).to(torch.int32),
                         torch.tensor(-int(t.numel() > 0), dtype=torch.int32, device=t.device)])
    # flags[0]: all of mine valid (1) or not (0) -> MIN;  flags[1]: -(I had proofs) -> MIN is -1 when any rank had some
    if dist is not None and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN, group=group)
    return bool(flags[0].item() == 1 and flags[1].item() == -1)


def native_sharded_tree(local_leaves, n: int, arity: int, mg=None, group=None):
    """The n-leaf tree sharded over the ranks of ``group`` through the library (cuzk_mg_tree_build): ``local_leaves`` is this
    rank's CUDA tensor of leaves (``mg.shard_leaves(n, arity, rank)``).  Returns (tree, mg); the tree serves proofs for this
    rank's leaves from the levels it keeps and verifies against the root every rank computed.  Collective."""
    from . import api

    if mg is None:
        mg = api.MultiGpu.from_torch_distributed(local_leaves.device.index, group=group)
    torch.cuda.synchronize(local_leaves.device)   # the library works on its own stream: the leaves must be complete
    return mg.build_tree([local_leaves], n=n, arity=arity), mg
