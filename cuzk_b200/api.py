"""Host-side mirror of the reference's CUDA batch interfaces, over the C ABI.

Names and semantics follow the reference:
  CudaFieldArithmetic   src/poseidon/cuda/field_arithmetic_cuda.cuh:25-81
  CudaPoseidonHash      src/poseidon/cuda/poseidon_interface_cuda.hpp:27-47, poseidon_cuda.cuh:23-59
  CudaNaryMerkleTree    src/merkle_tree/merkle_tree_cuda.cuh:39-106

Arrays of field elements are either numpy ``uint64`` arrays of shape (n, 4) (host memory: the call stages
through the library's device buffers, like the reference's std::vector entry points) or torch CUDA
tensors of dtype int64 and shape (n, 4) (device memory: asynchronous on the current torch stream).
Limbs are little-endian, plain canonical form -- the byte layout of the reference's FieldElement.
"""
from __future__ import annotations

import numpy as np

from .lib import FR_ADD, FR_MUL, FR_POW5, FR_SQR, FR_SUB, MEM_DEVICE, MEM_HOST, CuzkError, get_lib

try:  # torch is plumbing (device memory, streams); the host-array path works without it
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_tensor(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


def _ptr(x):
    if x is None:
        return None
    if _is_tensor(x):
        assert x.is_contiguous()
        return x.data_ptr()
    assert x.flags["C_CONTIGUOUS"]
    return x.ctypes.data


def _stream(x):
    if _is_tensor(x) and x.is_cuda:
        return torch.cuda.current_stream(x.device).cuda_stream
    return None


def _mem(x) -> int:
    if _is_tensor(x):
        if not x.is_cuda:
            raise CuzkError("torch tensors must live on the GPU; pass numpy arrays for host memory")
        return MEM_DEVICE
    return MEM_HOST


def _elems(x, cols: int = 4):
    """Normalise to a contiguous (n, cols) array of 64-bit words."""
    if _is_tensor(x):
        if x.dtype not in (torch.int64, torch.uint64):
            raise CuzkError("device elements must be int64/uint64 tensors")
        return x.reshape(-1, cols).contiguous()
    a = np.ascontiguousarray(x, dtype=np.uint64)
    return a.reshape(-1, cols)


def _empty_like(x, n: int, cols: int = 4):
    if _is_tensor(x):
        return torch.empty((n, cols), dtype=x.dtype, device=x.device)
    return np.empty((n, cols), dtype=np.uint64)


_init_count = 0


def initialize(device: int | None = None) -> None:
    """CudaFieldArithmetic::initialize / CudaNaryMerkleTree::initialize_cuda: idempotent, ref-counted."""
    global _init_count
    lib = get_lib()
    if device is None:
        device = torch.cuda.current_device() if (torch is not None and torch.cuda.is_available()) else 0
    lib.check(lib.cuzk_init(int(device)), "cuzk_init")
    _init_count += 1


def cleanup() -> None:
    """CudaFieldArithmetic::cleanup: drops one reference; never resets the device."""
    global _init_count
    if _init_count > 0:
        get_lib().cuzk_shutdown()
        _init_count -= 1


class CudaFieldArithmetic:
    """Element-wise batch field operations (reference semantics, arbitrary 256-bit inputs allowed)."""

    @staticmethod
    def _binary(op, a, b):
        a, b = _elems(a), _elems(b)
        if a.shape != b.shape:
            raise CuzkError("size mismatch")  # reference returns false (field_arithmetic_cuda.cu:366-370)
        lib = get_lib()
        out = _empty_like(a, a.shape[0])
        lib.check(lib.cuzk_fr_batch(op, _ptr(a), _ptr(b), _ptr(out), a.shape[0], _mem(a), _stream(a)), "cuzk_fr_batch")
        return out

    @staticmethod
    def _unary(op, a):
        a = _elems(a)
        lib = get_lib()
        out = _empty_like(a, a.shape[0])
        lib.check(lib.cuzk_fr_batch(op, _ptr(a), None, _ptr(out), a.shape[0], _mem(a), _stream(a)), "cuzk_fr_batch")
        return out

    @staticmethod
    def batch_add(a, b):
        return CudaFieldArithmetic._binary(FR_ADD, a, b)

    @staticmethod
    def batch_subtract(a, b):
        return CudaFieldArithmetic._binary(FR_SUB, a, b)

    @staticmethod
    def batch_multiply(a, b):
        return CudaFieldArithmetic._binary(FR_MUL, a, b)

    @staticmethod
    def batch_square(a):
        return CudaFieldArithmetic._unary(FR_SQR, a)

    @staticmethod
    def batch_power5(a):
        return CudaFieldArithmetic._unary(FR_POW5, a)

    @staticmethod
    def get_device_count() -> int:
        return get_lib().cuzk_device_count()


class CudaPoseidonHash:
    """IPoseidonCudaHash over the C ABI (one implementation backs both reference classes)."""

    def __init__(self, device: int | None = None):
        initialize(device)
        self._live = True

    def close(self):
        if self._live:
            cleanup()
            self._live = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def is_initialized(self) -> bool:
        return self._live and bool(get_lib().cuzk_is_initialized())

    @staticmethod
    def get_optimal_batch_size() -> int:
        return 148 * 6 * 128  # one resident wave: 148 SMs x 6 CTAs x 128 threads

    @staticmethod
    def get_max_batch_size() -> int:
        return 1 << 31

    def batch_hash_single(self, inputs):
        x = _elems(inputs)
        lib = get_lib()
        out = _empty_like(x, x.shape[0])
        lib.check(lib.cuzk_poseidon_hash_single(_ptr(x), _ptr(out), x.shape[0], _mem(x), _stream(x)), "cuzk_poseidon_hash_single")
        return out

    def batch_hash_pairs(self, left, right, out=None):
        l, r = _elems(left), _elems(right)
        if l.shape != r.shape:
            raise CuzkError("size mismatch")  # reference returns false (poseidon_cuda.cu:342-345)
        lib = get_lib()
        if out is None:
            out = _empty_like(l, l.shape[0])
        lib.check(
            lib.cuzk_poseidon_hash_pairs(_ptr(l), _ptr(r), _ptr(out), l.shape[0], _mem(l), _stream(l)), "cuzk_poseidon_hash_pairs"
        )
        return out

    def batch_permutation(self, states):
        """In place on (n, 3, 4) states; returns the same array."""
        s = _elems(states, 12)
        lib = get_lib()
        lib.check(lib.cuzk_poseidon_permutation(_ptr(s), s.shape[0], _mem(s), _stream(s)), "cuzk_poseidon_permutation")
        return s.reshape(-1, 3, 4)

    def batch_sponge(self, inputs, width: int, domain_separator: int = 3):
        """out[i] = sponge(inputs[i*width:(i+1)*width], domain_separator); 3 = hash_multiple."""
        x = _elems(inputs)
        n = x.shape[0] // width if width else 0
        lib = get_lib()
        out = _empty_like(x, n)
        lib.check(
            lib.cuzk_poseidon_sponge(_ptr(x), width, domain_separator, _ptr(out), n, _mem(x), _stream(x)), "cuzk_poseidon_sponge"
        )
        return out

    @staticmethod
    def constants():
        lib = get_lib()
        rc = np.empty((192, 4), dtype=np.uint64)
        mds = np.empty((9, 4), dtype=np.uint64)
        lib.check(lib.cuzk_poseidon_constants(rc.ctypes.data, mds.ctypes.data), "cuzk_poseidon_constants")
        return rc, mds


# ---- Merkle geometry (pure host arithmetic exported by the library) ----
def padded_leaves(n: int, arity: int) -> int:
    return get_lib().cuzk_merkle_padded_leaves(n, arity)


def num_levels(n: int, arity: int) -> int:
    return get_lib().cuzk_merkle_num_levels(n, arity)


def total_nodes(n: int, arity: int) -> int:
    return get_lib().cuzk_merkle_total_nodes(n, arity)


def tree_height(n: int, arity: int) -> int:
    return get_lib().cuzk_merkle_tree_height(n, arity)


def empty_hash(arity: int) -> np.ndarray:
    lib = get_lib()
    out = np.empty(4, dtype=np.uint64)
    lib.check(lib.cuzk_merkle_empty_hash(arity, out.ctypes.data), "cuzk_merkle_empty_hash")
    return out


def padding_root(arity: int, height: int) -> np.ndarray:
    lib = get_lib()
    out = np.empty(4, dtype=np.uint64)
    lib.check(lib.cuzk_merkle_padding_root(arity, height, out.ctypes.data), "cuzk_merkle_padding_root")
    return out


def build_batch_trees(batch_leaves, arity: int = 2):
    """CudaNaryMerkleTree::build_batch_trees for equal-sized trees: ``batch_leaves`` is (T, n, 4); returns T trees built by
    one forest pass (cuzk_merkle_build_batch)."""
    x = _elems(batch_leaves)
    T = int(batch_leaves.shape[0])
    n = x.shape[0] // T
    lib = get_lib()
    tot = total_nodes(n, arity)
    out = _empty_like(x, tot * T)
    lib.check(lib.cuzk_merkle_build_batch(_ptr(x), n, T, arity, _ptr(out), _mem(x), _stream(x)), "cuzk_merkle_build_batch")
    trees = []
    for t in range(T):
        tree = CudaNaryMerkleTree(arity=arity)
        tree.leaf_count, tree.levels = n, out[t * tot : (t + 1) * tot]
        trees.append(tree)
    return trees


class MerkleProofBatch:
    """Flat, level-uniform proof batch: siblings (q, L, arity-1, 4), positions (q, L) uint32, leaf indices (q,)."""

    def __init__(self, siblings, positions, indices, arity):
        self.siblings, self.positions, self.indices, self.arity = siblings, positions, indices, arity

    def __len__(self):
        return int(self.positions.shape[0])


class CudaNaryMerkleTree:
    """CudaNaryMerkleTree (merkle_tree_cuda.cuh:39-106): build on the GPU, keep every level for proof serving.

    ``device=True`` keeps the level arrays in HBM (torch tensor) and takes/returns device tensors;
    otherwise leaves and levels are host numpy arrays like the reference's std::vector members.
    """

    MIN_ARITY, MAX_ARITY = 2, 8

    def __init__(self, leaves=None, arity: int = 2, device: bool | None = None):
        if not (self.MIN_ARITY <= arity <= self.MAX_ARITY):
            # MerkleTreeConfig ctor -> ValidationError (std::invalid_argument), merkle_tree.hpp:27-31
            raise ValueError(f"arity must be between 2 and 8, got {arity}")
        initialize()
        self.arity = arity
        self.leaf_count = 0
        self.levels = None  # flat (total_nodes, 4)
        self._device = device
        if leaves is not None:
            self.build_tree(leaves)

    # -- build --
    def build_tree(self, leaves) -> bool:
        x = _elems(leaves)
        n = x.shape[0]
        if n == 0:  # merkle_tree_cuda.cu:143-148: empty input clears the tree
            self.leaf_count, self.levels = 0, None
            return True
        lib = get_lib()
        tot = total_nodes(n, self.arity)
        out = _empty_like(x, tot)
        lib.check(lib.cuzk_merkle_build(_ptr(x), n, self.arity, _ptr(out), _mem(x), _stream(x)), "cuzk_merkle_build")
        self.leaf_count, self.levels = n, out
        return True

    # -- getters --
    def get_leaf_count(self) -> int:
        return self.leaf_count

    def get_arity(self) -> int:
        return self.arity

    def get_tree_height(self) -> int:
        """The reference's float formula (getter only): 0 for an empty tree (merkle_tree.cpp:311-316)."""
        return 0 if self.leaf_count == 0 else tree_height(self.leaf_count, self.arity)

    def get_root_hash(self):
        if self.leaf_count == 0:
            return empty_hash(self.arity)  # merkle_tree.cpp:304-309
        r = self.levels[-1]
        return r.cpu().numpy().view(np.uint64) if _is_tensor(r) else r.copy()

    def get_tree_levels(self):
        """List of per-level arrays, level 0 = padded leaves, last = root (merkle_tree_cuda.cuh:89)."""
        out, off, p = [], 0, padded_leaves(self.leaf_count, self.arity)
        if self.leaf_count == 0:
            return out
        while True:
            out.append(self.levels[off : off + p])
            off += p
            if p == 1:
                break
            p //= self.arity
        return out

    # -- proofs --
    def generate_batch_proofs(self, indices) -> MerkleProofBatch:
        lib = get_lib()
        L = num_levels(self.leaf_count, self.arity) - 1 if self.leaf_count else 0
        dev = _is_tensor(self.levels)
        if dev:
            idx = torch.as_tensor(indices, dtype=torch.int64, device=self.levels.device).contiguous()
            q = idx.numel()
            sib = torch.empty((q, L, self.arity - 1, 4), dtype=self.levels.dtype, device=idx.device)
            pos = torch.empty((q, L), dtype=torch.int32, device=idx.device)
        else:
            idx = np.ascontiguousarray(indices, dtype=np.uint64)
            q = idx.size
            sib = np.empty((q, L, self.arity - 1, 4), dtype=np.uint64)
            pos = np.empty((q, L), dtype=np.uint32)
        if q and L:
            lib.check(
                lib.cuzk_merkle_prove_batch(
                    _ptr(self.levels), self.leaf_count, self.arity, _ptr(idx), q, _ptr(sib), _ptr(pos),
                    MEM_DEVICE if dev else MEM_HOST, _stream(self.levels),
                ),
                "cuzk_merkle_prove_batch",
            )
        return MerkleProofBatch(sib, pos, idx, self.arity)

    def verify_batch_proofs(self, proofs: MerkleProofBatch, leaf_values, root=None):
        """Per-proof results (uint8).  The reference returns the AND; see ``all_valid``."""
        lib = get_lib()
        lv = _elems(leaf_values)
        q = lv.shape[0]
        if q != len(proofs):
            raise CuzkError("size mismatch")
        L = int(proofs.positions.shape[1])
        dev = _is_tensor(lv)
        if root is None:
            root = self.levels[-1:] if self.leaf_count else empty_hash(self.arity)
        if dev:
            root_t = root if _is_tensor(root) else torch.from_numpy(np.asarray(root).view(np.int64)).to(lv.device)
            root_t = root_t.reshape(-1).contiguous()
            res = torch.empty(q, dtype=torch.uint8, device=lv.device)
        else:
            root_t = np.ascontiguousarray(root.cpu().numpy() if _is_tensor(root) else root).view(np.uint64).reshape(-1)
            res = np.empty(q, dtype=np.uint8)
        if q:
            lib.check(
                lib.cuzk_merkle_verify_batch(
                    _ptr(lv), _ptr(proofs.siblings), _ptr(proofs.positions), L, self.arity, _ptr(root_t), _ptr(res), q,
                    MEM_DEVICE if dev else MEM_HOST, _stream(lv),
                ),
                "cuzk_merkle_verify_batch",
            )
        return res

    def all_valid(self, proofs, leaf_values, root=None) -> bool:
        """verify_batch_proofs as the reference CUDA class returns it: false on an empty batch
        (merkle_tree_cuda.cu:343), else the AND of all results."""
        if len(proofs) == 0:
            return False
        res = self.verify_batch_proofs(proofs, leaf_values, root)
        return bool(res.all())


class DeviceMerkleTree:
    """Device-resident tree handle (cuzk_tree_*): the level arrays stay in HBM, proofs are generated and verified against
    them there, and leaves can be updated incrementally (path re-hash instead of the reference's full rebuild,
    merkle_tree.cpp:290-301).  Inputs/outputs are numpy arrays (host) or int64 CUDA tensors (device), as elsewhere."""

    def __init__(self, leaves, arity: int = 2):
        import ctypes as C

        if not (2 <= arity <= 8):
            raise ValueError(f"arity must be between 2 and 8, got {arity}")
        initialize()
        x = _elems(leaves)
        if x.shape[0] == 0:
            raise CuzkError("DeviceMerkleTree needs at least one leaf")
        lib = get_lib()
        h = C.c_void_p()
        lib.check(lib.cuzk_tree_build(_ptr(x), x.shape[0], arity, _mem(x), _stream(x), C.byref(h)), "cuzk_tree_build")
        self._h, self.arity, self.leaf_count = h, arity, x.shape[0]
        self.num_levels = lib.cuzk_tree_num_levels(h)
        self.total_nodes = lib.cuzk_tree_total_nodes(h)

    def close(self):
        if getattr(self, "_h", None):
            get_lib().cuzk_tree_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_root_hash(self) -> np.ndarray:
        out = np.empty(4, dtype=np.uint64)
        lib = get_lib()
        lib.check(lib.cuzk_tree_root(self._h, out.ctypes.data, MEM_HOST, None), "cuzk_tree_root")
        return out

    def get_tree_levels(self):
        """every level as host arrays (one download), level 0 = padded leaves, last = root"""
        flat = np.empty((self.total_nodes, 4), dtype=np.uint64)
        lib = get_lib()
        lib.check(lib.cuzk_tree_levels(self._h, flat.ctypes.data, MEM_HOST, None), "cuzk_tree_levels")
        out, off, p = [], 0, padded_leaves(self.leaf_count, self.arity)
        while True:
            out.append(flat[off : off + p])
            off += p
            if p == 1:
                return out
            p //= self.arity

    def generate_batch_proofs(self, indices) -> MerkleProofBatch:
        lib = get_lib()
        L = self.num_levels - 1
        dev = _is_tensor(indices)
        if dev:
            idx = indices.to(torch.int64).contiguous()
            q = idx.numel()
            sib = torch.empty((q, L, self.arity - 1, 4), dtype=torch.int64, device=idx.device)
            pos = torch.empty((q, L), dtype=torch.int32, device=idx.device)
        else:
            idx = np.ascontiguousarray(indices, dtype=np.uint64)
            q = idx.size
            sib = np.empty((q, L, self.arity - 1, 4), dtype=np.uint64)
            pos = np.empty((q, L), dtype=np.uint32)
        if q and L:
            lib.check(lib.cuzk_tree_prove_batch(self._h, _ptr(idx), q, _ptr(sib), _ptr(pos), MEM_DEVICE if dev else MEM_HOST,
                                                _stream(idx)), "cuzk_tree_prove_batch")
        return MerkleProofBatch(sib, pos, idx, self.arity)

    def verify_batch_proofs(self, proofs: MerkleProofBatch, leaf_values):
        lib = get_lib()
        lv = _elems(leaf_values)
        q = lv.shape[0]
        if q != len(proofs):
            raise CuzkError("size mismatch")
        dev = _is_tensor(lv)
        res = torch.empty(q, dtype=torch.uint8, device=lv.device) if dev else np.empty(q, dtype=np.uint8)
        if q:
            lib.check(lib.cuzk_tree_verify_batch(self._h, _ptr(lv), _ptr(proofs.siblings), _ptr(proofs.positions), _ptr(res), q,
                                                 MEM_DEVICE if dev else MEM_HOST, _stream(lv)), "cuzk_tree_verify_batch")
        return res

    def update_leaves(self, indices, values) -> None:
        """values[q] replaces leaf indices[q] (distinct indices < leaf_count); only the ancestors are re-hashed"""
        lib = get_lib()
        v = _elems(values)
        dev = _is_tensor(v)
        idx = indices.to(torch.int64).contiguous() if dev else np.ascontiguousarray(indices, dtype=np.uint64)
        lib.check(lib.cuzk_tree_update_leaves(self._h, _ptr(idx), _ptr(v), v.shape[0], MEM_DEVICE if dev else MEM_HOST, _stream(v)),
                  "cuzk_tree_update_leaves")

    def append_leaves(self, values) -> None:
        """NaryMerkleTree::insert_leaf for a batch: the tree afterwards equals a fresh build over all leaves"""
        lib = get_lib()
        v = _elems(values)
        lib.check(lib.cuzk_tree_append_leaves(self._h, _ptr(v), v.shape[0], _mem(v), _stream(v)), "cuzk_tree_append_leaves")
        self.leaf_count = lib.cuzk_tree_leaf_count(self._h)
        self.num_levels = lib.cuzk_tree_num_levels(self._h)
        self.total_nodes = lib.cuzk_tree_total_nodes(self._h)


class MultiGpu:
    """cuzk_mg_* handle: the hot path over several GPUs of one box, NCCL called from the library.

    ``MultiGpu.local(ngpus)`` -- this process drives ``ngpus`` devices.
    ``MultiGpu.from_torch_distributed()`` -- one process per GPU (torchrun): rank 0 creates the NCCL id, torch.distributed
    only carries its 128 bytes to the other ranks; the collective itself (one all-gather of subtree roots per build) is
    issued by the library."""

    def __init__(self, handle, nranks: int, nlocal: int, first_rank: int):
        self._h, self.nranks, self.nlocal, self.first_rank = handle, nranks, nlocal, first_rank

    @classmethod
    def local(cls, ngpus: int, devices=None) -> "MultiGpu":
        import ctypes as C

        lib = get_lib()
        h = C.c_void_p()
        arr = (C.c_int * ngpus)(*devices) if devices is not None else None
        lib.check(lib.cuzk_mg_init_local(ngpus, arr, C.byref(h)), "cuzk_mg_init_local")
        return cls(h, ngpus, ngpus, 0)

    @classmethod
    def from_torch_distributed(cls, device: int | None = None, group=None) -> "MultiGpu":
        import ctypes as C

        import torch.distributed as dist

        lib = get_lib()
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        else:
            rank, world = 0, 1      # a single process without a process group: one rank, no collective
        if device is None:
            device = torch.cuda.current_device()
        box = [None]
        if world > 1:
            if rank == 0:
                buf = (C.c_uint8 * 128)()
                lib.check(lib.cuzk_mg_unique_id(buf), "cuzk_mg_unique_id")
                box[0] = bytes(buf)
            dist.broadcast_object_list(box, src=0, group=group)
        h = C.c_void_p()
        idbuf = (C.c_uint8 * 128).from_buffer_copy(box[0]) if box[0] is not None else None
        lib.check(lib.cuzk_mg_init_rank(rank, world, int(device), idbuf, C.byref(h)), "cuzk_mg_init_rank")
        return cls(h, world, 1, rank)

    def close(self):
        if getattr(self, "_h", None):
            get_lib().cuzk_mg_free(self._h)
            self._h = None

    def device(self, local: int = 0) -> int:
        return get_lib().cuzk_mg_device(self._h, local)

    def stream(self, local: int = 0):
        return get_lib().cuzk_mg_stream(self._h, local)

    def shard_leaves(self, n: int, arity: int, rank: int) -> tuple[int, int]:
        """[first, first + count) of the leaves that shard `rank` holds in an n-leaf tree"""
        import ctypes as C

        first, count = C.c_size_t(), C.c_size_t()
        lib = get_lib()
        lib.check(lib.cuzk_mg_shard_leaves(n, arity, self.nranks, rank, C.byref(first), C.byref(count)), "cuzk_mg_shard_leaves")
        return first.value, count.value

    def batch_hash_pairs(self, left, right) -> np.ndarray:
        l, r = _elems(left), _elems(right)
        if l.shape != r.shape:
            raise CuzkError("left/right size mismatch")
        out = np.empty_like(l)
        lib = get_lib()
        lib.check(lib.cuzk_mg_poseidon_hash_pairs(self._h, l.ctypes.data, r.ctypes.data, out.ctypes.data, l.shape[0]), "cuzk_mg_poseidon_hash_pairs")
        return out

    def build_tree(self, leaves, n: int | None = None, arity: int = 2) -> "ShardedMerkleTree":
        """leaves: the whole leaf array in host memory (numpy), or a list with one int64 CUDA tensor per local device holding
        that shard's leaves (``shard_leaves``); ``n`` (total leaf count) is needed in the second form."""
        import ctypes as C

        lib = get_lib()
        h = C.c_void_p()
        if isinstance(leaves, (list, tuple)):
            if n is None:
                raise CuzkError("n (total leaf count) is needed with per-device leaves")
            ptrs = (C.c_void_p * self.nlocal)(*[(t.data_ptr() if t is not None and t.numel() else None) for t in leaves])
            lib.check(lib.cuzk_mg_tree_build(self._h, ptrs, n, arity, MEM_DEVICE, C.byref(h)), "cuzk_mg_tree_build")
        else:
            x = _elems(leaves)
            n = x.shape[0]
            ptrs = (C.c_void_p * 1)(x.ctypes.data)
            lib.check(lib.cuzk_mg_tree_build(self._h, ptrs, n, arity, MEM_HOST, C.byref(h)), "cuzk_mg_tree_build")
        return ShardedMerkleTree(h, n, arity)


class ShardedMerkleTree:
    """cuzk_mg_tree_*: every GPU keeps all levels of its subtrees and serves proofs from them; the top levels are replicated."""

    def __init__(self, handle, n: int, arity: int):
        lib = get_lib()
        self._h, self.leaf_count, self.arity = handle, n, arity
        self.num_levels = lib.cuzk_mg_tree_num_levels(handle)
        self.subtree_height = lib.cuzk_mg_tree_subtree_height(handle)

    def close(self):
        if getattr(self, "_h", None):
            get_lib().cuzk_mg_tree_free(self._h)
            self._h = None

    def get_root_hash(self) -> np.ndarray:
        out = np.empty(4, dtype=np.uint64)
        lib = get_lib()
        lib.check(lib.cuzk_mg_tree_root(self._h, out.ctypes.data), "cuzk_mg_tree_root")
        return out

    def generate_batch_proofs(self, indices) -> MerkleProofBatch:
        lib = get_lib()
        L = self.num_levels - 1
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        sib = np.zeros((idx.size, L, self.arity - 1, 4), dtype=np.uint64)
        pos = np.zeros((idx.size, L), dtype=np.uint32)
        if idx.size and L:
            lib.check(lib.cuzk_mg_tree_prove_batch(self._h, idx.ctypes.data, idx.size, sib.ctypes.data, pos.ctypes.data), "cuzk_mg_tree_prove_batch")
        return MerkleProofBatch(sib, pos, idx, self.arity)

    def verify_batch_proofs(self, proofs: MerkleProofBatch, leaf_values) -> np.ndarray:
        lib = get_lib()
        lv = _elems(leaf_values)
        if lv.shape[0] != len(proofs):
            raise CuzkError("size mismatch")
        res = np.zeros(lv.shape[0], dtype=np.uint8)
        if lv.shape[0]:
            lib.check(lib.cuzk_mg_tree_verify_batch(self._h, lv.ctypes.data, proofs.siblings.ctypes.data, proofs.positions.ctypes.data,
                                                    res.ctypes.data, lv.shape[0]), "cuzk_mg_tree_verify_batch")
        return res
