// merkle_tree/merkle_tree_cuda.cuh -- CudaNaryMerkleTree: GPU build, proofs and batch verification.
//
// Replaces the reference's src/merkle_tree/merkle_tree_cuda.cuh:39-151 (same class, benchmark helpers and utility
// namespace).  The two reference kernels (:24-36) are gone from the header: building, proving and verifying run in
// libcuzk_b200.so through cuzk_merkle_build / cuzk_merkle_verify_batch (include/cuzk_b200.h).
//
// Behaviour kept from the reference class: leaves are stored un-hashed and padded with empty_hash(arity) to a power
// of the arity; get_tree_levels() exposes every level on the host (level 0 = padded leaves, last = root);
// verify_proof compares against the tree's own root; verify_batch_proofs returns false for an empty batch or a
// size mismatch; get_tree_height() returns the reference's floating-point formula.
// Deliberate differences: the level arrays are sized by the integer padding loop, never by the float height (the
// reference builds a spurious extra level at exact powers); the whole tree is built in one
// host call (one upload; the levels stay in HBM until a caller asks for them) instead of a PCIe round trip per level; batches of every size are verified
// on the GPU (the reference verifies fewer than 32 proofs on the CPU, merkle_tree_cuda.cu:348-355).
#pragma once

#include <memory>
#include <optional>
#include <vector>

#include "../poseidon/cuda/cuda_field_element.cuh"
#include "../poseidon/cuda/poseidon_cuda.cuh"
#include "merkle_tree.hpp"

namespace MerkleTree {
namespace MerkleTreeCUDA {

using CudaFieldElement = Poseidon::CudaFieldElement;

// extension (no reference counterpart): a whole batch of proofs in the level-uniform wire format of the C ABI -- what a
// batch verifier wants instead of one heap-allocated vector per proof level.  For proof q and level l (0 = leaf level):
//   positions[q * levels + l]                               the own slot            (MerkleProof::indices[l])
//   siblings[(q * levels + l) * (arity - 1) + s]            the s-th other child    (MerkleProof::path[l][s])
struct FlatProofBatch {
  size_t arity = 0, levels = 0;
  std::vector<uint64_t> leaf_indices;
  std::vector<uint32_t> positions;
  std::vector<FieldElement> siblings;
  size_t size() const { return leaf_indices.size(); }
};

class CudaNaryMerkleTree {
public:
  // -- construction: the library is brought up on first use; a tree given leaves is built at once
  explicit CudaNaryMerkleTree(const MerkleTreeConfig &cfg = MerkleTreeConfig());
  explicit CudaNaryMerkleTree(const std::vector<FieldElement> &leaf_values, const MerkleTreeConfig &cfg = MerkleTreeConfig());
  // extension (no reference counterpart; opt-in): the tree sharded over `gpus` GPUs of the box through cuzk_mg_*.  The leaves
  // are cut into contiguous subtrees, one block per GPU; every GPU keeps ALL levels of its subtrees in HBM and serves proofs
  // from them; only the 32-byte subtree roots cross NVLink, in one NCCL all-gather issued by the library; the top levels are
  // hashed on every GPU.  Root, proofs and verdicts are those of the single-GPU tree.  get_tree_levels() stays empty: no
  // single device (and not the host) holds the whole tree.  build_tree() on such an object shards again.
  CudaNaryMerkleTree(const std::vector<FieldElement> &leaf_values, const MerkleTreeConfig &cfg, int gpus);
  ~CudaNaryMerkleTree();
  CudaNaryMerkleTree(CudaNaryMerkleTree &&) = default;
  CudaNaryMerkleTree &operator=(CudaNaryMerkleTree &&) = default;
  CudaNaryMerkleTree(const CudaNaryMerkleTree &) = default;
  CudaNaryMerkleTree &operator=(const CudaNaryMerkleTree &) = default;

  // -- GPU build: one upload, every level hashed on the device, the root comes back.  An empty vector clears the tree (true).
  bool build_tree(const std::vector<FieldElement> &leaf_values);
  // equal-sized trees are built as one forest pass (cuzk_merkle_build_batch); ragged batches fall back to one build per tree
  static bool build_batch_trees(const std::vector<std::vector<FieldElement>> &leaf_sets, std::vector<CudaNaryMerkleTree> &out_trees,
                                const MerkleTreeConfig &cfg = MerkleTreeConfig());

  // -- proofs: path[l] = the arity-1 siblings at level l in index order, indices[l] = own slot; nullopt past the last leaf
  std::optional<MerkleProof> generate_proof(size_t leaf) const;
  std::vector<MerkleProof> generate_batch_proofs(const std::vector<size_t> &leaf_indices) const;  // invalid indices skipped

  // -- verification on the GPU against this tree's root; the batch form is false for an empty batch or a size mismatch
  bool verify_proof(const MerkleProof &proof, const FieldElement &leaf_value) const;
  bool verify_batch_proofs(const std::vector<MerkleProof> &proofs, const std::vector<FieldElement> &leaf_values) const;
  // extension: one verdict per proof (1 = valid) instead of their conjunction
  bool verify_batch_proofs_each(const std::vector<MerkleProof> &proofs, const std::vector<FieldElement> &leaf_values,
                                std::vector<uint8_t> &verdicts) const;

  // extension: flat proof batches, generated and verified on the GPU without per-proof host objects
  bool generate_flat_proofs(const std::vector<size_t> &leaf_indices, FlatProofBatch &out) const;   // false on an invalid index
  bool verify_flat_proofs(const FlatProofBatch &batch, const std::vector<FieldElement> &leaf_values, std::vector<uint8_t> &verdicts) const;

  // -- getters (get_tree_height: the reference's floating-point formula; levels: 0 = padded leaves ... last = root)
  FieldElement get_root_hash() const;
  size_t get_arity() const { return config_.arity; }
  size_t get_leaf_count() const { return leaf_count_; }
  size_t get_tree_height() const { return tree_height_; }
  int get_gpu_count() const { return gpus_; }
  const std::vector<FieldElement> &get_leaves() const { return leaves_; }
  const std::vector<std::vector<FieldElement>> &get_tree_levels() const {
    fetch_levels();
    return tree_levels_;
  }
  void print_tree() const;

  // root / leaf count / arity equality with the reference's CPU tree.  A template so that this header does not need the
  // CPU class to be complete; it is instantiated only by callers that have the CPU tree (the reference's tests do).
  template <class CpuTree = NaryMerkleTree>
  bool compare_with_cpu_tree(const CpuTree &cpu) const {
    return get_leaf_count() == cpu.get_leaf_count() && get_arity() == cpu.get_arity() && get_root_hash() == cpu.get_root_hash();
  }

  // -- library lifetime and sizing hints
  static bool initialize_cuda();
  static void cleanup_cuda();
  static size_t get_optimal_batch_size();
  static size_t get_max_batch_size();

private:
  MerkleTreeConfig config_;
  std::vector<FieldElement> leaves_;
  size_t leaf_count_;
  size_t tree_height_;
  // A single-tree build leaves the levels in HBM (a cuzk_tree_t, shared by copies of this object) and fetches only the
  // root; the host copy the reference class keeps (level 0 = padded leaves ... last = root) is downloaded the first time
  // a caller asks for levels or proofs.
  std::shared_ptr<void> device_tree_;
  // multi-GPU mode (gpus_ > 1): the cuzk_mg_t handle and the sharded tree, shared by copies of this object
  int gpus_ = 1;
  std::shared_ptr<void> mg_, mg_tree_;
  FieldElement root_;
  mutable std::vector<std::vector<FieldElement>> tree_levels_;
  mutable bool levels_on_host_ = true;

  mutable size_t device_proofs_served_ = 0;

  void fetch_levels() const;
  bool proofs_from_device(const std::vector<size_t> &valid_leaves, std::vector<MerkleProof> &out) const;
  bool build_sharded(const std::vector<FieldElement> &leaf_values);
  // flat proofs of valid leaf indices from wherever the levels live in HBM (one device tree, or the shards)
  bool device_flat_proofs(const uint64_t *idx, size_t q, uint64_t *siblings, uint32_t *positions, size_t &levels_out) const;
  bool has_tree() const { return leaf_count_ != 0; }
  FieldElement compute_empty_hash(size_t arity) const;
  void adopt_levels(const std::vector<FieldElement> &leaves, const FieldElement *flat_levels);
};

// what the benchmark helpers below report (wall clock over host-vector calls)
struct CudaMerkleTreeStats {
  double total_time_ms;
  double build_time_ms;
  double proof_generation_time_ms;
  double proof_verification_time_ms;
  size_t trees_per_second;
  size_t proofs_per_second;
  size_t total_trees;
  size_t total_proofs;
  double speedup_vs_cpu;
  size_t leaf_count;
  size_t tree_height;
  size_t arity;
};

// leaves FieldElement(seed + tree * n + i) with the reference's seeds; batch_size is accepted and ignored (the library sizes its own launches)
CudaMerkleTreeStats benchmark_cuda_tree_building(size_t trees, size_t leaves_each, size_t arity = 2, size_t batch_size = 32);
CudaMerkleTreeStats benchmark_cuda_proof_generation(size_t proofs, size_t leaves_each, size_t arity = 2, size_t batch_size = 256);
CudaMerkleTreeStats benchmark_cuda_proof_verification(size_t proofs, size_t leaves_each, size_t arity = 2, size_t batch_size = 512);
// needs the reference's CPU tree at link time; lives in its own source file (merkle_tree_cuda_vs_cpu.cpp)
CudaMerkleTreeStats benchmark_cuda_vs_cpu_merkle(size_t trees, size_t leaves_each, size_t arity = 2, size_t batch_size = 32);

namespace CudaMerkleUtils {
MerkleTreeConfig get_optimal_config_for_gpu(size_t leaf_count);
bool check_cuda_compatibility();
std::vector<std::vector<FieldElement>> generate_batch_test_leaves(size_t num_trees, size_t leaves_per_tree, uint64_t seed = 0);
}  // namespace CudaMerkleUtils

}  // namespace MerkleTreeCUDA
}  // namespace MerkleTree
