// merkle_tree/merkle_tree.hpp -- Merkle configuration and proof types of the cuZK host interface.
//
// Stand-alone counterpart of the reference's src/merkle_tree/merkle_tree.hpp:17-136.  MerkleTreeConfig and
// MerkleProof are the value types the GPU tree exchanges with its callers.  NaryMerkleTree (the reference's CPU tree,
// i.e. the parity oracle) is only forward-declared so that CudaNaryMerkleTree::compare_with_cpu_tree keeps its
// signature; inside the reference tree the reference's own header provides the full class.
#pragma once

#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../common/error_handling.hpp"
#include "../common/namespace_utils.hpp"
#include "../poseidon/poseidon.hpp"

namespace MerkleTree {

using FieldElement = Poseidon::FieldElement;

struct MerkleTreeConfig {
  static constexpr size_t DEFAULT_ARITY = 2;
  static constexpr size_t MIN_ARITY = 2;
  static constexpr size_t MAX_ARITY = 8;
  static constexpr size_t DEFAULT_TREE_HEIGHT = 20;

  size_t arity;
  size_t tree_height;  // carried, never used to size a build (reference: merkle_tree.hpp:24-31)

  explicit MerkleTreeConfig(size_t arity = DEFAULT_ARITY, size_t tree_height = DEFAULT_TREE_HEIGHT)
      : arity(arity), tree_height(tree_height) {
    cuZK::ErrorHandling::validate_range(arity, MIN_ARITY, MAX_ARITY, "arity");  // throws std::invalid_argument
  }
};

// path[l] = the arity-1 siblings at level l (0 = leaf level) in index order with the own slot skipped,
// indices[l] = the own slot at level l, leaf_index = index of the proven leaf
struct MerkleProof {
  std::vector<std::vector<FieldElement>> path;
  std::vector<size_t> indices;
  size_t leaf_index;
  MerkleProof() : leaf_index(0) {}
};

class NaryMerkleTree;  // the reference's CPU tree (the parity oracle): never defined by this product

}  // namespace MerkleTree
