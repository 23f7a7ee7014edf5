// common/namespace_utils.hpp -- USING_* convenience macros (reference: src/common/namespace_utils.hpp:34-54).
// Stand-alone counterpart; the reference tree's own header wins when the host layer is dropped into it.
#pragma once

namespace Poseidon {
struct FieldElement;
class PoseidonHash;
}  // namespace Poseidon
namespace MerkleTree {
class NaryMerkleTree;
struct MerkleTreeConfig;
struct MerkleProof;
}  // namespace MerkleTree

#define USING_CUZK_TYPES()                     \
  using FieldElement = Poseidon::FieldElement; \
  using Hash = Poseidon::PoseidonHash;
#define USING_MERKLE_TYPES()                       \
  using NaryTree = MerkleTree::NaryMerkleTree;     \
  using TreeConfig = MerkleTree::MerkleTreeConfig; \
  using Proof = MerkleTree::MerkleProof;
// USING_FIELD_CONSTANTS / USING_FIELD_OPS of the reference name its CPU arithmetic and have no stand-alone counterpart.
