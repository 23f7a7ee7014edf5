// ref_shim.cpp -- extern "C" doorway into the UNMODIFIED reference CPU implementation
// (TEST INFRASTRUCTURE ONLY; see oracle/Makefile).
//
// This file is ours; it is compiled together with the reference's own sources
//   /root/reference/src/poseidon/{field_arithmetic,poseidon}.cpp
//   /root/reference/src/merkle_tree/merkle_tree.cpp
// where they lie, into oracle/_ref/libcuzk_ref.so.  Nothing from the reference is copied
// into this repository.  The library is used (a) to pin oracle/cuzk_oracle.c, (b) to generate
// tests/golden/, (c) as bench.py's cpu_baseline / --impl reference arm ("kind": "reference").
#include <chrono>
#include <cstring>
#include <optional>
#include <thread>
#include <vector>

#include "merkle_tree.hpp"
#include "poseidon.hpp"

using Poseidon::FieldElement;
namespace FA = Poseidon::FieldArithmetic;

static inline FieldElement ld(const uint64_t *p) { return FieldElement(p[0], p[1], p[2], p[3]); }
static inline void st(uint64_t *p, const FieldElement &f) { std::memcpy(p, f.limbs, 32); }

extern "C" {

// op: 0 add, 1 subtract, 2 multiply, 3 square, 4 power5 (field_arithmetic.hpp:47-61)
void cuzk_ref_batch_fr(int op, const uint64_t *a, const uint64_t *b, uint64_t *r, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    FieldElement x = ld(a + 4 * i), y = b ? ld(b + 4 * i) : FieldElement(), z;
    switch (op) {
    case 0: FA::add(x, y, z); break;
    case 1: FA::subtract(x, y, z); break;
    case 2: FA::multiply(x, y, z); break;
    case 3: FA::square(x, z); break;
    case 4: FA::power5(x, z); break;
    }
    st(r + 4 * i, z);
  }
}

void cuzk_ref_reduce(uint64_t a[4]) {
  FieldElement x = ld(a);
  FA::reduce(x);
  st(a, x);
}

void cuzk_ref_reduce_512(const uint64_t product[8], uint64_t r[4]) {
  FieldElement z;
  FA::reduce_512(product, z);
  st(r, z);
}

void cuzk_ref_round_constants(uint64_t *out) {
  Poseidon::PoseidonConstants::init();
  for (size_t i = 0; i < Poseidon::PoseidonConstants::ROUND_CONSTANTS.size(); ++i)
    st(out + 4 * i, Poseidon::PoseidonConstants::ROUND_CONSTANTS[i]);
}

void cuzk_ref_mds(uint64_t *out) {
  Poseidon::PoseidonConstants::init();
  for (size_t i = 0; i < Poseidon::PoseidonConstants::MDS_MATRIX.size(); ++i)
    st(out + 4 * i, Poseidon::PoseidonConstants::MDS_MATRIX[i]);
}

void cuzk_ref_batch_permutation(uint64_t *states, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    FieldElement s[3] = {ld(states + 12 * i), ld(states + 12 * i + 4), ld(states + 12 * i + 8)};
    Poseidon::PoseidonHash::permutation(s);
    for (int j = 0; j < 3; ++j) st(states + 12 * i + 4 * j, s[j]);
  }
}

void cuzk_ref_batch_hash_single(const uint64_t *in, uint64_t *out, size_t n) {
  for (size_t i = 0; i < n; ++i) st(out + 4 * i, Poseidon::PoseidonHash::hash_single(ld(in + 4 * i)));
}

void cuzk_ref_batch_hash_pairs(const uint64_t *l, const uint64_t *r, uint64_t *out, size_t n) {
  for (size_t i = 0; i < n; ++i)
    st(out + 4 * i, Poseidon::PoseidonHash::hash_pair(ld(l + 4 * i), ld(r + 4 * i)));
}

void cuzk_ref_batch_sponge(const uint64_t *in, size_t width, uint64_t ds, uint64_t *out, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    std::vector<FieldElement> v;
    for (size_t j = 0; j < width; ++j) v.push_back(ld(in + 4 * (i * width + j)));
    st(out + 4 * i, Poseidon::PoseidonHash::sponge(v, FieldElement(ds)));
  }
}

// all-core CPU baseline: hardware threads each hash_pair a disjoint slice
void cuzk_ref_batch_hash_pairs_mt(const uint64_t *l, const uint64_t *r, uint64_t *out, size_t n, int threads) {
  Poseidon::PoseidonConstants::init();
  {  // touch reduce_512's function-local static before going parallel
    uint64_t o[4];
    cuzk_ref_batch_hash_pairs(l, r, o, n ? 1 : 0);
  }
  if (threads < 1) threads = 1;
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) {
    size_t lo = n * t / threads, hi = n * (t + 1) / threads;
    th.emplace_back([=] { cuzk_ref_batch_hash_pairs(l + 4 * lo, r + 4 * lo, out + 4 * lo, hi - lo); });
  }
  for (auto &x : th) x.join();
}

// the reference's own timed loop (poseidon.cpp:195-219); returns hashes per second
double cuzk_ref_benchmark_poseidon_pairs(size_t n) {
  auto s = Poseidon::benchmark_poseidon_pairs(n);
  return 1e9 / s.avg_time_per_hash_ns;
}

void cuzk_ref_empty_hash(size_t arity, uint64_t out[4]) {
  st(out, MerkleTree::NaryMerkleTree::compute_empty_hash(arity));
}

size_t cuzk_ref_tree_height(size_t leaf_count, size_t arity) {
  return MerkleTree::NaryMerkleTree::calculate_tree_height(leaf_count, arity);
}

static std::vector<FieldElement> leaves_of(const uint64_t *leaves, size_t n) {
  std::vector<FieldElement> v;
  v.reserve(n);
  for (size_t i = 0; i < n; ++i) v.push_back(ld(leaves + 4 * i));
  return v;
}

void *cuzk_ref_tree_new(const uint64_t *leaves, size_t n, size_t arity) {
  return new MerkleTree::NaryMerkleTree(leaves_of(leaves, n), MerkleTree::MerkleTreeConfig(arity));
}
void cuzk_ref_tree_free(void *t) { delete static_cast<MerkleTree::NaryMerkleTree *>(t); }
void cuzk_ref_tree_root(void *t, uint64_t root[4]) {
  st(root, static_cast<MerkleTree::NaryMerkleTree *>(t)->get_root_hash());
}
size_t cuzk_ref_tree_get_height(void *t) { return static_cast<MerkleTree::NaryMerkleTree *>(t)->get_tree_height(); }

// proof -> flat (levels x (arity-1)) siblings + positions; returns levels or -1 for nullopt
long cuzk_ref_tree_prove(void *t, size_t leaf_index, uint64_t *siblings_out, uint64_t *positions_out) {
  auto *tree = static_cast<MerkleTree::NaryMerkleTree *>(t);
  auto proof = tree->generate_proof(leaf_index);
  if (!proof) return -1;
  size_t a1 = tree->get_arity() - 1;
  for (size_t l = 0; l < proof->path.size(); ++l) {
    positions_out[l] = proof->indices[l];
    for (size_t s = 0; s < proof->path[l].size(); ++s) st(siblings_out + 4 * (l * a1 + s), proof->path[l][s]);
  }
  return (long)proof->path.size();
}

int cuzk_ref_tree_verify(void *t, const uint64_t leaf[4], const uint64_t *siblings, const uint64_t *positions,
                         size_t levels, const uint64_t root[4]) {
  auto *tree = static_cast<MerkleTree::NaryMerkleTree *>(t);
  size_t a1 = tree->get_arity() - 1;
  MerkleTree::MerkleProof proof;
  for (size_t l = 0; l < levels; ++l) {
    proof.indices.push_back(positions[l]);
    std::vector<FieldElement> sib;
    for (size_t s = 0; s < a1; ++s) sib.push_back(ld(siblings + 4 * (l * a1 + s)));
    proof.path.push_back(sib);
  }
  return tree->verify_proof(proof, ld(leaf), ld(root)) ? 1 : 0;
}

// generate_test_leaves : merkle_tree.cpp:448-460
void cuzk_ref_generate_test_leaves(size_t count, uint64_t seed, uint64_t *out) {
  auto v = MerkleTree::MerkleUtils::generate_test_leaves(count, seed);
  for (size_t i = 0; i < count; ++i) st(out + 4 * i, v[i]);
}

// build time of the reference CPU tree in milliseconds (for the cpu_baseline leg)
double cuzk_ref_tree_build_ms(const uint64_t *leaves, size_t n, size_t arity, uint64_t root[4]) {
  auto v = leaves_of(leaves, n);
  auto t0 = std::chrono::high_resolution_clock::now();
  MerkleTree::NaryMerkleTree tree(v, MerkleTree::MerkleTreeConfig(arity));
  auto t1 = std::chrono::high_resolution_clock::now();
  st(root, tree.get_root_hash());
  return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

}  // extern "C"
