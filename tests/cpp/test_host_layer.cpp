// test_host_layer.cpp -- parity tests of the C++ host layer (the reference's CUDA-side class names over the C ABI)
// against the plain-C oracle.  Modelled on the reference's own GPU suites (src/poseidon/test/test_field_arithmetic_cuda.cpp,
// test_poseidon_cuda.cpp, src/merkle_tree/test/test_merkle_tree_cuda.cpp): every GPU result must equal the CPU result
// bit for bit.  Here the CPU side is oracle/libcuzk_oracle.so, so the suite runs on the GPU box without the reference.
#include <gtest/gtest.h>

#include <random>
#include <thread>
#include <vector>

#include "cuzk_b200.h"
#include "../../cuzk_b200/host/src/merkle_tree/merkle_tree_cuda.cuh"
#include "../../cuzk_b200/host/src/poseidon/cuda/poseidon_cuda.cuh"
#include "../../cuzk_b200/host/src/poseidon/cuda/poseidon_cuda_benchmarks.hpp"
#include "../../cuzk_b200/host/src/poseidon/cuda/poseidon_cuda_optimized.cuh"

extern "C" {  // oracle/cuzk_oracle.c
void cuzk_oracle_batch_fr(int op, const uint64_t *a, const uint64_t *b, uint64_t *r, size_t n);
void cuzk_oracle_permutation(uint64_t *state);
void cuzk_oracle_hash_single(const uint64_t in[4], uint64_t out[4]);
void cuzk_oracle_hash_pair(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]);
void cuzk_oracle_batch_sponge(const uint64_t *in, size_t width, uint64_t ds, uint64_t *out, size_t n);
void cuzk_oracle_empty_hash(size_t arity, uint64_t out[4]);
size_t cuzk_oracle_tree_height_float(size_t leaf_count, size_t arity);
size_t cuzk_oracle_total_nodes(size_t n, size_t arity);
size_t cuzk_oracle_padded_size(size_t n, size_t arity);
size_t cuzk_oracle_num_levels(size_t n, size_t arity);
void cuzk_oracle_merkle_build(const uint64_t *leaves, size_t n, size_t arity, uint64_t *levels_out);
}

using Poseidon::CudaFieldElement;
using Poseidon::FieldElement;
using namespace Poseidon::CudaFieldOps;
using namespace Poseidon::PoseidonCUDA;
using namespace MerkleTree;
using namespace MerkleTree::MerkleTreeCUDA;

namespace {

const uint64_t *raw(const std::vector<FieldElement> &v) { return reinterpret_cast<const uint64_t *>(v.data()); }
uint64_t *raw(std::vector<FieldElement> &v) { return reinterpret_cast<uint64_t *>(v.data()); }

std::vector<FieldElement> random_elements(size_t n, uint64_t seed, bool canonical) {
  std::mt19937_64 gen(seed);
  std::vector<FieldElement> v(n);
  for (auto &e : v) e = FieldElement(gen(), gen(), gen(), canonical ? gen() >> 4 : gen());
  return v;
}
std::vector<FieldElement> edge_elements() {
  const uint64_t M = ~0ull;
  return {FieldElement(0), FieldElement(1), FieldElement(2), FieldElement(M, M, M, M), FieldElement(M, M, 0, 0),
          FieldElement(0x43e1f593f0000000ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull),   // p - 1
          FieldElement(0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull),   // p
          FieldElement(0x43e1f593f0000002ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull),   // p + 1
          FieldElement(0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full),   // k
          FieldElement(1, M, 0, 0), FieldElement(0, 5, 0, 1), FieldElement(0, 0, 0, 1ull << 63)};
}
std::vector<FieldElement> u64_leaves(size_t n, uint32_t seed) {  // the reference GPU suite's leaves (test_merkle_tree_cuda.cpp:31-44)
  std::mt19937 gen(seed);
  std::uniform_int_distribution<uint64_t> dist(1, UINT64_MAX);
  std::vector<FieldElement> v;
  v.reserve(n);
  for (size_t i = 0; i < n; ++i) v.emplace_back(dist(gen));
  return v;
}
std::vector<std::vector<FieldElement>> oracle_levels(const std::vector<FieldElement> &leaves, size_t arity) {
  std::vector<FieldElement> flat(cuzk_oracle_total_nodes(leaves.size(), arity));
  cuzk_oracle_merkle_build(raw(leaves), leaves.size(), arity, raw(flat));
  std::vector<std::vector<FieldElement>> out;
  size_t w = cuzk_oracle_padded_size(leaves.size(), arity), at = 0;
  for (size_t l = 0; l < cuzk_oracle_num_levels(leaves.size(), arity); ++l) {
    out.emplace_back(flat.begin() + at, flat.begin() + at + w);
    at += w;
    w /= arity;
  }
  return out;
}

}  // namespace

// ------------------------------------------------------------------------------------------ field arithmetic
class HostFieldArithmetic : public ::testing::Test {
protected:
  void SetUp() override { ASSERT_TRUE(CudaFieldArithmetic::initialize()) << "no usable GPU: cuzk_b200 has no CPU fallback"; }
  void TearDown() override { CudaFieldArithmetic::cleanup(); }
};

TEST_F(HostFieldArithmetic, BatchOpsEqualOracle) {
  for (bool canonical : {true, false}) {
    auto a = random_elements(1500, 1, canonical), b = random_elements(1500, 2, canonical);
    const auto edges = edge_elements();
    for (const auto &x : edges)
      for (const auto &y : edges) {
        a.push_back(x);
        b.push_back(y);
      }
    std::vector<FieldElement> got, want(a.size());
    struct Op { int code; bool binary; const char *name; };
    for (Op op : {Op{0, true, "add"}, Op{1, true, "subtract"}, Op{2, true, "multiply"}, Op{3, false, "square"}, Op{4, false, "power5"}}) {
      SCOPED_TRACE(op.name);
      bool ok = false;
      switch (op.code) {
        case 0: ok = CudaFieldArithmetic::batch_add(a, b, got); break;
        case 1: ok = CudaFieldArithmetic::batch_subtract(a, b, got); break;
        case 2: ok = CudaFieldArithmetic::batch_multiply(a, b, got); break;
        case 3: ok = CudaFieldArithmetic::batch_square(a, got); break;
        case 4: ok = CudaFieldArithmetic::batch_power5(a, got); break;
      }
      ASSERT_TRUE(ok);
      ASSERT_EQ(got.size(), a.size());
      cuzk_oracle_batch_fr(op.code, raw(a), raw(b), raw(want), a.size());
      size_t bad = 0;
      for (size_t i = 0; i < a.size(); ++i) bad += got[i] != want[i];
      EXPECT_EQ(bad, 0u);
    }
  }
}

TEST_F(HostFieldArithmetic, EmptyMismatchedAndSingle) {
  std::vector<FieldElement> none, out{FieldElement(7)}, three(3), two(2);
  EXPECT_TRUE(CudaFieldArithmetic::batch_add(none, none, out));
  EXPECT_TRUE(out.empty());
  EXPECT_FALSE(CudaFieldArithmetic::batch_multiply(three, two, out));
  FieldElement r, want;
  const FieldElement a(5), b(3);
  ASSERT_TRUE(CudaFieldArithmetic::gpu_multiply(a, b, r));
  EXPECT_EQ(r, FieldElement(15));
  ASSERT_TRUE(CudaFieldArithmetic::gpu_power5(FieldElement(2), r));
  EXPECT_EQ(r, FieldElement(32));
  EXPECT_GT(CudaFieldArithmetic::get_device_count(), 0);
  EXPECT_GT(CudaFieldArithmetic::get_optimal_block_size(), 0u);
}

TEST_F(HostFieldArithmetic, ResultMayAliasInput) {
  auto a = random_elements(257, 9, true), b = random_elements(257, 10, true);
  std::vector<FieldElement> want(a.size());
  cuzk_oracle_batch_fr(2, raw(a), raw(b), raw(want), a.size());
  ASSERT_TRUE(CudaFieldArithmetic::batch_multiply(a, b, a));
  EXPECT_TRUE(a == want);
}

// ------------------------------------------------------------------------------------------ Poseidon
class HostPoseidon : public ::testing::Test {
protected:
  void SetUp() override {
    hasher = std::make_unique<CudaPoseidonHash>();
    ASSERT_TRUE(hasher->is_initialized()) << "no usable GPU: cuzk_b200 has no CPU fallback";
  }
  void TearDown() override { hasher.reset(); }
  std::unique_ptr<CudaPoseidonHash> hasher;
};

TEST_F(HostPoseidon, SingleAndPairHashesEqualOracle) {
  auto l = random_elements(3000, 3, true), r = random_elements(3000, 4, true);
  for (const auto &e : edge_elements()) {  // includes non-canonical inputs: the sponge reduces them on absorption
    l.push_back(e);
    r.push_back(FieldElement(e.limbs[3], e.limbs[2], e.limbs[1], e.limbs[0]));
  }
  std::vector<FieldElement> got;
  ASSERT_TRUE(hasher->batch_hash_single(l, got));
  ASSERT_EQ(got.size(), l.size());
  size_t bad = 0;
  for (size_t i = 0; i < l.size(); ++i) {
    FieldElement want;
    cuzk_oracle_hash_single(l[i].limbs, want.limbs);
    bad += got[i] != want;
  }
  EXPECT_EQ(bad, 0u);
  ASSERT_TRUE(hasher->batch_hash_pairs(l, r, got));
  bad = 0;
  for (size_t i = 0; i < l.size(); ++i) {
    FieldElement want;
    cuzk_oracle_hash_pair(l[i].limbs, r[i].limbs, want.limbs);
    bad += got[i] != want;
  }
  EXPECT_EQ(bad, 0u);
}

TEST_F(HostPoseidon, LargeBatchCrossesPipelineChunks) {
  const size_t n = 300000;  // > 2 chunks of the double-buffered host path
  auto l = random_elements(n, 5, true), r = random_elements(n, 6, true);
  std::vector<FieldElement> got;
  ASSERT_TRUE(hasher->batch_hash_pairs(l, r, got));
  ASSERT_EQ(got.size(), n);
  size_t bad = 0;
  for (size_t i = 0; i < n; i += 997) {
    FieldElement want;
    cuzk_oracle_hash_pair(l[i].limbs, r[i].limbs, want.limbs);
    bad += got[i] != want;
  }
  for (size_t i : {size_t(0), size_t(113663), size_t(113664), size_t(227327), size_t(227328), n - 1}) {
    FieldElement want;
    cuzk_oracle_hash_pair(l[i].limbs, r[i].limbs, want.limbs);
    bad += got[i] != want;
  }
  EXPECT_EQ(bad, 0u);
}

TEST_F(HostPoseidon, PermutationAndSpongeEqualOracle) {
  auto flat = random_elements(3 * 500, 7, false);  // caller-supplied states may be non-canonical
  std::vector<std::array<CudaFieldElement, 3>> states(500);
  for (size_t i = 0; i < 500; ++i)
    for (int j = 0; j < 3; ++j) states[i][j] = CudaFieldElement(flat[3 * i + j]);
  ASSERT_TRUE(hasher->batch_permutation(states));
  size_t bad = 0;
  for (size_t i = 0; i < 500; ++i) {
    uint64_t st[12];
    for (int j = 0; j < 3; ++j) std::memcpy(st + 4 * j, flat[3 * i + j].limbs, 32);
    cuzk_oracle_permutation(st);
    for (int j = 0; j < 3; ++j) bad += std::memcmp(st + 4 * j, states[i][j].limbs, 32) != 0;
  }
  EXPECT_EQ(bad, 0u);
  for (size_t width : {1, 2, 3, 5, 8}) {
    auto in = random_elements(width * 200, 8 + width, true);
    std::vector<FieldElement> got, want(200);
    ASSERT_TRUE(hasher->batch_sponge(in, width, 3, got));
    cuzk_oracle_batch_sponge(raw(in), width, 3, raw(want), 200);
    EXPECT_TRUE(got == want) << "width " << width;
  }
}

TEST_F(HostPoseidon, EdgeCasesAndBothClassNames) {
  std::vector<FieldElement> none, out{FieldElement(1)}, three(3), two(2);
  EXPECT_TRUE(hasher->batch_hash_single(none, out));
  EXPECT_TRUE(out.empty());
  EXPECT_FALSE(hasher->batch_hash_pairs(three, two, out));
  EXPECT_GT(hasher->get_optimal_batch_size(), 0u);
  EXPECT_GT(hasher->get_max_batch_size(), hasher->get_optimal_batch_size());
  // outputs may be one of the input vectors (written in place, chunk by chunk)
  auto a = random_elements(250000, 31, true), b = random_elements(250000, 32, true);
  std::vector<FieldElement> want_pairs, want_single;
  ASSERT_TRUE(hasher->batch_hash_pairs(a, b, want_pairs));
  ASSERT_TRUE(hasher->batch_hash_single(b, want_single));
  ASSERT_TRUE(hasher->batch_hash_pairs(a, b, a));
  EXPECT_TRUE(a == want_pairs);
  ASSERT_TRUE(hasher->batch_hash_single(b, b));
  EXPECT_TRUE(b == want_single);
  Poseidon::PoseidonCUDAOptimized::CudaPoseidonHashOptimized optimized;
  ASSERT_TRUE(optimized.is_initialized());
  EXPECT_TRUE(verify_cuda_implementations_match(*hasher, optimized, "CUDA Original", "CUDA Optimized", 100));
  const CudaPoseidonStats s = benchmark_cuda_poseidon_pairs(*hasher, 20000, 4096);
  EXPECT_EQ(s.total_hashes, 20000u);
  EXPECT_GT(s.hashes_per_second, 0u);
}

TEST(HostLifecycle, DroppingOneUserKeepsTheOthersAlive) {
  auto a = std::make_unique<CudaPoseidonHash>();
  ASSERT_TRUE(a->is_initialized());
  {
    CudaPoseidonHash b;
    ASSERT_TRUE(b.is_initialized());
    ASSERT_TRUE(CudaFieldArithmetic::initialize());
    CudaFieldArithmetic::cleanup();  // the reference resets the device here and kills `a` and `b`
  }
  std::vector<FieldElement> in{FieldElement(42)}, out;
  ASSERT_TRUE(a->batch_hash_single(in, out));
  FieldElement want;
  cuzk_oracle_hash_single(in[0].limbs, want.limbs);
  EXPECT_EQ(out[0], want);
  a.reset();
  CudaFieldArithmetic::cleanup();  // extra cleanup is harmless
  CudaPoseidonHash c;              // and the library comes back up
  EXPECT_TRUE(c.is_initialized());
}

TEST(HostLifecycle, ConcurrentCallersGetTheirOwnResults) {
  // the reference is single-threaded; here host-buffer calls serialise inside the library and must not mix up buffers
  CudaPoseidonHash shared;
  ASSERT_TRUE(shared.is_initialized());
  const int nthreads = 4;
  std::vector<std::vector<FieldElement>> in(nthreads), want(nthreads);
  for (int t = 0; t < nthreads; ++t) {
    in[t] = random_elements(40000 + 1000 * t, 900 + t, true);
    want[t].resize(in[t].size());
    for (size_t i = 0; i < in[t].size(); i += 97) cuzk_oracle_hash_single(in[t][i].limbs, want[t][i].limbs);
  }
  std::vector<int> bad(nthreads, 0);
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; ++t)
    pool.emplace_back([&, t] {
      CudaPoseidonHash mine;   // a second reference on the library, taken and dropped concurrently
      for (int rep = 0; rep < 5; ++rep) {
        std::vector<FieldElement> out;
        IPoseidonCudaHash &h = (rep & 1) ? static_cast<IPoseidonCudaHash &>(mine) : static_cast<IPoseidonCudaHash &>(shared);
        if (!h.batch_hash_single(in[t], out) || out.size() != in[t].size()) { bad[t] += 1000; continue; }
        for (size_t i = 0; i < out.size(); i += 97) bad[t] += out[i] != want[t][i];
        std::vector<FieldElement> leaves(in[t].begin(), in[t].begin() + 500);
        CudaNaryMerkleTree tree(leaves, MerkleTreeConfig(2 + t));
        auto p = tree.generate_proof(17);
        bad[t] += !(p && tree.verify_proof(*p, leaves[17]));
      }
    });
  for (auto &th : pool) th.join();
  for (int t = 0; t < nthreads; ++t) EXPECT_EQ(bad[t], 0) << "thread " << t;
}

// ------------------------------------------------------------------------------------------ Merkle tree
class HostMerkle : public ::testing::Test {
protected:
  void SetUp() override { ASSERT_TRUE(CudaNaryMerkleTree::initialize_cuda()) << "no usable GPU: cuzk_b200 has no CPU fallback"; }
  void TearDown() override { CudaNaryMerkleTree::cleanup_cuda(); }
};

TEST_F(HostMerkle, ConfigValidation) {
  EXPECT_THROW(MerkleTreeConfig(1), std::invalid_argument);
  EXPECT_THROW(MerkleTreeConfig(9), std::invalid_argument);
  EXPECT_TRUE(CudaMerkleUtils::check_cuda_compatibility());
  EXPECT_GT(CudaNaryMerkleTree::get_max_batch_size(), CudaNaryMerkleTree::get_optimal_batch_size());
}

TEST_F(HostMerkle, EveryLevelEqualsOracle) {
  for (size_t arity = 2; arity <= 8; ++arity)
    for (size_t n : {1, 2, 3, 4, 15, 16, 17, 64, 100, 256, 1000}) {
      SCOPED_TRACE("arity " + std::to_string(arity) + ", leaves " + std::to_string(n));
      const auto leaves = u64_leaves(n, 42);
      CudaNaryMerkleTree tree((MerkleTreeConfig(arity)));
      ASSERT_TRUE(tree.build_tree(leaves));
      const auto want = oracle_levels(leaves, arity);
      ASSERT_EQ(tree.get_tree_levels().size(), want.size());
      for (size_t l = 0; l < want.size(); ++l) EXPECT_TRUE(tree.get_tree_levels()[l] == want[l]) << "level " << l;
      EXPECT_EQ(tree.get_root_hash(), want.back()[0]);
      EXPECT_EQ(tree.get_leaf_count(), n);
      EXPECT_EQ(tree.get_tree_height(), cuzk_oracle_tree_height_float(n, arity));
    }
}

TEST_F(HostMerkle, FloatHeightHazardDoesNotAddALevel) {
  // 8^3 = 512 and 5^3 = 125 leaves: the float formula over-counts the height at exact powers; the level
  // arrays must still come from the integer loop and the root must equal the oracle's
  for (auto [arity, n] : {std::pair<size_t, size_t>{5, 125}, {6, 216}, {8, 512}}) {
    const auto leaves = u64_leaves(n, 7);
    CudaNaryMerkleTree tree(leaves, MerkleTreeConfig(arity));
    const auto want = oracle_levels(leaves, arity);
    EXPECT_EQ(tree.get_tree_levels().size(), want.size());
    EXPECT_EQ(tree.get_root_hash(), want.back()[0]);
    EXPECT_EQ(tree.get_tree_height(), cuzk_oracle_tree_height_float(n, arity));
  }
}

TEST_F(HostMerkle, EmptyAndSingleLeafTrees) {
  for (size_t arity : {2, 3, 8}) {
    CudaNaryMerkleTree tree((MerkleTreeConfig(arity)));
    ASSERT_TRUE(tree.build_tree({}));
    FieldElement e;
    cuzk_oracle_empty_hash(arity, e.limbs);
    EXPECT_EQ(tree.get_root_hash(), e);
    EXPECT_EQ(tree.get_leaf_count(), 0u);
    EXPECT_EQ(tree.get_tree_height(), 0u);
    EXPECT_FALSE(tree.generate_proof(0).has_value());
    const std::vector<FieldElement> one{FieldElement(0xc151df7d6ee5e2d6ull)};
    ASSERT_TRUE(tree.build_tree(one));
    EXPECT_EQ(tree.get_root_hash(), one[0]);
    auto p = tree.generate_proof(0);
    ASSERT_TRUE(p.has_value());
    EXPECT_EQ(p->path.size(), 0u);
    EXPECT_TRUE(tree.verify_proof(*p, one[0]));
  }
}

TEST_F(HostMerkle, ProofsVerifyAndCorruptionIsCaught) {
  for (size_t arity : {2, 4, 8}) {
    SCOPED_TRACE("arity " + std::to_string(arity));
    const size_t n = 150;
    const auto leaves = u64_leaves(n, 12345);
    CudaNaryMerkleTree tree(leaves, MerkleTreeConfig(arity));
    const auto levels = oracle_levels(leaves, arity);
    std::vector<size_t> idx(n);
    for (size_t i = 0; i < n; ++i) idx[i] = i;
    idx.push_back(n + 5);  // out of range: skipped silently
    auto proofs = tree.generate_batch_proofs(idx);
    ASSERT_EQ(proofs.size(), n);
    for (size_t i = 0; i < n; ++i) {  // layout: path[l] = siblings in index order without the own slot, indices[l] = own slot
      size_t at = i;
      ASSERT_EQ(proofs[i].path.size(), levels.size() - 1);
      for (size_t l = 0; l + 1 < levels.size(); ++l) {
        EXPECT_EQ(proofs[i].indices[l], at % arity);
        size_t s = 0;
        for (size_t c = 0; c < arity; ++c) {
          if (c == at % arity) continue;
          EXPECT_EQ(proofs[i].path[l][s++], levels[l][at - at % arity + c]);
        }
        at /= arity;
      }
    }
    EXPECT_TRUE(tree.verify_batch_proofs(proofs, leaves));
    // a small batch (the reference would take its CPU branch below 32 proofs) gives the same answer on the GPU
    EXPECT_TRUE(tree.verify_batch_proofs({proofs[3], proofs[77]}, {leaves[3], leaves[77]}));
    EXPECT_FALSE(tree.verify_batch_proofs({}, {}));               // CUDA class: empty batch -> false
    EXPECT_FALSE(tree.verify_batch_proofs(proofs, {leaves[0]}));  // size mismatch -> false
    // corruptions: wrong leaf, flipped sibling bit, wrong slot, slot out of range, short path
    std::vector<MerkleProof> bad(proofs.begin(), proofs.begin() + 6);
    std::vector<FieldElement> vals(leaves.begin(), leaves.begin() + 6);
    vals[1] = FieldElement(vals[1].limbs[0] ^ 1);
    bad[2].path[1][0].limbs[2] ^= 1ull << 17;
    bad[3].indices[0] = (bad[3].indices[0] + 1) % arity;
    bad[4].indices[1] = arity;
    bad[5].path.pop_back();
    bad[5].indices.pop_back();
    std::vector<uint8_t> res;
    ASSERT_TRUE(tree.verify_batch_proofs_each(bad, vals, res));
    EXPECT_EQ(res, (std::vector<uint8_t>{1, 0, 0, 0, 0, 0}));
    EXPECT_FALSE(tree.verify_batch_proofs(bad, vals));
    EXPECT_FALSE(tree.verify_proof(proofs[9], leaves[10]));
    EXPECT_TRUE(tree.verify_proof(proofs[9], leaves[9]));
    // flat batches carry the same siblings and slots as the MerkleProof objects, and verify the same way
    FlatProofBatch flat;
    std::vector<size_t> all(n);
    for (size_t i = 0; i < n; ++i) all[i] = i;
    ASSERT_TRUE(tree.generate_flat_proofs(all, flat));
    ASSERT_EQ(flat.levels, levels.size() - 1);
    size_t mismatches = 0;
    for (size_t i = 0; i < n; ++i)
      for (size_t l = 0; l < flat.levels; ++l) {
        mismatches += flat.positions[i * flat.levels + l] != proofs[i].indices[l];
        for (size_t s2 = 0; s2 + 1 < arity; ++s2)
          mismatches += flat.siblings[(i * flat.levels + l) * (arity - 1) + s2] != proofs[i].path[l][s2];
      }
    EXPECT_EQ(mismatches, 0u);
    std::vector<uint8_t> verdicts;
    std::vector<FieldElement> tampered = leaves;
    tampered[5] = FieldElement(tampered[5].limbs[0] + 1);
    ASSERT_TRUE(tree.verify_flat_proofs(flat, tampered, verdicts));
    for (size_t i = 0; i < n; ++i) EXPECT_EQ(verdicts[i], i == 5 ? 0 : 1);
    EXPECT_FALSE(tree.generate_flat_proofs({n}, flat));
  }
}

TEST_F(HostMerkle, MalformedSiblingListsAreRejectedLikeTheCpuVerifier) {
  // NaryMerkleTree::verify_proof (merkle_tree.cpp:228-230) rejects a level that does not carry exactly arity - 1 siblings:
  // neither junk appended to a valid level nor a trailing sibling that happens to be padding may be dropped or filled in
  const size_t arity = 4, n = 13;   // 13 leaves of 16: the last group of level 0 ends in padding
  const auto leaves = u64_leaves(n, 99);
  CudaNaryMerkleTree tree(leaves, MerkleTreeConfig(arity));
  auto proofs = tree.generate_batch_proofs({0, 12, 5});
  ASSERT_EQ(proofs.size(), 3u);
  std::vector<FieldElement> vals{leaves[0], leaves[12], leaves[5]};
  ASSERT_TRUE(tree.verify_batch_proofs(proofs, vals));
  proofs[0].path[0].push_back(FieldElement(7));   // extra junk sibling
  proofs[1].path[0].pop_back();                   // missing trailing sibling (a padding constant)
  std::vector<uint8_t> res;
  ASSERT_TRUE(tree.verify_batch_proofs_each(proofs, vals, res));
  EXPECT_EQ(res, (std::vector<uint8_t>{0, 0, 1}));
}

TEST_F(HostMerkle, ShardedOverSeveralGpusEqualsTheSingleGpuTree) {
  // opt-in multi-GPU constructor (cuzk_mg_*): with one GPU in the box it still exercises the sharded code path (one shard)
  const int gpus = std::min(cuzk_device_count(), 8);
  for (size_t arity : {2, 8}) {
    const size_t n = arity == 2 ? 3000 : 40000;
    const auto leaves = u64_leaves(n, 4242);
    CudaNaryMerkleTree single(leaves, MerkleTreeConfig(arity));
    CudaNaryMerkleTree sharded(leaves, MerkleTreeConfig(arity), gpus);
    EXPECT_EQ(sharded.get_gpu_count(), std::max(gpus, 1));
    ASSERT_EQ(sharded.get_leaf_count(), n);
    EXPECT_EQ(sharded.get_root_hash(), single.get_root_hash());
    std::vector<size_t> idx;
    for (size_t i = 0; i < n; i += 37) idx.push_back(i);
    idx.push_back(n - 1);
    const auto a = single.generate_batch_proofs(idx), b = sharded.generate_batch_proofs(idx);
    ASSERT_EQ(a.size(), b.size());
    size_t differ = 0;
    std::vector<FieldElement> vals;
    for (size_t q = 0; q < a.size(); ++q) {
      differ += a[q].path != b[q].path || a[q].indices != b[q].indices;
      vals.push_back(leaves[idx[q]]);
    }
    EXPECT_EQ(differ, 0u);
    EXPECT_TRUE(sharded.verify_batch_proofs(b, vals));
    FlatProofBatch flat;
    ASSERT_TRUE(sharded.generate_flat_proofs(idx, flat));
    std::vector<uint8_t> verdicts;
    vals[1] = FieldElement(vals[1].limbs[0] ^ 2);
    ASSERT_TRUE(sharded.verify_flat_proofs(flat, vals, verdicts));
    EXPECT_EQ(verdicts[0], 1);
    EXPECT_EQ(verdicts[1], 0);
  }
}

TEST_F(HostMerkle, BatchTreesAndBenchmarkHelpers) {
  const auto batch = CudaMerkleUtils::generate_batch_test_leaves(10, 32, 12345);
  std::vector<CudaNaryMerkleTree> trees;
  ASSERT_TRUE(CudaNaryMerkleTree::build_batch_trees(batch, trees, MerkleTreeConfig(4)));
  ASSERT_EQ(trees.size(), 10u);
  for (size_t t = 0; t < 10; ++t) EXPECT_EQ(trees[t].get_root_hash(), oracle_levels(batch[t], 4).back()[0]);
  // ragged batch (falls back to one build per tree) and a larger uniform forest
  std::vector<std::vector<FieldElement>> ragged{u64_leaves(5, 1), u64_leaves(64, 2), {}, u64_leaves(9, 3)};
  ASSERT_TRUE(CudaNaryMerkleTree::build_batch_trees(ragged, trees, MerkleTreeConfig(2)));
  ASSERT_EQ(trees.size(), 4u);
  EXPECT_EQ(trees[1].get_root_hash(), oracle_levels(ragged[1], 2).back()[0]);
  EXPECT_EQ(trees[2].get_leaf_count(), 0u);
  const auto forest = CudaMerkleUtils::generate_batch_test_leaves(50, 2048, 777);
  ASSERT_TRUE(CudaNaryMerkleTree::build_batch_trees(forest, trees, MerkleTreeConfig(8)));
  for (size_t t : {size_t(0), size_t(31), size_t(49)}) {
    const auto want = oracle_levels(forest[t], 8);
    ASSERT_EQ(trees[t].get_tree_levels().size(), want.size());
    for (size_t l = 0; l < want.size(); ++l) EXPECT_TRUE(trees[t].get_tree_levels()[l] == want[l]);
  }
  const CudaMerkleTreeStats b = benchmark_cuda_tree_building(4, 256, 4);
  EXPECT_GT(b.build_time_ms, 0.0);
  const CudaMerkleTreeStats v = benchmark_cuda_proof_verification(200, 256, 4);
  EXPECT_GT(v.proof_verification_time_ms, 0.0);
  EXPECT_EQ(CudaMerkleUtils::get_optimal_config_for_gpu(500).arity, 2u);
  EXPECT_EQ(CudaMerkleUtils::get_optimal_config_for_gpu(200000).arity, 8u);
}
