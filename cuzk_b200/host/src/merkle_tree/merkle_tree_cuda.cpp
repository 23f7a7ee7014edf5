// merkle_tree_cuda.cpp -- CudaNaryMerkleTree over the C ABI (replaces src/merkle_tree/merkle_tree_cuda.cu:120-740; its two
// kernels, :45-118, are in libcuzk_b200.so).  Host C++ only: marshals std::vector data to the flat layouts the C ABI takes.
#include "merkle_tree_cuda.cuh"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <iostream>
#include <map>
#include <mutex>
#include <random>

#include "cuzk_b200.h"

namespace MerkleTree {
namespace MerkleTreeCUDA {

namespace {

std::mutex g_mu;
int g_init_refs = 0;  // references this translation unit holds on the library through initialize_cuda()

const uint64_t *raw(const std::vector<FieldElement> &v) { return reinterpret_cast<const uint64_t *>(v.data()); }
uint64_t *raw(std::vector<FieldElement> &v) { return reinterpret_cast<uint64_t *>(v.data()); }

bool ensure_library() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_init_refs > 0) return true;
  if (cuzk_init(0) != CUZK_OK) {
    std::cerr << "Failed to initialize CUDA Poseidon: " << cuzk_last_error() << std::endl;
    return false;
  }
  g_init_refs = 1;
  return true;
}

}  // namespace

CudaNaryMerkleTree::CudaNaryMerkleTree(const MerkleTreeConfig &config) : config_(config), leaf_count_(0), tree_height_(0) {
  ensure_library();
}

CudaNaryMerkleTree::CudaNaryMerkleTree(const std::vector<FieldElement> &leaves, const MerkleTreeConfig &config)
    : config_(config), leaf_count_(0), tree_height_(0) {
  ensure_library();
  build_tree(leaves);
}

CudaNaryMerkleTree::CudaNaryMerkleTree(const std::vector<FieldElement> &leaves, const MerkleTreeConfig &config, int gpus)
    : config_(config), leaf_count_(0), tree_height_(0), gpus_(gpus < 1 ? 1 : gpus) {
  ensure_library();
  build_tree(leaves);
}

CudaNaryMerkleTree::~CudaNaryMerkleTree() = default;

// multi-GPU mode: one cuzk_mg_t per object family (created on first use), the tree sharded over its devices
bool CudaNaryMerkleTree::build_sharded(const std::vector<FieldElement> &leaves) {
  if (!mg_) {
    cuzk_mg_t *mg = nullptr;
    if (cuzk_mg_init_local(gpus_, nullptr, &mg) != CUZK_OK) {
      std::cerr << "CudaNaryMerkleTree: multi-GPU initialisation failed: " << cuzk_last_error() << std::endl;
      return false;
    }
    mg_ = std::shared_ptr<void>(mg, [](void *h) { cuzk_mg_free(static_cast<cuzk_mg_t *>(h)); });
  }
  cuzk_mg_tree_t *t = nullptr;
  const uint64_t *whole[1] = {raw(leaves)};
  if (cuzk_mg_tree_build(static_cast<cuzk_mg_t *>(mg_.get()), whole, leaves.size(), (unsigned)config_.arity, CUZK_MEM_HOST, &t) != CUZK_OK ||
      cuzk_mg_tree_root(t, root_.limbs) != CUZK_OK) {
    std::cerr << "CudaNaryMerkleTree::build_tree (sharded): " << cuzk_last_error() << std::endl;
    if (t) cuzk_mg_tree_free(t);
    return false;
  }
  mg_tree_ = std::shared_ptr<void>(t, [](void *h) { cuzk_mg_tree_free(static_cast<cuzk_mg_tree_t *>(h)); });
  return true;
}

bool CudaNaryMerkleTree::device_flat_proofs(const uint64_t *idx, size_t q, uint64_t *siblings, uint32_t *positions, size_t &levels_out) const {
  if (mg_tree_) {
    const cuzk_mg_tree_t *t = static_cast<const cuzk_mg_tree_t *>(mg_tree_.get());
    levels_out = cuzk_mg_tree_num_levels(t) - 1;
    return !q || !levels_out || cuzk_mg_tree_prove_batch(t, idx, q, siblings, positions) == CUZK_OK;
  }
  const cuzk_tree_t *t = static_cast<const cuzk_tree_t *>(device_tree_.get());
  levels_out = cuzk_tree_num_levels(t) - 1;
  return !q || !levels_out || cuzk_tree_prove_batch(t, idx, q, siblings, positions, CUZK_MEM_HOST, nullptr) == CUZK_OK;
}

bool CudaNaryMerkleTree::build_tree(const std::vector<FieldElement> &leaves) {
  tree_levels_.clear();
  device_tree_.reset();
  mg_tree_.reset();
  levels_on_host_ = true;
  if (leaves.empty()) {  // an empty input clears the tree (merkle_tree_cuda.cu:142-148)
    leaf_count_ = 0;
    tree_height_ = 0;
    leaves_.clear();
    return true;
  }
  if (!ensure_library()) return false;
  const unsigned arity = (unsigned)config_.arity;
  const size_t n = leaves.size();
  if (gpus_ > 1) {   // sharded over several GPUs: the levels stay on their devices, the host keeps the root
    if (!build_sharded(leaves)) {
      leaf_count_ = 0;
      tree_height_ = 0;
      leaves_.clear();
      return false;
    }
    if (&leaves != &leaves_) leaves_ = leaves;
    leaf_count_ = n;
    tree_height_ = cuzk_merkle_tree_height(n, arity);
    return true;
  }
  cuzk_tree_t *handle = nullptr;
  if (cuzk_tree_build(raw(leaves), n, arity, CUZK_MEM_HOST, nullptr, &handle) != CUZK_OK ||
      cuzk_tree_root(handle, root_.limbs, CUZK_MEM_HOST, nullptr) != CUZK_OK) {
    std::cerr << "CudaNaryMerkleTree::build_tree: " << cuzk_last_error() << std::endl;
    if (handle) cuzk_tree_free(handle);
    leaf_count_ = 0;
    tree_height_ = 0;
    leaves_.clear();
    return false;
  }
  device_tree_ = std::shared_ptr<void>(handle, [](void *h) { cuzk_tree_free(static_cast<cuzk_tree_t *>(h)); });
  if (&leaves != &leaves_) leaves_ = leaves;
  leaf_count_ = n;
  tree_height_ = cuzk_merkle_tree_height(n, arity);  // the reference's float formula: a getter value only
  levels_on_host_ = false;
  return true;
}

// downloads the level arrays the first time they are needed on the host, one copy per level straight into its vector
void CudaNaryMerkleTree::fetch_levels() const {
  if (levels_on_host_) return;
  const cuzk_tree_t *t = static_cast<const cuzk_tree_t *>(device_tree_.get());
  const size_t nlevels = cuzk_tree_num_levels(t);
  size_t width = cuzk_merkle_padded_leaves(leaf_count_, (unsigned)config_.arity);
  tree_levels_.assign(nlevels, {});
  for (size_t l = 0; l < nlevels; ++l) {
    tree_levels_[l].resize(width);
    if (cuzk_tree_level(t, l, raw(tree_levels_[l]), CUZK_MEM_HOST, nullptr) != CUZK_OK) {
      std::cerr << "CudaNaryMerkleTree: level download failed: " << cuzk_last_error() << std::endl;
      tree_levels_.clear();
      return;
    }
    width /= config_.arity;
  }
  levels_on_host_ = true;
}

// takes over the flat level-major array the library produced for `leaves` (level 0 = padded leaves, last = root)
void CudaNaryMerkleTree::adopt_levels(const std::vector<FieldElement> &leaves, const FieldElement *flat) {
  const unsigned arity = (unsigned)config_.arity;
  const size_t n = leaves.size();
  leaves_ = leaves;
  leaf_count_ = n;
  tree_height_ = cuzk_merkle_tree_height(n, arity);  // the reference's float formula: a getter value only
  device_tree_.reset();   // forest builds hand every level over on the host
  levels_on_host_ = true;
  const size_t nlevels = cuzk_merkle_num_levels(n, arity);
  tree_levels_.clear();
  tree_levels_.reserve(nlevels);
  size_t width = cuzk_merkle_padded_leaves(n, arity), at = 0;
  for (size_t l = 0; l < nlevels; ++l) {
    tree_levels_.emplace_back(flat + at, flat + at + width);
    at += width;
    width /= arity;
  }
  root_ = tree_levels_.back()[0];
}

// proofs gathered by the GPU from the levels in HBM (cuzk_tree_prove_batch) and unpacked into MerkleProof objects
bool CudaNaryMerkleTree::proofs_from_device(const std::vector<size_t> &valid_leaves, std::vector<MerkleProof> &out) const {
  const size_t arity = config_.arity, q = valid_leaves.size(), per = arity - 1;
  const size_t nlv = mg_tree_ ? cuzk_mg_tree_num_levels(static_cast<const cuzk_mg_tree_t *>(mg_tree_.get())) - 1
                              : cuzk_tree_num_levels(static_cast<const cuzk_tree_t *>(device_tree_.get())) - 1;
  std::vector<uint64_t> idx(valid_leaves.begin(), valid_leaves.end());
  std::vector<FieldElement> sib(q * nlv * per);
  std::vector<uint32_t> pos(q * nlv);
  size_t levels_seen = 0;
  if (!device_flat_proofs(idx.data(), q, raw(sib), pos.data(), levels_seen)) {
    std::cerr << "CudaNaryMerkleTree: device proof generation failed: " << cuzk_last_error() << std::endl;
    return false;
  }
  out.reserve(out.size() + q);
  for (size_t k = 0; k < q; ++k) {
    out.emplace_back();
    MerkleProof &p = out.back();
    p.leaf_index = valid_leaves[k];
    p.path.resize(nlv);
    p.indices.resize(nlv);
    for (size_t l = 0; l < nlv; ++l) {
      const FieldElement *src = sib.data() + (k * nlv + l) * per;
      p.path[l].assign(src, src + per);
      p.indices[l] = pos[k * nlv + l];
    }
  }
  device_proofs_served_ += q;
  return true;
}

std::optional<MerkleProof> CudaNaryMerkleTree::generate_proof(size_t leaf_index) const {
  if (leaf_index >= leaf_count_ || !has_tree()) return std::nullopt;
  if (mg_tree_) {   // sharded tree: always served by the GPU that owns the leaf
    std::vector<MerkleProof> one;
    if (proofs_from_device({leaf_index}, one) && !one.empty()) return std::move(one[0]);
    return std::nullopt;
  }
  // a few single proofs are served from the device tree; a caller that keeps asking gets the host copy (one download)
  if (!levels_on_host_ && device_proofs_served_ < 64) {
    std::vector<MerkleProof> one;
    if (proofs_from_device({leaf_index}, one)) return std::move(one[0]);
  }
  fetch_levels();
  if (tree_levels_.empty()) return std::nullopt;
  const size_t arity = config_.arity, nlv = tree_levels_.size() - 1;
  MerkleProof proof;
  proof.leaf_index = leaf_index;
  proof.path.resize(nlv);
  proof.indices.resize(nlv);
  size_t idx = leaf_index;
  for (size_t l = 0; l < nlv; ++l) {
    const size_t slot = idx % arity, first = idx - slot;
    const std::vector<FieldElement> &level = tree_levels_[l];
    std::vector<FieldElement> &sib = proof.path[l];
    sib.reserve(arity - 1);
    for (size_t c = 0; c < arity; ++c)
      if (c != slot) sib.push_back(level[first + c]);
    proof.indices[l] = slot;
    idx /= arity;
  }
  return proof;
}

std::vector<MerkleProof> CudaNaryMerkleTree::generate_batch_proofs(const std::vector<size_t> &indices) const {
  std::vector<MerkleProof> proofs;
  if (!has_tree()) return proofs;
  if (mg_tree_) {
    std::vector<size_t> valid;
    valid.reserve(indices.size());
    for (size_t i : indices)
      if (i < leaf_count_) valid.push_back(i);
    if (!proofs_from_device(valid, proofs)) proofs.clear();
    return proofs;
  }
  if (!levels_on_host_) {
    // the batch is gathered on the GPU unless it is so large that downloading the tree once is cheaper
    std::vector<size_t> valid;
    valid.reserve(indices.size());
    for (size_t i : indices)
      if (i < leaf_count_) valid.push_back(i);   // invalid indices are skipped silently, as in the reference
    const cuzk_tree_t *t = static_cast<const cuzk_tree_t *>(device_tree_.get());
    const size_t proof_elems = valid.size() * (cuzk_tree_num_levels(t) - 1) * (config_.arity - 1);
    if (proof_elems < cuzk_tree_total_nodes(t) * 2 && proofs_from_device(valid, proofs)) return proofs;
    proofs.clear();
    fetch_levels();
  }
  proofs.reserve(indices.size());
  for (size_t i : indices) {
    auto p = generate_proof(i);
    if (p) proofs.push_back(std::move(*p));
  }
  return proofs;
}

bool CudaNaryMerkleTree::generate_flat_proofs(const std::vector<size_t> &leaf_indices, FlatProofBatch &out) const {
  out = FlatProofBatch();
  if (!has_tree()) return false;
  for (size_t i : leaf_indices)
    if (i >= leaf_count_) return false;
  out.arity = config_.arity;
  out.leaf_indices.assign(leaf_indices.begin(), leaf_indices.end());
  const size_t q = leaf_indices.size();
  if (device_tree_ || mg_tree_) {
    out.levels = mg_tree_ ? cuzk_mg_tree_num_levels(static_cast<const cuzk_mg_tree_t *>(mg_tree_.get())) - 1
                          : cuzk_tree_num_levels(static_cast<const cuzk_tree_t *>(device_tree_.get())) - 1;
    out.positions.resize(q * out.levels);
    out.siblings.resize(q * out.levels * (out.arity - 1));
    size_t levels_seen = 0;
    if (!device_flat_proofs(out.leaf_indices.data(), q, raw(out.siblings), out.positions.data(), levels_seen)) {
      std::cerr << "CudaNaryMerkleTree::generate_flat_proofs: " << cuzk_last_error() << std::endl;
      return false;
    }
    return true;
  }
  // forest-built trees hold their levels on the host only
  out.levels = tree_levels_.size() - 1;
  out.positions.resize(q * out.levels);
  out.siblings.resize(q * out.levels * (out.arity - 1));
  for (size_t k = 0; k < q; ++k) {
    size_t at = leaf_indices[k];
    for (size_t l = 0; l < out.levels; ++l) {
      const size_t slot = at % out.arity, first = at - slot;
      FieldElement *dst = out.siblings.data() + (k * out.levels + l) * (out.arity - 1);
      for (size_t c = 0; c < out.arity; ++c)
        if (c != slot) *dst++ = tree_levels_[l][first + c];
      out.positions[k * out.levels + l] = (uint32_t)slot;
      at /= out.arity;
    }
  }
  return true;
}

bool CudaNaryMerkleTree::verify_flat_proofs(const FlatProofBatch &batch, const std::vector<FieldElement> &leaf_values,
                                            std::vector<uint8_t> &verdicts) const {
  verdicts.assign(batch.size(), 0);
  if (!has_tree() || batch.size() != leaf_values.size() || batch.arity != config_.arity ||
      batch.positions.size() != batch.size() * batch.levels || batch.siblings.size() != batch.positions.size() * (batch.arity - 1))
    return false;
  if (batch.size() == 0) return true;
  if (!ensure_library()) return false;
  const FieldElement root = get_root_hash();
  const bool sharded = mg_tree_ && batch.levels + 1 == cuzk_mg_tree_num_levels(static_cast<const cuzk_mg_tree_t *>(mg_tree_.get()));
  const int rc = sharded ? cuzk_mg_tree_verify_batch(static_cast<const cuzk_mg_tree_t *>(mg_tree_.get()), raw(leaf_values), raw(batch.siblings),
                                                     batch.positions.data(), verdicts.data(), batch.size())
                         : cuzk_merkle_verify_batch(raw(leaf_values), raw(batch.siblings), batch.positions.data(), batch.levels,
                                                    (unsigned)batch.arity, root.limbs, verdicts.data(), batch.size(), CUZK_MEM_HOST, nullptr);
  if (rc != CUZK_OK) {
    std::cerr << "CudaNaryMerkleTree::verify_flat_proofs: " << cuzk_last_error() << std::endl;
    return false;
  }
  return true;
}

bool CudaNaryMerkleTree::verify_batch_proofs_each(const std::vector<MerkleProof> &proofs, const std::vector<FieldElement> &leaf_values,
                                                  std::vector<uint8_t> &results) const {
  results.assign(proofs.size(), 0);
  if (proofs.size() != leaf_values.size() || proofs.empty() || !has_tree()) return false;
  if (!ensure_library()) return false;
  const size_t arity = config_.arity, sib_per_level = arity - 1;
  const FieldElement root = get_root_hash();
  // Proofs of equal length go to the GPU as one level-uniform batch.  Rejected outright (verdict 0), as the reference's CPU
  // verify_proof does (merkle_tree.cpp:217-219, :228-230): path and indices of different length, or a level that does not
  // carry exactly arity - 1 siblings.
  std::map<size_t, std::vector<size_t>> by_len;
  for (size_t q = 0; q < proofs.size(); ++q) {
    const MerkleProof &p = proofs[q];
    bool well_formed = p.path.size() == p.indices.size();
    for (size_t l = 0; well_formed && l < p.path.size(); ++l) well_formed = p.path[l].size() == sib_per_level;
    if (well_formed) by_len[p.path.size()].push_back(q);
  }
  for (const auto &group : by_len) {
    const size_t L = group.first, m = group.second.size();
    // flat, level-uniform copies of the batch (the layout of cuzk_merkle_verify_batch); raw buffers: every slot is written
    // exactly once below, so there is nothing to gain from value-initialising ~100 bytes per proof level first
    std::unique_ptr<uint64_t[]> leaves(new uint64_t[m * 4]), sib(new uint64_t[std::max<size_t>(1, m * L * sib_per_level * 4)]);
    std::unique_ptr<uint32_t[]> pos(new uint32_t[std::max<size_t>(1, m * L)]);
    for (size_t k = 0; k < m; ++k) {
      const MerkleProof &p = proofs[group.second[k]];
      std::memcpy(leaves.get() + 4 * k, leaf_values[group.second[k]].limbs, 32);
      for (size_t l = 0; l < L; ++l) {
        pos[k * L + l] = p.indices[l] < arity ? (uint32_t)p.indices[l] : 0xFFFFFFFFu;
        std::memcpy(sib.get() + (k * L + l) * sib_per_level * 4, p.path[l].data(), sib_per_level * 32);
      }
    }
    std::vector<uint8_t> res(m);
    if (cuzk_merkle_verify_batch(leaves.get(), sib.get(), pos.get(), L, (unsigned)arity, root.limbs, res.data(), m, CUZK_MEM_HOST, nullptr) !=
        CUZK_OK) {
      std::cerr << "CudaNaryMerkleTree::verify_batch_proofs: " << cuzk_last_error() << std::endl;
      return false;
    }
    for (size_t k = 0; k < m; ++k) results[group.second[k]] = res[k];
  }
  return true;
}

bool CudaNaryMerkleTree::verify_batch_proofs(const std::vector<MerkleProof> &proofs, const std::vector<FieldElement> &leaf_values) const {
  std::vector<uint8_t> results;
  if (!verify_batch_proofs_each(proofs, leaf_values, results)) return false;
  return std::all_of(results.begin(), results.end(), [](uint8_t r) { return r != 0; });
}

bool CudaNaryMerkleTree::verify_proof(const MerkleProof &proof, const FieldElement &leaf_value) const {
  if (!has_tree()) return false;
  return verify_batch_proofs(std::vector<MerkleProof>{proof}, std::vector<FieldElement>{leaf_value});
}

bool CudaNaryMerkleTree::build_batch_trees(const std::vector<std::vector<FieldElement>> &batch_leaves, std::vector<CudaNaryMerkleTree> &trees,
                                           const MerkleTreeConfig &config) {
  trees.clear();
  trees.reserve(batch_leaves.size());
  if (batch_leaves.empty()) return true;
  const size_t n = batch_leaves[0].size();
  bool uniform = n > 0;
  for (const auto &leaves : batch_leaves) uniform = uniform && leaves.size() == n;
  if (!uniform) {  // ragged batch: one build per tree, like the reference's loop (merkle_tree_cuda.cu:467-482)
    for (const auto &leaves : batch_leaves) {
      CudaNaryMerkleTree tree(config);
      if (!tree.build_tree(leaves)) return false;
      trees.push_back(std::move(tree));
    }
    return true;
  }
  // equal-sized trees: the whole forest is built by one pass of level launches (cuzk_merkle_build_batch)
  if (!ensure_library()) return false;
  const size_t count = batch_leaves.size(), per_tree = cuzk_merkle_total_nodes(n, (unsigned)config.arity);
  std::vector<FieldElement> all_leaves;
  all_leaves.reserve(count * n);
  for (const auto &leaves : batch_leaves) all_leaves.insert(all_leaves.end(), leaves.begin(), leaves.end());
  std::vector<FieldElement> flat(count * per_tree);
  if (cuzk_merkle_build_batch(raw(all_leaves), n, count, (unsigned)config.arity, raw(flat), CUZK_MEM_HOST, nullptr) != CUZK_OK) {
    std::cerr << "CudaNaryMerkleTree::build_batch_trees: " << cuzk_last_error() << std::endl;
    return false;
  }
  for (size_t t = 0; t < count; ++t) {
    trees.emplace_back(config);
    trees.back().adopt_levels(batch_leaves[t], flat.data() + t * per_tree);
  }
  return true;
}

FieldElement CudaNaryMerkleTree::get_root_hash() const {
  if (!has_tree()) return compute_empty_hash(config_.arity);  // merkle_tree.cpp:304-309
  return root_;
}

FieldElement CudaNaryMerkleTree::compute_empty_hash(size_t arity) const {
  FieldElement e;
  if (!ensure_library() || cuzk_merkle_empty_hash((unsigned)arity, e.limbs) != CUZK_OK)
    std::cerr << "CudaNaryMerkleTree: empty hash unavailable: " << cuzk_last_error() << std::endl;
  return e;
}

void CudaNaryMerkleTree::print_tree() const {
  fetch_levels();
  std::cout << "CUDA Merkle Tree (arity=" << config_.arity << ", height=" << tree_height_ << "):" << std::endl;
  for (size_t l = tree_levels_.size(); l-- > 0;) {
    const auto &level = tree_levels_[l];
    std::cout << "Level " << l << ": ";
    for (size_t i = 0; i < std::min<size_t>(level.size(), 8); ++i) std::cout << level[i].to_hex().substr(0, 8) << "... ";
    if (level.size() > 8) std::cout << "(+" << (level.size() - 8) << " more)";
    std::cout << std::endl;
  }
}

bool CudaNaryMerkleTree::initialize_cuda() { return ensure_library(); }

void CudaNaryMerkleTree::cleanup_cuda() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_init_refs == 0) return;
  g_init_refs = 0;
  cuzk_shutdown();  // one reference; the library stays up while hashers or CudaFieldArithmetic still hold theirs
}

size_t CudaNaryMerkleTree::get_optimal_batch_size() {
  cuzk_device_info_t info;
  if (cuzk_device_info(0, &info) != CUZK_OK) return 1024;
  return (size_t)info.sm_count * 6 * 128;
}
size_t CudaNaryMerkleTree::get_max_batch_size() { return (size_t)1 << 31; }

namespace CudaMerkleUtils {

MerkleTreeConfig get_optimal_config_for_gpu(size_t leaf_count) {
  // wider nodes mean fewer levels and fewer permutations per leaf: arity 2 costs 1 permutation per leaf, arity 8 costs 4/7
  const size_t arity = leaf_count < 1000 ? 2 : (leaf_count > 100000 ? 8 : 4);
  return MerkleTreeConfig(arity, cuzk_merkle_tree_height(leaf_count, (unsigned)arity));
}

bool check_cuda_compatibility() {
  if (cuzk_device_count() <= 0) {
    std::cerr << "No CUDA-capable devices found" << std::endl;
    return false;
  }
  cuzk_device_info_t info;
  if (cuzk_device_info(0, &info) != CUZK_OK) return false;
  if (info.cc_major < 10) {
    std::cerr << "cuzk_b200 needs a compute capability 10.0 (B200) device, found " << info.cc_major << "." << info.cc_minor << std::endl;
    return false;
  }
  return true;
}

std::vector<std::vector<FieldElement>> generate_batch_test_leaves(size_t num_trees, size_t leaves_per_tree, uint64_t seed) {
  std::vector<std::vector<FieldElement>> batch(num_trees);
  for (size_t t = 0; t < num_trees; ++t) {
    batch[t].reserve(leaves_per_tree);
    for (size_t i = 0; i < leaves_per_tree; ++i) batch[t].emplace_back(seed + t * leaves_per_tree + i);  // merkle_tree_cuda.cu:623-643
  }
  return batch;
}

}  // namespace CudaMerkleUtils

namespace {
double ms_since(std::chrono::steady_clock::time_point t0) {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
CudaMerkleTreeStats fresh_stats(size_t leaves, size_t arity) {
  CudaMerkleTreeStats s = {};
  s.leaf_count = leaves;
  s.arity = arity;
  return s;
}
}  // namespace

CudaMerkleTreeStats benchmark_cuda_tree_building(size_t num_trees, size_t leaves_per_tree, size_t arity, size_t /*batch_size*/) {
  CudaMerkleTreeStats stats = fresh_stats(leaves_per_tree, arity);
  stats.total_trees = num_trees;
  if (!CudaNaryMerkleTree::initialize_cuda()) return stats;
  const auto batch = CudaMerkleUtils::generate_batch_test_leaves(num_trees, leaves_per_tree, 12345);
  std::vector<CudaNaryMerkleTree> trees;
  const auto t0 = std::chrono::steady_clock::now();
  const bool ok = CudaNaryMerkleTree::build_batch_trees(batch, trees, MerkleTreeConfig(arity));
  const double ms = ms_since(t0);
  if (!ok || trees.empty()) {
    std::cerr << "CUDA tree building failed" << std::endl;
    return stats;
  }
  stats.build_time_ms = stats.total_time_ms = ms;
  stats.tree_height = trees[0].get_tree_height();
  if (ms > 0) stats.trees_per_second = static_cast<size_t>(num_trees * 1000.0 / ms);
  return stats;
}

namespace {
std::vector<size_t> random_indices(size_t count, size_t bound) {
  std::mt19937_64 gen(std::random_device{}());
  std::uniform_int_distribution<size_t> pick(0, bound - 1);
  std::vector<size_t> idx(count);
  for (auto &i : idx) i = pick(gen);
  return idx;
}
}  // namespace

CudaMerkleTreeStats benchmark_cuda_proof_generation(size_t num_proofs, size_t leaves_per_tree, size_t arity, size_t /*batch_size*/) {
  CudaMerkleTreeStats stats = fresh_stats(leaves_per_tree, arity);
  stats.total_proofs = num_proofs;
  if (!CudaNaryMerkleTree::initialize_cuda() || leaves_per_tree == 0) return stats;
  CudaNaryMerkleTree tree(CudaMerkleUtils::generate_batch_test_leaves(1, leaves_per_tree, 54321)[0], MerkleTreeConfig(arity));
  stats.tree_height = tree.get_tree_height();
  const auto idx = random_indices(num_proofs, leaves_per_tree);
  const auto t0 = std::chrono::steady_clock::now();
  const auto proofs = tree.generate_batch_proofs(idx);
  const double ms = ms_since(t0);
  stats.proof_generation_time_ms = stats.total_time_ms = ms;
  if (ms > 0) stats.proofs_per_second = static_cast<size_t>(proofs.size() * 1000.0 / ms);
  return stats;
}

CudaMerkleTreeStats benchmark_cuda_proof_verification(size_t num_proofs, size_t leaves_per_tree, size_t arity, size_t /*batch_size*/) {
  CudaMerkleTreeStats stats = fresh_stats(leaves_per_tree, arity);
  stats.total_proofs = num_proofs;
  if (!CudaNaryMerkleTree::initialize_cuda() || leaves_per_tree == 0) return stats;
  const auto leaves = CudaMerkleUtils::generate_batch_test_leaves(1, leaves_per_tree, 98765)[0];
  CudaNaryMerkleTree tree(leaves, MerkleTreeConfig(arity));
  stats.tree_height = tree.get_tree_height();
  const auto idx = random_indices(num_proofs, leaves_per_tree);
  std::vector<FieldElement> values;
  values.reserve(num_proofs);
  for (size_t i : idx) values.push_back(leaves[i]);
  const auto proofs = tree.generate_batch_proofs(idx);
  const auto t0 = std::chrono::steady_clock::now();
  const bool ok = tree.verify_batch_proofs(proofs, values);
  const double ms = ms_since(t0);
  stats.proof_verification_time_ms = stats.total_time_ms = ms;
  if (ms > 0) stats.proofs_per_second = static_cast<size_t>(num_proofs * 1000.0 / ms);
  if (!ok) std::cerr << "Warning: Batch proof verification failed in benchmark" << std::endl;
  return stats;
}

}  // namespace MerkleTreeCUDA
}  // namespace MerkleTree
