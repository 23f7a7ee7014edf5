// merkle_tree_cuda_vs_cpu.cpp -- benchmark_cuda_vs_cpu_merkle (replaces src/merkle_tree/merkle_tree_cuda.cu:742-806).
// Builds the same batches with the reference's CPU NaryMerkleTree and with CudaNaryMerkleTree, reports the ratio and
// checks the roots agree; compiled only where the reference's CPU library is linked -- never into libcuzk_host itself.
#include <chrono>
#include <iostream>

#include "merkle_tree_cuda.cuh"

namespace MerkleTree {
namespace MerkleTreeCUDA {

CudaMerkleTreeStats benchmark_cuda_vs_cpu_merkle(size_t num_trees, size_t leaves_per_tree, size_t arity, size_t /*batch_size*/) {
  CudaMerkleTreeStats stats = {};
  stats.leaf_count = leaves_per_tree;
  stats.arity = arity;
  stats.total_trees = num_trees;
  if (!CudaNaryMerkleTree::initialize_cuda()) return stats;
  const auto batch = CudaMerkleUtils::generate_batch_test_leaves(num_trees, leaves_per_tree, 11111);
  const MerkleTreeConfig config(arity);
  using clock = std::chrono::steady_clock;

  auto t0 = clock::now();
  std::vector<NaryMerkleTree> cpu_trees;
  cpu_trees.reserve(num_trees);
  for (const auto &leaves : batch) cpu_trees.emplace_back(leaves, config);
  const double cpu_ms = std::chrono::duration<double, std::milli>(clock::now() - t0).count();

  t0 = clock::now();
  std::vector<CudaNaryMerkleTree> gpu_trees;
  const bool ok = CudaNaryMerkleTree::build_batch_trees(batch, gpu_trees, config);
  const double gpu_ms = std::chrono::duration<double, std::milli>(clock::now() - t0).count();
  if (!ok || gpu_trees.empty()) {
    std::cerr << "CUDA tree building failed in comparison benchmark" << std::endl;
    return stats;
  }
  stats.build_time_ms = stats.total_time_ms = gpu_ms;
  stats.tree_height = gpu_trees[0].get_tree_height();
  if (gpu_ms > 0) {
    stats.speedup_vs_cpu = cpu_ms / gpu_ms;
    stats.trees_per_second = static_cast<size_t>(num_trees * 1000.0 / gpu_ms);
  }
  for (size_t i = 0; i < gpu_trees.size() && i < cpu_trees.size(); ++i) {
    if (!gpu_trees[i].compare_with_cpu_tree(cpu_trees[i])) {
      std::cerr << "Warning: CPU and CUDA trees are not consistent!" << std::endl;
      break;
    }
  }
  return stats;
}

}  // namespace MerkleTreeCUDA
}  // namespace MerkleTree
