// cuda/poseidon_cuda_optimized.cuh -- CudaPoseidonHashOptimized.
//
// Replaces the reference's src/poseidon/cuda/poseidon_cuda_optimized.cuh:26-62.  The reference keeps two copies of
// its kernels ("original" and "optimized"); here one set of sm_100a kernels backs both class names, so the benchmark
// driver that instantiates both (src/poseidon/test/benchmark.cpp:132-133) and verify_cuda_implementations_match
// keep working unchanged.
#pragma once

#include "poseidon_cuda.cuh"

namespace Poseidon {
namespace PoseidonCUDAOptimized {

using namespace Poseidon::CudaFieldOps;
using Poseidon::PoseidonCUDA::IPoseidonCudaHash;

class CudaPoseidonHashOptimized : public IPoseidonCudaHash {
public:
  CudaPoseidonHashOptimized() = default;
  ~CudaPoseidonHashOptimized() override = default;

  bool batch_hash_single(const std::vector<FieldElement> &inputs, std::vector<FieldElement> &outputs) override {
    return impl_.batch_hash_single(inputs, outputs);
  }
  bool batch_hash_pairs(const std::vector<FieldElement> &left_inputs, const std::vector<FieldElement> &right_inputs,
                        std::vector<FieldElement> &outputs) override {
    return impl_.batch_hash_pairs(left_inputs, right_inputs, outputs);
  }
  bool batch_permutation(std::vector<std::array<CudaFieldElement, PoseidonParams::STATE_SIZE>> &states) override {
    return impl_.batch_permutation(states);
  }
  size_t get_optimal_batch_size() const override { return impl_.get_optimal_batch_size(); }
  size_t get_max_batch_size() const override { return impl_.get_max_batch_size(); }
  bool is_initialized() const override { return impl_.is_initialized(); }

private:
  Poseidon::PoseidonCUDA::CudaPoseidonHash impl_;
};

}  // namespace PoseidonCUDAOptimized
}  // namespace Poseidon
