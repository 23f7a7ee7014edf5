// cuda/poseidon_interface_cuda.hpp -- IPoseidonCudaHash, the GPU batch-hash contract.
// Replaces the reference's src/poseidon/cuda/poseidon_interface_cuda.hpp:15-47 (identical virtual interface).
#pragma once

#include <array>
#include <vector>

#include "../poseidon.hpp"
#include "cuda_field_element.cuh"

namespace Poseidon {
namespace PoseidonCUDA {

using namespace Poseidon::CudaFieldOps;

struct CudaPoseidonStats {
  double total_time_ms;
  double avg_time_per_hash_ns;
  size_t hashes_per_second;
  size_t total_hashes;
  double speedup_vs_cpu;
};

class IPoseidonCudaHash {
public:
  virtual ~IPoseidonCudaHash() = default;

  // outputs[i] = PoseidonHash::hash_single(inputs[i]); outputs is resized; empty input -> true, empty output
  virtual bool batch_hash_single(const std::vector<FieldElement> &inputs, std::vector<FieldElement> &outputs) = 0;
  // outputs[i] = PoseidonHash::hash_pair(left[i], right[i]); size mismatch -> false
  virtual bool batch_hash_pairs(const std::vector<FieldElement> &left_inputs, const std::vector<FieldElement> &right_inputs,
                                std::vector<FieldElement> &outputs) = 0;
  // in-place PoseidonHash::permutation on every state
  virtual bool batch_permutation(std::vector<std::array<CudaFieldElement, PoseidonParams::STATE_SIZE>> &states) = 0;

  virtual size_t get_optimal_batch_size() const = 0;
  virtual size_t get_max_batch_size() const = 0;
  virtual bool is_initialized() const = 0;
};

}  // namespace PoseidonCUDA
}  // namespace Poseidon
