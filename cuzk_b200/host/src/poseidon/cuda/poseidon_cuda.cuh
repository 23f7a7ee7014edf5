// cuda/poseidon_cuda.cuh -- CudaPoseidonHash: IPoseidonCudaHash over libcuzk_b200.so.
//
// Replaces the reference's src/poseidon/cuda/poseidon_cuda.cuh:23-59.  Construction takes a reference on the
// library (cuzk_init), destruction drops it (cuzk_shutdown) -- never a device reset.  The device function
// device_hash_n the reference exports from this header (:20) is internal to libcuzk_b200.so (csrc/poseidon.cuh,
// sponge_n); host code reaches it through cuzk_poseidon_sponge with domain separator 3.
#pragma once

#include <vector>

#include "../poseidon.hpp"
#include "cuda_field_element.cuh"
#include "field_arithmetic_cuda.cuh"
#include "poseidon_interface_cuda.hpp"

namespace Poseidon {
namespace PoseidonCUDA {

using namespace Poseidon::CudaFieldOps;

class CudaPoseidonHash : public IPoseidonCudaHash {
public:
  CudaPoseidonHash();
  ~CudaPoseidonHash() override;
  CudaPoseidonHash(const CudaPoseidonHash &) = delete;
  CudaPoseidonHash &operator=(const CudaPoseidonHash &) = delete;

  bool batch_hash_single(const std::vector<FieldElement> &inputs, std::vector<FieldElement> &outputs) override;
  bool batch_hash_pairs(const std::vector<FieldElement> &left_inputs, const std::vector<FieldElement> &right_inputs,
                        std::vector<FieldElement> &outputs) override;
  bool batch_permutation(std::vector<std::array<CudaFieldElement, PoseidonParams::STATE_SIZE>> &states) override;

  // extension (no reference counterpart): outputs[i] = PoseidonHash::sponge(inputs[i*width .. +width), domain_separator);
  // domain separator 3 = hash_multiple, the Merkle node hash
  bool batch_sponge(const std::vector<FieldElement> &inputs, size_t width, uint64_t domain_separator, std::vector<FieldElement> &outputs);

  size_t get_optimal_batch_size() const override;
  size_t get_max_batch_size() const override;
  bool is_initialized() const override;

private:
  bool initialized_;
  size_t optimal_batch_size_;
  size_t max_batch_size_;
};

}  // namespace PoseidonCUDA
}  // namespace Poseidon
