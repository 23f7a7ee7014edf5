// cuda/poseidon_interface_cuda.hpp -- the GPU batch-hash contract of the cuZK host interface.
//
// Stands in for the reference's src/poseidon/cuda/poseidon_interface_cuda.hpp:15-47: the abstract class IPoseidonCudaHash
// (three batch operations, two sizing hints, one status query) and the statistics record the benchmark helpers fill in.
// Signatures are the reference's, so callers written against it -- its tests, src/poseidon/test/benchmark.cpp, the helpers in
// poseidon_cuda_benchmarks.hpp -- compile unchanged; everything behind the interface is libcuzk_b200.so.
//
// Semantics every implementation here guarantees (bit-exact with the reference's CPU PoseidonHash):
//   batch_hash_single : y[i] = hash_single(x[i])        -- sponge, domain separator 1, inputs may be any 256-bit value
//   batch_hash_pairs  : y[i] = hash_pair(l[i], r[i])    -- sponge, domain separator 2; l and r must have equal length
//   batch_permutation : state[i] <- permutation(state[i]) in place, 3 elements per state
// Return value true = done (output vector resized to the batch size; an empty batch gives an empty output and true),
// false = not initialised, mismatched sizes or a CUDA error (message on std::cerr).  Nothing throws.
#pragma once

#include <array>
#include <cstddef>
#include <vector>

#include "../poseidon.hpp"
#include "cuda_field_element.cuh"

namespace Poseidon {
namespace PoseidonCUDA {

using namespace Poseidon::CudaFieldOps;

// one Poseidon state as the batch_permutation entry point takes it: t = 3 elements of 32 bytes, packed
using CudaPoseidonState = std::array<CudaFieldElement, PoseidonParams::STATE_SIZE>;

class IPoseidonCudaHash {
public:
  virtual ~IPoseidonCudaHash() = default;

  virtual bool batch_hash_single(const std::vector<FieldElement> &x, std::vector<FieldElement> &y) = 0;
  virtual bool batch_hash_pairs(const std::vector<FieldElement> &l, const std::vector<FieldElement> &r, std::vector<FieldElement> &y) = 0;
  virtual bool batch_permutation(std::vector<CudaPoseidonState> &states) = 0;

  virtual size_t get_optimal_batch_size() const = 0;  // units in one resident wave of the hashing kernels
  virtual size_t get_max_batch_size() const = 0;      // sanity bound; larger host batches are chunked by the library anyway
  virtual bool is_initialized() const = 0;
};

// filled in by benchmark_cuda_poseidon_* (poseidon_cuda_benchmarks.hpp); wall-clock figures over host-vector calls
struct CudaPoseidonStats {
  double total_time_ms;
  double avg_time_per_hash_ns;
  size_t hashes_per_second;
  size_t total_hashes;
  double speedup_vs_cpu;  // only benchmark_cuda_vs_cpu_poseidon sets it
};

}  // namespace PoseidonCUDA
}  // namespace Poseidon
