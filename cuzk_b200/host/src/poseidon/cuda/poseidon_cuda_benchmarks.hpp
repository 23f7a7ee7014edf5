// cuda/poseidon_cuda_benchmarks.hpp -- measurement helpers over IPoseidonCudaHash.
// Replaces the reference's src/poseidon/cuda/poseidon_cuda_benchmarks.hpp:10-17 (same four functions).
#pragma once

#include <string>

#include "poseidon_cuda.cuh"
#include "poseidon_interface_cuda.hpp"

namespace Poseidon {
namespace PoseidonCUDA {

CudaPoseidonStats benchmark_cuda_poseidon_single(IPoseidonCudaHash &hasher, size_t num_hashes, size_t batch_size = 1024);
CudaPoseidonStats benchmark_cuda_poseidon_pairs(IPoseidonCudaHash &hasher, size_t num_pairs, size_t batch_size = 1024);
// needs the reference's CPU library at link time (Poseidon::benchmark_poseidon); lives in its own source file
CudaPoseidonStats benchmark_cuda_vs_cpu_poseidon(IPoseidonCudaHash &hasher, size_t num_hashes, size_t batch_size = 1024);
bool verify_cuda_implementations_match(IPoseidonCudaHash &hasher1, IPoseidonCudaHash &hasher2, const std::string &name1,
                                       const std::string &name2, size_t num_tests = 100);

}  // namespace PoseidonCUDA
}  // namespace Poseidon
