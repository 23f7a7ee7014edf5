// cuda/poseidon_cuda_benchmarks.hpp -- timing and cross-check helpers that work on any IPoseidonCudaHash.
//
// Stands in for the reference's src/poseidon/cuda/poseidon_cuda_benchmarks.hpp:10-17 (same four entry points, called by
// src/poseidon/test/benchmark.cpp and test_poseidon_cuda.cpp).  Protocol, as in the reference harness: one batch of random
// inputs is generated up front, then ceil(total / batch) synchronous calls on host vectors are timed with the wall clock.
//
//   benchmark_cuda_poseidon_single / _pairs   hashes per second through batch_hash_single / batch_hash_pairs
//   benchmark_cuda_vs_cpu_poseidon            the single-hash figure next to the reference CPU's benchmark_poseidon();
//                                             defined in poseidon_cuda_vs_cpu.cpp because it needs the CPU library at link time
//   verify_cuda_implementations_match         two hashers, the reference's `count` deterministic four-limb inputs, outputs
//                                             compared element by element (the benchmark driver refuses to run otherwise)
#pragma once

#include <cstddef>
#include <string>

#include "poseidon_interface_cuda.hpp"
#include "poseidon_cuda.cuh"

namespace Poseidon {
namespace PoseidonCUDA {

CudaPoseidonStats benchmark_cuda_poseidon_single(IPoseidonCudaHash &gpu, size_t total, size_t batch = 1024);

CudaPoseidonStats benchmark_cuda_poseidon_pairs(IPoseidonCudaHash &gpu, size_t total, size_t batch = 1024);

CudaPoseidonStats benchmark_cuda_vs_cpu_poseidon(IPoseidonCudaHash &gpu, size_t total, size_t batch = 1024);

bool verify_cuda_implementations_match(IPoseidonCudaHash &first, IPoseidonCudaHash &second, const std::string &first_name,
                                       const std::string &second_name, size_t count = 100);

}  // namespace PoseidonCUDA
}  // namespace Poseidon
