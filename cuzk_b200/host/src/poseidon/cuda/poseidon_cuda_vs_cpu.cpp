// poseidon_cuda_vs_cpu.cpp -- benchmark_cuda_vs_cpu_poseidon (replaces src/poseidon/cuda/poseidon_cuda_benchmarks.cpp:119-135).
// The one helper that times the reference's CPU implementation next to the GPU: it calls Poseidon::benchmark_poseidon,
// so it is compiled only where the reference's CPU library (poseidon.cpp) is linked -- never into libcuzk_host itself.
#include "poseidon_cuda_benchmarks.hpp"

namespace Poseidon {
namespace PoseidonCUDA {

CudaPoseidonStats benchmark_cuda_vs_cpu_poseidon(IPoseidonCudaHash &hasher, size_t num_hashes, size_t batch_size) {
  const HashingStats cpu = Poseidon::benchmark_poseidon(num_hashes);
  CudaPoseidonStats stats = benchmark_cuda_poseidon_single(hasher, num_hashes, batch_size);
  if (cpu.avg_time_per_hash_ns > 0 && stats.avg_time_per_hash_ns > 0) stats.speedup_vs_cpu = cpu.avg_time_per_hash_ns / stats.avg_time_per_hash_ns;
  return stats;
}

}  // namespace PoseidonCUDA
}  // namespace Poseidon
