// poseidon_cuda_benchmarks.cpp -- measurement helpers over IPoseidonCudaHash (replaces
// src/poseidon/cuda/poseidon_cuda_benchmarks.cpp:12-117 and :137-259).  Same protocol as the reference harness:
// `batch_size` random inputs generated once, then ceil(n / batch_size) synchronous batch calls on host vectors,
// wall-clock time over the whole loop.  Inputs come from a local mt19937_64 (canonical, < 2^252) instead of
// FieldElement::random(), so these helpers do not need the CPU library.
#include "poseidon_cuda_benchmarks.hpp"

#include <algorithm>
#include <chrono>
#include <iostream>
#include <random>

namespace Poseidon {
namespace PoseidonCUDA {

namespace {

std::vector<FieldElement> random_elements(size_t n, uint64_t seed) {
  std::mt19937_64 gen(seed);
  std::vector<FieldElement> v(n);
  for (auto &e : v) e = FieldElement(gen(), gen(), gen(), gen() >> 4);
  return v;
}

template <class Call>
CudaPoseidonStats timed_batches(IPoseidonCudaHash &hasher, size_t total, size_t batch_size, Call call) {
  CudaPoseidonStats stats = {};
  if (!hasher.is_initialized()) {
    std::cerr << "CUDA Poseidon hasher not initialized for benchmarking" << std::endl;
    return stats;
  }
  if (batch_size == 0) return stats;
  size_t done = 0;
  const auto t0 = std::chrono::steady_clock::now();
  for (size_t at = 0; at < total; at += batch_size) {
    const size_t m = std::min(batch_size, total - at);
    if (call(m)) done += m;
  }
  const double ns = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count();
  stats.total_time_ms = ns / 1e6;
  stats.total_hashes = done;
  if (done) {
    stats.avg_time_per_hash_ns = ns / done;
    stats.hashes_per_second = static_cast<size_t>(1e9 / stats.avg_time_per_hash_ns);
  }
  return stats;
}

}  // namespace

CudaPoseidonStats benchmark_cuda_poseidon_single(IPoseidonCudaHash &hasher, size_t num_hashes, size_t batch_size) {
  std::vector<FieldElement> in = random_elements(batch_size, 11), out;
  return timed_batches(hasher, num_hashes, batch_size, [&](size_t m) {
    if (m != in.size()) in.resize(m);
    return hasher.batch_hash_single(in, out);
  });
}

CudaPoseidonStats benchmark_cuda_poseidon_pairs(IPoseidonCudaHash &hasher, size_t num_pairs, size_t batch_size) {
  std::vector<FieldElement> l = random_elements(batch_size, 12), r = random_elements(batch_size, 13), out;
  return timed_batches(hasher, num_pairs, batch_size, [&](size_t m) {
    if (m != l.size()) {
      l.resize(m);
      r.resize(m);
    }
    return hasher.batch_hash_pairs(l, r, out);
  });
}

bool verify_cuda_implementations_match(IPoseidonCudaHash &h1, IPoseidonCudaHash &h2, const std::string &name1, const std::string &name2,
                                       size_t num_tests) {
  std::cout << "\nVerifying " << name1 << " and " << name2 << " implementations match...\n";
  if (!h1.is_initialized() || !h2.is_initialized()) {
    std::cerr << "One or both hashers not initialized for verification" << std::endl;
    return false;
  }
  // the reference's deterministic inputs (poseidon_cuda_benchmarks.cpp:160-163, :208-212): full four-limb patterns
  std::vector<FieldElement> left(num_tests), right(num_tests);
  for (size_t i = 0; i < num_tests; ++i) {
    left[i] = FieldElement(i + 1, 2 * i + 1, 3 * i + 1, 4 * i + 1);
    right[i] = FieldElement(5 * i + 1, 6 * i + 1, 7 * i + 1, 8 * i + 1);
  }
  auto compare = [&](const char *kind, const std::vector<FieldElement> &a, const std::vector<FieldElement> &b) {
    if (a.size() != b.size()) {
      std::cerr << kind << " output sizes differ: " << a.size() << " vs " << b.size() << std::endl;
      return false;
    }
    size_t bad = 0;
    for (size_t i = 0; i < a.size(); ++i) {
      if (a[i] == b[i]) continue;
      ++bad;
      std::cerr << kind << " hash mismatch at index " << i << ":\n  " << name1 << ": " << a[i].to_hex() << "\n  " << name2 << ": "
                << b[i].to_hex() << std::endl;
    }
    if (bad) std::cout << "✗ " << bad << " " << kind << " hash mismatches found" << std::endl;
    else std::cout << "✓ All " << a.size() << " " << kind << " hashes match" << std::endl;
    return bad == 0;
  };
  std::vector<FieldElement> o1, o2;
  std::cout << "Testing single hash operations..." << std::endl;
  if (!h1.batch_hash_single(left, o1) || !h2.batch_hash_single(left, o2)) {
    std::cerr << "Failed to hash with " << name1 << " or " << name2 << std::endl;
    return false;
  }
  bool ok = compare("single", o1, o2);
  std::cout << "Testing pair hash operations..." << std::endl;
  if (!h1.batch_hash_pairs(left, right, o1) || !h2.batch_hash_pairs(left, right, o2)) {
    std::cerr << "Failed to hash pairs with " << name1 << " or " << name2 << std::endl;
    return false;
  }
  ok = compare("pair", o1, o2) && ok;
  if (ok) std::cout << "✓ All verification tests passed! " << name1 << " and " << name2 << " produce identical results." << std::endl;
  else std::cout << "✗ Verification failed! " << name1 << " and " << name2 << " produce different results." << std::endl;
  return ok;
}

}  // namespace PoseidonCUDA
}  // namespace Poseidon
