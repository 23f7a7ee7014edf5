// field_arithmetic_cuda.cpp -- CudaFieldArithmetic over the C ABI (replaces src/poseidon/cuda/field_arithmetic_cuda.cu:315-700).
// Host C++ only: the kernels are in libcuzk_b200.so (cuzk_fr_batch).  std::vector<FieldElement> is already the
// packed 4 x u64 layout the C ABI takes, so there is no per-element conversion pass.
#include "field_arithmetic_cuda.cuh"

#include <chrono>
#include <iostream>
#include <mutex>
#include <random>

#include "cuzk_b200.h"

namespace Poseidon {
namespace CudaFieldOps {

namespace {
std::mutex g_mu;
const uint64_t *raw(const std::vector<FieldElement> &v) { return reinterpret_cast<const uint64_t *>(v.data()); }
uint64_t *raw(std::vector<FieldElement> &v) { return reinterpret_cast<uint64_t *>(v.data()); }
}  // namespace

int CudaFieldArithmetic::init_refs_ = 0;

bool CudaFieldArithmetic::initialize() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (init_refs_ > 0) return true;  // idempotent, like the reference (field_arithmetic_cuda.cu:316)
  if (cuzk_init(0) != CUZK_OK) {
    std::cerr << "CudaFieldArithmetic::initialize: " << cuzk_last_error() << std::endl;
    return false;
  }
  init_refs_ = 1;
  return true;
}

void CudaFieldArithmetic::cleanup() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (init_refs_ == 0) return;
  init_refs_ = 0;
  cuzk_shutdown();  // drops this class's reference only; other live hashers / trees keep the library up
}

bool CudaFieldArithmetic::run(int op, const std::vector<FieldElement> &a, const std::vector<FieldElement> *b,
                              std::vector<FieldElement> &result, const char *what) {
  if (init_refs_ == 0 && !cuzk_is_initialized()) {
    std::cerr << "CUDA not initialized" << std::endl;
    return false;
  }
  if (b && a.size() != b->size()) {
    std::cerr << what << ": input vectors must have the same size" << std::endl;
    return false;
  }
  if (a.empty()) {
    result.clear();
    return true;
  }
  std::vector<FieldElement> out(a.size());  // separate buffer: `result` may alias an input
  if (cuzk_fr_batch(op, raw(a), b ? raw(*b) : nullptr, raw(out), a.size(), CUZK_MEM_HOST, nullptr) != CUZK_OK) {
    std::cerr << what << ": " << cuzk_last_error() << std::endl;
    return false;
  }
  result.swap(out);
  return true;
}

bool CudaFieldArithmetic::batch_add(const std::vector<FieldElement> &a, const std::vector<FieldElement> &b, std::vector<FieldElement> &r) {
  return run(CUZK_FR_ADD, a, &b, r, "batch_add");
}
bool CudaFieldArithmetic::batch_subtract(const std::vector<FieldElement> &a, const std::vector<FieldElement> &b, std::vector<FieldElement> &r) {
  return run(CUZK_FR_SUB, a, &b, r, "batch_subtract");
}
bool CudaFieldArithmetic::batch_multiply(const std::vector<FieldElement> &a, const std::vector<FieldElement> &b, std::vector<FieldElement> &r) {
  return run(CUZK_FR_MUL, a, &b, r, "batch_multiply");
}
bool CudaFieldArithmetic::batch_square(const std::vector<FieldElement> &in, std::vector<FieldElement> &r) {
  return run(CUZK_FR_SQR, in, nullptr, r, "batch_square");
}
bool CudaFieldArithmetic::batch_power5(const std::vector<FieldElement> &in, std::vector<FieldElement> &r) {
  return run(CUZK_FR_POW5, in, nullptr, r, "batch_power5");
}

namespace {
bool single(int op, const FieldElement &a, const FieldElement *b, FieldElement &result) {
  FieldElement out;
  if (cuzk_fr_batch(op, a.limbs, b ? b->limbs : nullptr, out.limbs, 1, CUZK_MEM_HOST, nullptr) != CUZK_OK) {
    std::cerr << "CudaFieldArithmetic: " << cuzk_last_error() << std::endl;
    return false;
  }
  result = out;
  return true;
}
}  // namespace
bool CudaFieldArithmetic::gpu_add(const FieldElement &a, const FieldElement &b, FieldElement &r) { return single(CUZK_FR_ADD, a, &b, r); }
bool CudaFieldArithmetic::gpu_subtract(const FieldElement &a, const FieldElement &b, FieldElement &r) { return single(CUZK_FR_SUB, a, &b, r); }
bool CudaFieldArithmetic::gpu_multiply(const FieldElement &a, const FieldElement &b, FieldElement &r) { return single(CUZK_FR_MUL, a, &b, r); }
bool CudaFieldArithmetic::gpu_square(const FieldElement &a, FieldElement &r) { return single(CUZK_FR_SQR, a, nullptr, r); }
bool CudaFieldArithmetic::gpu_power5(const FieldElement &a, FieldElement &r) { return single(CUZK_FR_POW5, a, nullptr, r); }

int CudaFieldArithmetic::get_device_count() { return cuzk_device_count(); }

void CudaFieldArithmetic::print_device_info() {
  cuzk_device_info_t info;
  if (cuzk_device_info(0, &info) != CUZK_OK) {
    std::cout << "CUDA Device Info: unavailable (" << cuzk_last_error() << ")" << std::endl;
    return;
  }
  std::cout << "CUDA Device Info:" << std::endl;
  std::cout << "  Name: " << info.name << std::endl;
  std::cout << "  Compute Capability: " << info.cc_major << "." << info.cc_minor << std::endl;
  std::cout << "  Memory: " << info.total_mem_bytes / (1024 * 1024) << " MB" << std::endl;
  std::cout << "  Max Threads per Block: " << info.max_threads_per_block << std::endl;
  std::cout << "  Multiprocessors: " << info.sm_count << std::endl;
  std::cout << "  Optimal Block Size: " << get_optimal_block_size() << std::endl;
}

size_t CudaFieldArithmetic::get_optimal_block_size() { return 128; }  // CTA size of every cuzk_b200 kernel

// element-wise multiplies through batch_multiply in batches of `batch_size`, like the reference's helper
// (field_arithmetic_cuda.cu:702-774): host vectors in, host vectors out, wall-clock time
CudaHashingStats benchmark_cuda_field_operations(size_t num_operations, size_t batch_size) {
  CudaHashingStats stats = {};
  if (batch_size == 0 || !CudaFieldArithmetic::initialize()) return stats;
  std::mt19937_64 gen(20261018);
  std::vector<FieldElement> a(batch_size), b(batch_size), out;
  for (size_t i = 0; i < batch_size; ++i) {
    a[i] = FieldElement(gen(), gen(), gen(), gen() >> 4);
    b[i] = FieldElement(gen(), gen(), gen(), gen() >> 4);
  }
  size_t done = 0;
  const auto t0 = std::chrono::steady_clock::now();
  while (done < num_operations) {
    const size_t m = std::min(batch_size, num_operations - done);
    if (m != a.size()) {
      a.resize(m);
      b.resize(m);
    }
    if (!CudaFieldArithmetic::batch_multiply(a, b, out)) break;
    done += m;
  }
  const double ns = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count();
  stats.total_time_ms = ns / 1e6;
  stats.total_operations = done;
  if (done) {
    stats.avg_time_per_operation_ns = ns / done;
    stats.operations_per_second = static_cast<size_t>(1e9 / stats.avg_time_per_operation_ns);
  }
  return stats;
}

}  // namespace CudaFieldOps
}  // namespace Poseidon
