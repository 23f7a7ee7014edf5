// cuda/field_arithmetic_cuda.cuh -- CudaFieldArithmetic: element-wise batch field operations on the GPU.
//
// Replaces the reference's src/poseidon/cuda/field_arithmetic_cuda.cuh:25-91 (same class, same static methods,
// same bool-and-stderr error style).  The five reference kernels (:18-22) are gone from the header: the kernels
// live in libcuzk_b200.so and are reached through cuzk_fr_batch (include/cuzk_b200.h).  Differences on purpose:
// initialize()/cleanup() are reference-counted and cleanup() never resets the device (the reference's does,
// field_arithmetic_cuda.cu:355-360, invalidating every other live CUDA object); the gpu_* single-element
// functions, which the reference declares (:53-57) but never defines, are defined here as one-element batches.
#pragma once

#include <vector>

#include "../field_arithmetic.hpp"
#include "cuda_field_element.cuh"

namespace Poseidon {
namespace CudaFieldOps {

class CudaFieldArithmetic {
public:
  static bool initialize();
  static void cleanup();

  static bool batch_add(const std::vector<FieldElement> &a, const std::vector<FieldElement> &b, std::vector<FieldElement> &result);
  static bool batch_subtract(const std::vector<FieldElement> &a, const std::vector<FieldElement> &b, std::vector<FieldElement> &result);
  static bool batch_multiply(const std::vector<FieldElement> &a, const std::vector<FieldElement> &b, std::vector<FieldElement> &result);
  static bool batch_square(const std::vector<FieldElement> &input, std::vector<FieldElement> &result);
  static bool batch_power5(const std::vector<FieldElement> &input, std::vector<FieldElement> &result);

  static bool gpu_add(const FieldElement &a, const FieldElement &b, FieldElement &result);
  static bool gpu_subtract(const FieldElement &a, const FieldElement &b, FieldElement &result);
  static bool gpu_multiply(const FieldElement &a, const FieldElement &b, FieldElement &result);
  static bool gpu_square(const FieldElement &a, FieldElement &result);
  static bool gpu_power5(const FieldElement &a, FieldElement &result);

  static int get_device_count();
  static void print_device_info();
  static size_t get_optimal_block_size();

private:
  static int init_refs_;  // how many initialize() calls this class still owes a cuzk_shutdown() for
  static bool run(int op, const std::vector<FieldElement> &a, const std::vector<FieldElement> *b, std::vector<FieldElement> &result,
                  const char *what);
};

struct CudaHashingStats {
  double total_time_ms;
  double avg_time_per_operation_ns;
  size_t operations_per_second;
  size_t total_operations;
};

CudaHashingStats benchmark_cuda_field_operations(size_t num_operations, size_t batch_size = 1024);

}  // namespace CudaFieldOps
}  // namespace Poseidon
