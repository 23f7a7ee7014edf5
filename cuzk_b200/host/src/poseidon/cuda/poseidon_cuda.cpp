// poseidon_cuda.cpp -- CudaPoseidonHash over the C ABI (replaces src/poseidon/cuda/poseidon_cuda.cu:212-487 and
// poseidon_cuda_optimized.cu:279-554; the kernels those files hold are in libcuzk_b200.so).
#include "poseidon_cuda.cuh"

#include <iostream>

#include "cuzk_b200.h"

namespace Poseidon {
namespace PoseidonCUDA {

namespace {
const uint64_t *raw(const std::vector<FieldElement> &v) { return reinterpret_cast<const uint64_t *>(v.data()); }
uint64_t *raw(std::vector<FieldElement> &v) { return reinterpret_cast<uint64_t *>(v.data()); }
bool report(const char *what) {
  std::cerr << "CudaPoseidonHash::" << what << ": " << cuzk_last_error() << std::endl;
  return false;
}
}  // namespace

CudaPoseidonHash::CudaPoseidonHash() : initialized_(false), optimal_batch_size_(0), max_batch_size_(0) {
  if (cuzk_init(0) != CUZK_OK) {
    std::cerr << "Failed to initialize CUDA Poseidon: " << cuzk_last_error() << std::endl;
    return;
  }
  cuzk_device_info_t info;
  const int sms = (cuzk_device_info(0, &info) == CUZK_OK) ? info.sm_count : 148;
  optimal_batch_size_ = (size_t)sms * 6 * 128;  // one resident wave: 6 CTAs of 128 one-hash threads per SM
  max_batch_size_ = (size_t)1 << 31;            // the library chunks host batches itself; this is a sanity bound
  initialized_ = true;
}

CudaPoseidonHash::~CudaPoseidonHash() {
  if (initialized_) cuzk_shutdown();
}

bool CudaPoseidonHash::batch_hash_single(const std::vector<FieldElement> &inputs, std::vector<FieldElement> &outputs) {
  if (!initialized_) {
    std::cerr << "CUDA Poseidon not initialized" << std::endl;
    return false;
  }
  if (inputs.empty()) {
    outputs.clear();
    return true;
  }
  // written in place: the library uploads a chunk's inputs before it stores that chunk's outputs, so `outputs` may even be
  // the input vector itself; reusing the caller's capacity avoids a 32 B/hash allocation and page-fault pass per call
  const size_t n = inputs.size();
  const uint64_t *in = raw(inputs);
  outputs.resize(n);
  if (&outputs == &inputs) in = raw(outputs);
  if (cuzk_poseidon_hash_single(in, raw(outputs), n, CUZK_MEM_HOST, nullptr) != CUZK_OK) return report("batch_hash_single");
  return true;
}

bool CudaPoseidonHash::batch_hash_pairs(const std::vector<FieldElement> &left, const std::vector<FieldElement> &right,
                                        std::vector<FieldElement> &outputs) {
  if (!initialized_) {
    std::cerr << "CUDA Poseidon not initialized" << std::endl;
    return false;
  }
  if (left.size() != right.size()) {
    std::cerr << "Left and right input vectors must have the same size" << std::endl;
    return false;
  }
  if (left.empty()) {
    outputs.clear();
    return true;
  }
  const size_t n = left.size();
  outputs.resize(n);  // in place, see batch_hash_single; sizes are equal, so resizing cannot move an aliased input
  if (cuzk_poseidon_hash_pairs(raw(left), raw(right), raw(outputs), n, CUZK_MEM_HOST, nullptr) != CUZK_OK) return report("batch_hash_pairs");
  return true;
}

bool CudaPoseidonHash::batch_permutation(std::vector<std::array<CudaFieldElement, PoseidonParams::STATE_SIZE>> &states) {
  if (!initialized_) {
    std::cerr << "CUDA Poseidon not initialized" << std::endl;
    return false;
  }
  if (states.empty()) return true;
  static_assert(sizeof(std::array<CudaFieldElement, PoseidonParams::STATE_SIZE>) == 96, "states are packed 3 x 32 bytes");
  if (cuzk_poseidon_permutation(reinterpret_cast<uint64_t *>(states.data()), states.size(), CUZK_MEM_HOST, nullptr) != CUZK_OK)
    return report("batch_permutation");
  return true;
}

bool CudaPoseidonHash::batch_sponge(const std::vector<FieldElement> &inputs, size_t width, uint64_t domain_separator,
                                    std::vector<FieldElement> &outputs) {
  if (!initialized_) {
    std::cerr << "CUDA Poseidon not initialized" << std::endl;
    return false;
  }
  if (width == 0 || inputs.size() % width != 0) {
    std::cerr << "batch_sponge: input count must be a multiple of a non-zero width" << std::endl;
    return false;
  }
  const size_t n = inputs.size() / width;
  std::vector<FieldElement> out(n);
  if (n && cuzk_poseidon_sponge(raw(inputs), width, domain_separator, raw(out), n, CUZK_MEM_HOST, nullptr) != CUZK_OK) return report("batch_sponge");
  outputs.swap(out);
  return true;
}

size_t CudaPoseidonHash::get_optimal_batch_size() const { return optimal_batch_size_; }
size_t CudaPoseidonHash::get_max_batch_size() const { return max_batch_size_; }
bool CudaPoseidonHash::is_initialized() const { return initialized_ && cuzk_is_initialized(); }

}  // namespace PoseidonCUDA
}  // namespace Poseidon
