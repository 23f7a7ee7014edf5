// cuda/cuda_field_element.cuh -- CudaFieldElement, the 32-byte-aligned twin of FieldElement.
//
// Replaces the reference's src/poseidon/cuda/cuda_field_element.cuh:13-113 (the struct) for HOST code.  The
// reference's __device__ arithmetic in the same file (:119-468) has no counterpart here: device arithmetic lives in
// libcuzk_b200.so (csrc/fr.cuh) behind the C ABI and is bit-exact with the reference's CPU arithmetic, which the
// reference's own device functions are not.  This header is plain C++; it also compiles under
// nvcc, where the members are usable from device code.
#pragma once

#include <cstdint>
#include <string>

#include "../field_arithmetic.hpp"

#if defined(__CUDACC__)
#define CUZK_HD __host__ __device__
#else
#define CUZK_HD
#endif

namespace Poseidon {

struct alignas(32) CudaFieldElement {
  uint64_t limbs[4];

  CUZK_HD CudaFieldElement() : limbs{0, 0, 0, 0} {}
  CUZK_HD explicit CudaFieldElement(uint64_t value) : limbs{value, 0, 0, 0} {}
  CUZK_HD CudaFieldElement(uint64_t v0, uint64_t v1, uint64_t v2, uint64_t v3) : limbs{v0, v1, v2, v3} {}
  CudaFieldElement(const CudaFieldElement &) = default;
  CudaFieldElement &operator=(const CudaFieldElement &) = default;

  CUZK_HD bool operator==(const CudaFieldElement &o) const {
    return ((limbs[0] ^ o.limbs[0]) | (limbs[1] ^ o.limbs[1]) | (limbs[2] ^ o.limbs[2]) | (limbs[3] ^ o.limbs[3])) == 0;
  }
  CUZK_HD bool operator!=(const CudaFieldElement &o) const { return !(*this == o); }
  CUZK_HD bool operator<(const CudaFieldElement &o) const {
    for (int i = 3; i >= 0; --i)
      if (limbs[i] != o.limbs[i]) return limbs[i] < o.limbs[i];
    return false;
  }
  CUZK_HD bool is_zero() const { return (limbs[0] | limbs[1] | limbs[2] | limbs[3]) == 0; }
  CUZK_HD void set_zero() { limbs[0] = limbs[1] = limbs[2] = limbs[3] = 0; }

  // host-side conversions to and from the CPU type (same limb order, same plain canonical form)
  explicit CudaFieldElement(const FieldElement &fe) : limbs{fe.limbs[0], fe.limbs[1], fe.limbs[2], fe.limbs[3]} {}
  operator FieldElement() const { return FieldElement(limbs[0], limbs[1], limbs[2], limbs[3]); }

#ifndef __CUDA_ARCH__
  std::string to_hex() const { return static_cast<FieldElement>(*this).to_hex(); }
  std::string to_dec() const { return static_cast<FieldElement>(*this).to_dec(); }
  static CudaFieldElement from_hex(const std::string &hex) { return CudaFieldElement(FieldElement::from_hex(hex)); }
  // uses the CPU library's FieldElement::random(); a template so that it is only looked up where a caller asks for it
  template <class FE = FieldElement>
  static CudaFieldElement random() { return CudaFieldElement(FE::random()); }
#endif
};

static_assert(sizeof(CudaFieldElement) == 32 && sizeof(FieldElement) == 32, "field elements are 4 x u64");

namespace CudaFieldOps {}  // device arithmetic namespace of the reference; empty here (see the note at the top)

}  // namespace Poseidon
