// fr_consts.cuh -- BN254 scalar-field constants shared by the one-thread-per-unit path (fr.cuh) and the cooperative
// sixteen-lanes-per-permutation path (coop.cuh).  Plain C++ (no device code) so the host-side lockstep emulation of the
// cooperative path (tests/cpp/coop_emul.cpp) can include it too.
//   p : src/poseidon/field_arithmetic.cpp:12-14      k = 2^256 mod p : field_arithmetic.cpp:256-258
#pragma once
#include <cstdint>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#endif

namespace cuzk {

typedef uint32_t u32;
typedef uint64_t u64;

// p = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001  (field_arithmetic.cpp:12-14)
#define CUZK_P0 0xf0000001u
#define CUZK_P1 0x43e1f593u
#define CUZK_P2 0x79b97091u
#define CUZK_P3 0x2833e848u
#define CUZK_P4 0x8181585du
#define CUZK_P5 0xb85045b6u
#define CUZK_P6 0xe131a029u
#define CUZK_P7 0x30644e72u
// k = 2^256 mod p = 0x0e0a77c19a07df2f666ea36f7879462e36fc76959f60cd29ac96341c4ffffffb  (:256-258)
#define CUZK_K0 0x4ffffffbu
#define CUZK_K1 0xac96341cu
#define CUZK_K2 0x9f60cd29u
#define CUZK_K3 0x36fc7695u
#define CUZK_K4 0x7879462eu
#define CUZK_K5 0x666ea36fu
#define CUZK_K6 0x9a07df2fu
#define CUZK_K7 0x0e0a77c1u

struct Fr {
  u32 v[8];
};

template <int I> struct PW;   // limbs of 1p
template <> struct PW<0> { static constexpr u32 v = CUZK_P0; };
template <> struct PW<1> { static constexpr u32 v = CUZK_P1; };
template <> struct PW<2> { static constexpr u32 v = CUZK_P2; };
template <> struct PW<3> { static constexpr u32 v = CUZK_P3; };
template <> struct PW<4> { static constexpr u32 v = CUZK_P4; };
template <> struct PW<5> { static constexpr u32 v = CUZK_P5; };
template <> struct PW<6> { static constexpr u32 v = CUZK_P6; };
template <> struct PW<7> { static constexpr u32 v = CUZK_P7; };

// limb i of m*p for m in {1,2,4} (all < 2^256), evaluated at compile time
__host__ __device__ constexpr u32 mulp_limb(int m, int i) {
  const u32 p[8] = {CUZK_P0, CUZK_P1, CUZK_P2, CUZK_P3, CUZK_P4, CUZK_P5, CUZK_P6, CUZK_P7};
  u64 carry = 0;
  u32 out = 0;
  for (int j = 0; j <= i; ++j) {
    u64 t = (u64)p[j] * (u64)m + carry;
    out = (u32)t;
    carry = t >> 32;
  }
  return out;
}

__host__ __device__ constexpr u32 k_limb(int i) {
  const u32 k[8] = {CUZK_K0, CUZK_K1, CUZK_K2, CUZK_K3, CUZK_K4, CUZK_K5, CUZK_K6, CUZK_K7};
  return k[i];
}

// debug builds only: > 0 makes the fast path's "undecided comparison" flags fire on near misses too, to exercise the exact
// fallback in tests (see cond_sub_top in fr.cuh)
#ifndef CUZK_UNC_WIDEN
#define CUZK_UNC_WIDEN 0
#endif

// W - p (W = 2^256) and the quotient-estimate constant of the linear-form MDS row (poseidon.cuh: mds_row_fast)
#define CUZK_NP0 (0u - CUZK_P0)
#define CUZK_NP1 (~CUZK_P1)
#define CUZK_NP2 (~CUZK_P2)
#define CUZK_NP3 (~CUZK_P3)
#define CUZK_NP4 (~CUZK_P4)
#define CUZK_NP5 (~CUZK_P5)
#define CUZK_NP6 (~CUZK_P6)
#define CUZK_NP7 (~CUZK_P7)
__host__ __device__ constexpr u32 np_limb(int i) {   // limbs of W - p (P0 != 0, so no borrow past limb 0)
  const u32 v[8] = {CUZK_NP0, CUZK_NP1, CUZK_NP2, CUZK_NP3, CUZK_NP4, CUZK_NP5, CUZK_NP6, CUZK_NP7};
  return v[i];
}
constexpr u32 kQuotMagic = (u32)((1ull << 61) / (u64)(CUZK_P7 + 1u));   // floor(2^61 / (p_top + 1))


}  // namespace cuzk
