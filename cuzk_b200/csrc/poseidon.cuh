// poseidon.cuh -- the reference's Poseidon permutation (t=3, R_F=8, R_P=56, x^5) on registers.
//
// Follows src/poseidon/poseidon.cpp:60-87 (permutation), :128-134 (add_round_constants),
// :136-146 (S-box), :148-167 (MDS), :103-126 (sponge), bit-exactly, using fr.cuh.
//
// Round constants (poseidon.cpp:33-44) are all < 2^64, so each lives in __constant__ memory as
// two 32-bit words; the library's init kernel generates them with the reference formula and the
// host refuses to start if any upper word is non-zero.  MDS = [[7,23,8],[26,5,4],[15,20,9]]
// (poseidon.cpp:46-58) is compiled in as immediates.
#pragma once
#include "fr.cuh"

namespace cuzk {

constexpr int kRounds = 64;
constexpr int kFullHalf = 4;
constexpr int kPartial = 56;

__constant__ u32 c_rc[kRounds * 3][2];

// s += RC (RC < 2^64), s canonical on entry -> canonical on exit.  add_round_constants :128-134
__device__ __forceinline__ void arc_canon(u32 (&s)[8], int idx) {
  const u32 c0 = c_rc[idx][0], c1 = c_rc[idx][1];
  s[0] = add_cc(s[0], c0);
  s[1] = addc_cc(s[1], c1);
#pragma unroll
  for (int i = 2; i < 7; ++i) s[i] = addc_cc(s[i], 0u);
  s[7] = addc(s[7], 0u);
  cond_sub_mp<1>(s);
}

// same for an arbitrary 256-bit s (first round of a caller-supplied state): wraps mod 2^256, full reduce
__device__ __forceinline__ void arc_general(u32 (&s)[8], int idx) {
  const u32 c0 = c_rc[idx][0], c1 = c_rc[idx][1];
  s[0] = add_cc(s[0], c0);
  s[1] = addc_cc(s[1], c1);
#pragma unroll
  for (int i = 2; i < 7; ++i) s[i] = addc_cc(s[i], 0u);
  s[7] = addc(s[7], 0u);
  fr_reduce(s);
}

// u = multiply(C, s) BEFORE its final reduce, for a small constant C (<= 26) and canonical s:
//   P = C*s (9 words), u = (P mod W + (P div W) * k) mod W          (field_arithmetic.cpp:221-238, :250-330
//   with high <= 4 so mult_high == 0 and high_contribution = high*k unreduced)
template <u32 C>
__device__ __forceinline__ void mds_term(u32 (&u)[8], const u32 (&s)[8]) {
  u32 e[10], o[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) { e[i] = 0; o[i] = 0; }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    mad_wide(e[2 * m], e[2 * m + 1], s[2 * m], C);
    mad_wide(o[2 * m + 1], o[2 * m + 2], s[2 * m + 1], C);
  }
  u32 P[9];
  P[0] = e[0];
  P[1] = add_cc(e[1], o[1]);
#pragma unroll
  for (int i = 2; i < 8; ++i) P[i] = addc_cc(e[i], o[i]);
  P[8] = addc(0u, o[8]);
  const u32 h = P[8];
  // e2 = P_low + h * k_even (carry chained), o2 = h * k_odd
  u32 e2[8], o2[9];
#pragma unroll
  for (int i = 0; i < 8; ++i) e2[i] = P[i];
#pragma unroll
  for (int i = 0; i < 9; ++i) o2[i] = 0;
  e2[0] = mad_lo_cc(h, k_limb(0), e2[0]);
  e2[1] = madc_hi_cc(h, k_limb(0), e2[1]);
#pragma unroll
  for (int m = 1; m < 4; ++m) {
    e2[2 * m] = madc_lo_cc(h, k_limb(2 * m), e2[2 * m]);
    e2[2 * m + 1] = madc_hi_cc(h, k_limb(2 * m), e2[2 * m + 1]);
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) mad_wide(o2[2 * m + 1], o2[2 * m + 2], h, k_limb(2 * m + 1));
  u[0] = e2[0];
  u[1] = add_cc(e2[1], o2[1]);
#pragma unroll
  for (int i = 2; i < 7; ++i) u[i] = addc_cc(e2[i], o2[i]);
  u[7] = addc(e2[7], o2[7]);
}

// n = add(add(add(0, mul(C0,s0)), mul(C1,s1)), mul(C2,s2))   (apply_mds_matrix :148-167)
template <u32 C0, u32 C1, u32 C2>
__device__ __forceinline__ void mds_row(u32 (&n)[8], const u32 (&s0)[8], const u32 (&s1)[8], const u32 (&s2)[8]) {
  u32 a[8], b[8];
  mds_term<C0>(n, s0);
  fr_reduce(n);
  mds_term<C1>(a, s1);
  fr_reduce(a);
  fr_add_canon(b, n, a);
  mds_term<C2>(a, s2);
  fr_reduce(a);
  fr_add_canon(n, b, a);
}

__device__ __forceinline__ void mds(u32 (&s0)[8], u32 (&s1)[8], u32 (&s2)[8]) {
  u32 n0[8], n1[8], n2[8];
  mds_row<7, 23, 8>(n0, s0, s1, s2);
  mds_row<26, 5, 4>(n1, s0, s1, s2);
  mds_row<15, 20, 9>(n2, s0, s1, s2);
#pragma unroll
  for (int i = 0; i < 8; ++i) { s0[i] = n0[i]; s1[i] = n1[i]; s2[i] = n2[i]; }
}

__device__ __forceinline__ void sbox(u32 (&s)[8]) {
  u32 r[8];
  fr_pow5(r, s);
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = r[i];
}

// permutation : poseidon.cpp:60-87.  CANON = every state word is already < p on entry.
template <bool CANON>
__device__ __forceinline__ void permute(u32 (&s0)[8], u32 (&s1)[8], u32 (&s2)[8]) {
#pragma unroll 1
  for (int round = 0; round < kRounds; ++round) {
    if (!CANON && round == 0) {
      arc_general(s0, 0);
      arc_general(s1, 1);
      arc_general(s2, 2);
    } else {
      arc_canon(s0, 3 * round);
      arc_canon(s1, 3 * round + 1);
      arc_canon(s2, 3 * round + 2);
    }
    const bool full = (round < kFullHalf) || (round >= kFullHalf + kPartial);
    const int nsbox = full ? 3 : 1;
    // one S-box body shared by all rounds: apply to s0, and in full rounds rotate the state so
    // the next element sits in s0 (three rotations restore the order).
#pragma unroll 1
    for (int i = 0; i < nsbox; ++i) {
      sbox(s0);
      if (full) {
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          u32 t = s0[w];
          s0[w] = s1[w];
          s1[w] = s2[w];
          s2[w] = t;
        }
      }
    }
    mds(s0, s1, s2);
  }
}

// sponge absorb step: state[i] = add(state[i], x) for a possibly non-canonical x and canonical state
// (poseidon.cpp:113-118).  General add: wraps mod 2^256, full reduce.
__device__ __forceinline__ void absorb(u32 (&s)[8], const u32 (&x)[8]) {
  u32 r[8];
  fr_add_general(r, s, x);
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = r[i];
}

// ---- 32-byte element I/O (AoS of 4 x u64 little-endian limbs == 8 x u32 little-endian words) ----
__device__ __forceinline__ void load_fr(u32 (&x)[8], const uint4 *__restrict__ p) {
  uint4 a = __ldg(p), b = __ldg(p + 1);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
  x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
__device__ __forceinline__ void load_fr_plain(u32 (&x)[8], const uint4 *p) {
  uint4 a = p[0], b = p[1];
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
  x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
__device__ __forceinline__ void store_fr(uint4 *p, const u32 (&x)[8]) {
  p[0] = make_uint4(x[0], x[1], x[2], x[3]);
  p[1] = make_uint4(x[4], x[5], x[6], x[7]);
}
__device__ __forceinline__ void set_small(u32 (&x)[8], u32 v) {
  x[0] = v;
#pragma unroll
  for (int i = 1; i < 8; ++i) x[i] = 0;
}

// hash_multiple / device sponge over `width` children (poseidon.cpp:98-126; domain separator DS):
// absorbs two per permutation, a final odd child alone; width == 0 -> zero permutations -> output 0.
template <class Loader>
__device__ __forceinline__ void sponge_n(u32 (&out)[8], u32 ds, int width, Loader load) {
  u32 s0[8], s1[8], s2[8];
  set_small(s0, ds);
  set_small(s1, 0);
  set_small(s2, 0);
#pragma unroll 1
  for (int i = 0; i < width; i += 2) {
    u32 x[8];
    load(x, i);
    absorb(s1, x);
    if (i + 1 < width) {
      load(x, i + 1);
      absorb(s2, x);
    }
    permute<true>(s0, s1, s2);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) out[i] = s1[i];
}

}  // namespace cuzk
