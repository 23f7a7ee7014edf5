// fr.cuh -- BN254 scalar-field "reference arithmetic" on 8 x 32-bit limbs for sm_100a.
//
// Every function here evaluates, bit for bit, what the reference CPU code computes
// (src/poseidon/field_arithmetic.cpp), including its non-modular corner semantics:
//   add      :172-182   (a + b) mod 2^256, then reduce            -> fr_add_general / fr_add_canon
//   subtract :184-219   conditional +p, limb-wise borrow quirk     -> fr_sub_ref
//   reduce   :244-248   while (a >= p) a -= p  (quotient <= 5)     -> fr_reduce
//   multiply :221-238 + reduce_512 :250-330                        -> fr_mul / fr_sqr / fr_pow5
// The exact formulas are restated in DESIGN.md ("Arithmetic the kernels evaluate").
//
// Implementation: 32x32->64 multiply-adds issue as IMAD.WIDE.U32(.X) -- ptxas fuses each
// mad.lo.cc / madc.hi.cc pair below into one IMAD.WIDE with predicate carry in/out.  Products
// are accumulated in two interleaved sets of 64-bit lanes (even / odd 32-bit positions), so a
// row of four products is one carry chain with no per-product carry fix-up.
#pragma once
#include <cstdint>

#include "fr_consts.cuh"

#ifndef CUZK_KARATSUBA
#define CUZK_KARATSUBA 0   // 1: one level of Karatsuba for the two full 8x8 products of a multiply (see mul_wide_by_k).
                           // Bit-exact and 8 % fewer multiplier-pipe cycles, but measured SLOWER on B200 (176.9 vs 185.8 M
                           // hashes/s): the extra serial carry chains cost more issue slots and latency than the pipe time saved.
#endif

namespace cuzk {

// ---- single-instruction carry-chain primitives (CGBN style; volatile keeps PTX order) ----
__device__ __forceinline__ u32 add_cc(u32 a, u32 b) { u32 r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ u32 addc_cc(u32 a, u32 b) { u32 r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ u32 addc(u32 a, u32 b) { u32 r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ u32 sub_cc(u32 a, u32 b) { u32 r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ u32 subc_cc(u32 a, u32 b) { u32 r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ u32 subc(u32 a, u32 b) { u32 r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ u32 mad_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ u32 madc_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ u32 madc_hi_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ u32 madc_hi(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ u32 madc_lo(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ u32 mad_lo(u32 a, u32 b, u32 c) { return a * b + c; }

// (hi:lo) = a*b + (hi:lo) as one IMAD.WIDE.U32 (no carry in/out)
__device__ __forceinline__ void mad_wide(u32 &lo, u32 &hi, u32 a, u32 b) {
  u64 t = (u64)a * (u64)b + (((u64)hi << 32) | lo);
  lo = (u32)t;
  hi = (u32)(t >> 32);
}

// ---- comparisons / conditional subtract ----
// x -= m*p if x >= m*p   (m in {1,2,4}); 8 IADD3.X + 8 SEL
template <int M>
__device__ __forceinline__ void cond_sub_mp(u32 (&x)[8]) {
  u32 d[8];
  d[0] = sub_cc(x[0], mulp_limb(M, 0));
#pragma unroll
  for (int i = 1; i < 8; ++i) d[i] = subc_cc(x[i], mulp_limb(M, i));
  u32 borrow = subc(0u, 0u);  // 0xffffffff when x < m*p
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = borrow ? x[i] : d[i];
}

// reduce : field_arithmetic.cpp:244-248.  Exact x mod p for ANY 256-bit x (quotient <= 5).
__device__ __forceinline__ void fr_reduce(u32 (&x)[8]) {
  cond_sub_mp<4>(x);
  cond_sub_mp<2>(x);
  cond_sub_mp<1>(x);
}

// ---- fast path: conditional subtract decided by the top word alone ------------------------------------------
// x -= M*p when x[7] > top word of M*p, which proves x >= M*p; x[7] < top proves x < M*p.  The undecided case
// x[7] == top (2^-32 per call on random data) is reported in `unc`: the caller then recomputes its whole unit (hash,
// node, proof level) on the exact path, so results stay bit-exact.  One ISETP + eight predicated IADD3.X instead of
// eight IADD3.X + eight SEL.  `enable` == 0 turns the subtraction off (used for the reference's "mh == 0" case).
// CUZK_UNC_WIDEN > 0 (debug builds only) flags near-misses too, to exercise the exact fallback in tests.
template <int M>
__device__ __forceinline__ void cond_sub_top(u32 (&x)[8], u32 &unc, u32 enable) {
  constexpr u32 T = mulp_limb(M, 7);
  unc |= (((x[7] ^ T) >> CUZK_UNC_WIDEN) == 0u) ? 1u : 0u;
  asm volatile(
      "{\n\t"
      ".reg .pred e, p;\n\t"
      "setp.ne.u32 e, %16, 0;\n\t"
      "setp.gt.and.u32 p, %7, %15, e;\n\t"
      "@p sub.cc.u32 %0, %0, %8;\n\t"
      "@p subc.cc.u32 %1, %1, %9;\n\t"
      "@p subc.cc.u32 %2, %2, %10;\n\t"
      "@p subc.cc.u32 %3, %3, %11;\n\t"
      "@p subc.cc.u32 %4, %4, %12;\n\t"
      "@p subc.cc.u32 %5, %5, %13;\n\t"
      "@p subc.cc.u32 %6, %6, %14;\n\t"
      "@p subc.u32 %7, %7, %15;\n\t"
      "}"
      : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7])
      : "n"(mulp_limb(M, 0)), "n"(mulp_limb(M, 1)), "n"(mulp_limb(M, 2)), "n"(mulp_limb(M, 3)), "n"(mulp_limb(M, 4)),
        "n"(mulp_limb(M, 5)), "n"(mulp_limb(M, 6)), "n"(mulp_limb(M, 7)), "r"(enable));
}
template <int M>
__device__ __forceinline__ void cond_sub_top(u32 (&x)[8], u32 &unc) {
  constexpr u32 T = mulp_limb(M, 7);
  unc |= (((x[7] ^ T) >> CUZK_UNC_WIDEN) == 0u) ? 1u : 0u;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.gt.u32 p, %7, %15;\n\t"
      "@p sub.cc.u32 %0, %0, %8;\n\t"
      "@p subc.cc.u32 %1, %1, %9;\n\t"
      "@p subc.cc.u32 %2, %2, %10;\n\t"
      "@p subc.cc.u32 %3, %3, %11;\n\t"
      "@p subc.cc.u32 %4, %4, %12;\n\t"
      "@p subc.cc.u32 %5, %5, %13;\n\t"
      "@p subc.cc.u32 %6, %6, %14;\n\t"
      "@p subc.u32 %7, %7, %15;\n\t"
      "}"
      : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7])
      : "n"(mulp_limb(M, 0)), "n"(mulp_limb(M, 1)), "n"(mulp_limb(M, 2)), "n"(mulp_limb(M, 3)), "n"(mulp_limb(M, 4)),
        "n"(mulp_limb(M, 5)), "n"(mulp_limb(M, 6)), "n"(mulp_limb(M, 7)));
}

// reduce for ANY 256-bit x on the fast path (quotient <= 5: 4p, 2p, p ladder)
__device__ __forceinline__ void fr_reduce_fast(u32 (&x)[8], u32 &unc) {
  cond_sub_top<4>(x, unc);
  cond_sub_top<2>(x, unc);
  cond_sub_top<1>(x, unc);
}
__device__ __forceinline__ void fr_reduce_fast(u32 (&x)[8], u32 &unc, u32 enable) {
  cond_sub_top<4>(x, unc, enable);
  cond_sub_top<2>(x, unc, enable);
  cond_sub_top<1>(x, unc, enable);
}

// ---- optional: reduce through a shared-memory table of the multiples 0p .. 5p -------------------------------------
// One quotient estimate from the top word (est in {q-1, q}), one subtraction of est*p fetched with two LDS.128, one
// top-word conditional subtract: ~21 ALU + 2 LDS instead of the three-step ladder's ~33 ALU.  Kernels call
// fr_table_init() before their bounds check (it contains a __syncthreads).
// Bit-exact (parity suite green) and 8 % fewer ALU instructions per S-box, but no faster on B200 (184.5 M hashes/s either
// way), so it is off by default.
#ifndef CUZK_REDUCE_TABLE
#define CUZK_REDUCE_TABLE 0
#endif
static __constant__ u32 c_mulp[6][8] = {
    {mulp_limb(0, 0), mulp_limb(0, 1), mulp_limb(0, 2), mulp_limb(0, 3), mulp_limb(0, 4), mulp_limb(0, 5), mulp_limb(0, 6), mulp_limb(0, 7)},
    {mulp_limb(1, 0), mulp_limb(1, 1), mulp_limb(1, 2), mulp_limb(1, 3), mulp_limb(1, 4), mulp_limb(1, 5), mulp_limb(1, 6), mulp_limb(1, 7)},
    {mulp_limb(2, 0), mulp_limb(2, 1), mulp_limb(2, 2), mulp_limb(2, 3), mulp_limb(2, 4), mulp_limb(2, 5), mulp_limb(2, 6), mulp_limb(2, 7)},
    {mulp_limb(3, 0), mulp_limb(3, 1), mulp_limb(3, 2), mulp_limb(3, 3), mulp_limb(3, 4), mulp_limb(3, 5), mulp_limb(3, 6), mulp_limb(3, 7)},
    {mulp_limb(4, 0), mulp_limb(4, 1), mulp_limb(4, 2), mulp_limb(4, 3), mulp_limb(4, 4), mulp_limb(4, 5), mulp_limb(4, 6), mulp_limb(4, 7)},
    {mulp_limb(5, 0), mulp_limb(5, 1), mulp_limb(5, 2), mulp_limb(5, 3), mulp_limb(5, 4), mulp_limb(5, 5), mulp_limb(5, 6), mulp_limb(5, 7)}};
__device__ __forceinline__ uint4 *fr_table() {
  __shared__ uint4 s_mulp[12];   // row m = m*p as two uint4
  return s_mulp;
}
__device__ __forceinline__ void fr_table_init() {
#if CUZK_REDUCE_TABLE
  for (unsigned w = threadIdx.x; w < 48; w += blockDim.x) reinterpret_cast<u32 *>(fr_table())[w] = c_mulp[w >> 3][w & 7];
  __syncthreads();
#endif
}
constexpr u32 kTopQuotMagic = (u32)((1ull << 61) / (u64)(CUZK_P7 + 1u));   // floor(2^61 / (p_top + 1))
__device__ __forceinline__ void fr_reduce_table(u32 (&x)[8], u32 &unc, u32 enable) {
  u32 est = __umulhi(x[7], kTopQuotMagic) >> 29;     // floor(x7 / (p7 + 1)) up to rounding: q - 1 <= est <= q, est <= 5
  est = enable ? est : 0u;
  const uint4 *row = fr_table() + 2 * est;
  const uint4 a = row[0], b = row[1];
  x[0] = sub_cc(x[0], a.x);
  x[1] = subc_cc(x[1], a.y);
  x[2] = subc_cc(x[2], a.z);
  x[3] = subc_cc(x[3], a.w);
  x[4] = subc_cc(x[4], b.x);
  x[5] = subc_cc(x[5], b.y);
  x[6] = subc_cc(x[6], b.z);
  x[7] = subc(x[7], b.w);
  cond_sub_top<1>(x, unc, enable);
}

// add for canonical operands (a, b < p): sum < 2p < 2^256, at most one subtraction.
__device__ __forceinline__ void fr_add_canon(u32 (&r)[8], const u32 (&a)[8], const u32 (&b)[8]) {
  r[0] = add_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < 7; ++i) r[i] = addc_cc(a[i], b[i]);
  r[7] = addc(a[7], b[7]);
  cond_sub_mp<1>(r);
}

// add : field_arithmetic.cpp:172-182 for arbitrary 256-bit operands (carry out dropped, full reduce)
__device__ __forceinline__ void fr_add_general(u32 (&r)[8], const u32 (&a)[8], const u32 (&b)[8]) {
  r[0] = add_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < 7; ++i) r[i] = addc_cc(a[i], b[i]);
  r[7] = addc(a[7], b[7]);
  fr_reduce(r);
}

// subtract : field_arithmetic.cpp:184-219 restated on 64-bit limb semantics.
// Works on u64 limbs because the reference's lost-borrow quirk is defined per 64-bit limb.
__device__ __forceinline__ void fr_sub_ref(u64 (&r)[4], const u64 (&a)[4], const u64 (&b)[4]) {
  const u64 p[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
  bool lt = false, decided = false;  // a < b, most-significant limb first (:60-68)
#pragma unroll
  for (int i = 3; i >= 0; --i) {
    if (!decided && a[i] != b[i]) { lt = a[i] < b[i]; decided = true; }
  }
  u64 x[4];
  if (lt) {  // a + p mod 2^256 (:190-197)
    u64 carry = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      u64 s = a[i] + p[i];
      u64 c1 = s < a[i];
      u64 s2 = s + carry;
      u64 c2 = s2 < s;
      x[i] = s2;
      carry = c1 | c2;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = a[i];
  }
  u64 borrow = 0;  // subtract_internal (:204-219): b[i] + borrow wraps in 64 bits
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    u64 y = b[i] + borrow;
    if (x[i] >= y) { r[i] = x[i] - y; borrow = 0; }
    else { r[i] = x[i] - y; borrow = 1; }   // == x + (2^64-1 - y) + 1 mod 2^64
  }
}

// ---- schoolbook products on even/odd 64-bit lanes ----
// One row chain: acc[0..2*N-1] += {a0,a1,..} * b at lane stride 2, carry chained; returns nothing,
// the carry out of the last lane is left in CC for the caller (addc) when needed.
// NW = number of full (wide) lanes, LO = 1 when a final low-half-only product follows.
template <int NW, int LO>
__device__ __forceinline__ void row_chain(u32 *acc, const u32 *a /* stride 2 */, u32 b) {
  if (NW > 0) {
    acc[0] = mad_lo_cc(a[0], b, acc[0]);
    acc[1] = madc_hi_cc(a[0], b, acc[1]);
#pragma unroll
    for (int j = 1; j < NW; ++j) {
      acc[2 * j] = madc_lo_cc(a[2 * j], b, acc[2 * j]);
      acc[2 * j + 1] = madc_hi_cc(a[2 * j], b, acc[2 * j + 1]);
    }
    if (LO) acc[2 * NW] = madc_lo(a[2 * NW], b, acc[2 * NW]);
  } else if (LO) {
    acc[0] = mad_lo(a[0], b, acc[0]);
  }
}

// r[0..15] = a * b  (full 512-bit product)
__device__ __forceinline__ void mul_wide_8x8(u32 (&r)[16], const u32 (&a)[8], const u32 (&b)[8]) {
  u32 e[16], o[16];  // e[w]: word w from even-position products; o[w]: word w from odd-position products
#pragma unroll
  for (int i = 0; i < 16; ++i) { e[i] = 0; o[i] = 0; }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if ((i & 1) == 0) {
      row_chain<4, 0>(e + i, a, b[i]);            // a_even * b_i -> even positions i..i+6
      if (i + 8 < 16) e[i + 8] = addc(0u, 0u);    // carry lands in a so-far untouched word
      row_chain<4, 0>(o + i + 1, a + 1, b[i]);    // a_odd * b_i -> odd positions i+1..i+7 (no carry out)
    } else {
      row_chain<4, 0>(o + i, a, b[i]);
      if (i + 8 < 16) o[i + 8] = addc(0u, 0u);
      row_chain<4, 0>(e + i + 1, a + 1, b[i]);
    }
  }
  r[0] = e[0];
  r[1] = add_cc(e[1], o[1]);
#pragma unroll
  for (int i = 2; i < 15; ++i) r[i] = addc_cc(e[i], o[i]);
  r[15] = addc(e[15], o[15]);
}

// r[0..7] = a[0..3] * b[0..3] on even/odd lanes (16 multiply-adds)
__device__ __forceinline__ void mul_wide_4x4(u32 (&r)[8], const u32 *a, const u32 *b) {
  u32 e[8], o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { e[i] = 0; o[i] = 0; }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if ((i & 1) == 0) {
      row_chain<2, 0>(e + i, a, b[i]);            // a0, a2 -> even positions i, i+2
      if (i + 4 < 8) e[i + 4] = addc(0u, 0u);
      row_chain<2, 0>(o + i + 1, a + 1, b[i]);    // a1, a3 -> odd positions i+1, i+3
    } else {
      row_chain<2, 0>(o + i, a, b[i]);
      if (i + 4 < 8) o[i + 4] = addc(0u, 0u);
      row_chain<2, 0>(e + i + 1, a + 1, b[i]);
    }
  }
  r[0] = e[0];
  r[1] = add_cc(e[1], o[1]);
#pragma unroll
  for (int i = 2; i < 7; ++i) r[i] = addc_cc(e[i], o[i]);
  r[7] = addc(e[7], o[7]);
}

// ---- one level of Karatsuba over the 128-bit halves: three 4x4 products (48 multiply-adds) instead of 64 -------------
// The 32x32->64 multiplier is the kernel's bound and the ALU has slack, so ~45 extra carry-chain adds per product are a
// good trade for 16 multiply-adds.  Exact for any operands:
//   a = a0 + a1 W2, b = b0 + b1 W2 (W2 = 2^128):  a b = z0 + (zm - z0 - z2) W2 + z2 W2^2,
//   z0 = a0 b0, z2 = a1 b1, zm = (a0 + a1)(b0 + b1) = as bs + ca bs W2 + cb as W2 + ca cb W2^2   (ca, cb = the sums' carry bits)
#define CUZK_KS0 0xc8794629u   // k0 + k1 for k = 2^256 mod p (no carry out: 0x4506ee573968ac591304d78bc8794629)
#define CUZK_KS1 0x1304d78bu
#define CUZK_KS2 0x3968ac59u
#define CUZK_KS3 0x4506ee57u

// shared tail: r = z0 + (zm9 - z0 - z2) * 2^128 + z2 * 2^256, zm9 = zm[0..7] + zm8 * 2^256
__device__ __forceinline__ void karatsuba_combine(u32 (&r)[16], const u32 (&z0)[8], const u32 (&z2)[8], u32 (&zm)[8], u32 zm8) {
  u32 s[9];
  s[0] = add_cc(z0[0], z2[0]);
#pragma unroll
  for (int i = 1; i < 8; ++i) s[i] = addc_cc(z0[i], z2[i]);
  s[8] = addc(0u, 0u);
  u32 z1[9];
  z1[0] = sub_cc(zm[0], s[0]);
#pragma unroll
  for (int i = 1; i < 8; ++i) z1[i] = subc_cc(zm[i], s[i]);
  z1[8] = subc(zm8, s[8]);
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = z0[i];
  r[4] = add_cc(z0[4], z1[0]);
#pragma unroll
  for (int i = 1; i < 4; ++i) r[4 + i] = addc_cc(z0[4 + i], z1[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) r[8 + i] = addc_cc(z2[i], z1[4 + i]);
  r[12] = addc_cc(z2[4], z1[8]);
  r[13] = addc_cc(z2[5], 0u);
  r[14] = addc_cc(z2[6], 0u);
  r[15] = addc(z2[7], 0u);
}

// r[0..15] = a * k for the constant k = 2^256 mod p
__device__ __forceinline__ void mul_wide_by_k(u32 (&r)[16], const u32 (&a)[8]) {
  const u32 klo[4] = {CUZK_K0, CUZK_K1, CUZK_K2, CUZK_K3}, khi[4] = {CUZK_K4, CUZK_K5, CUZK_K6, CUZK_K7};
  const u32 ks[4] = {CUZK_KS0, CUZK_KS1, CUZK_KS2, CUZK_KS3};
  u32 z0[8], z2[8], zm[8], as[4];
  mul_wide_4x4(z0, a, klo);
  mul_wide_4x4(z2, a + 4, khi);
  as[0] = add_cc(a[0], a[4]);
  as[1] = addc_cc(a[1], a[5]);
  as[2] = addc_cc(a[2], a[6]);
  as[3] = addc_cc(a[3], a[7]);
  const u32 ca = addc(0u, 0u);
  mul_wide_4x4(zm, as, ks);
  // zm += ca * ks * 2^128  (ks has no carry bit of its own)
  u32 zm8;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.u32 p, %5, 0;\n\t"
      "mov.u32 %4, 0;\n\t"
      "@p add.cc.u32 %0, %0, %6;\n\t"
      "@p addc.cc.u32 %1, %1, %7;\n\t"
      "@p addc.cc.u32 %2, %2, %8;\n\t"
      "@p addc.cc.u32 %3, %3, %9;\n\t"
      "@p addc.u32 %4, %4, 0;\n\t"
      "}"
      : "+r"(zm[4]), "+r"(zm[5]), "+r"(zm[6]), "+r"(zm[7]), "=r"(zm8)
      : "r"(ca), "n"(CUZK_KS0), "n"(CUZK_KS1), "n"(CUZK_KS2), "n"(CUZK_KS3));
  karatsuba_combine(r, z0, z2, zm, zm8);
}

// r[0..15] = a * b for two register operands
__device__ __forceinline__ void mul_wide_8x8_karatsuba(u32 (&r)[16], const u32 (&a)[8], const u32 (&b)[8]) {
  u32 z0[8], z2[8], zm[8], as[4], bs[4];
  mul_wide_4x4(z0, a, b);
  mul_wide_4x4(z2, a + 4, b + 4);
  as[0] = add_cc(a[0], a[4]);
  as[1] = addc_cc(a[1], a[5]);
  as[2] = addc_cc(a[2], a[6]);
  as[3] = addc_cc(a[3], a[7]);
  const u32 ca = addc(0u, 0u);
  bs[0] = add_cc(b[0], b[4]);
  bs[1] = addc_cc(b[1], b[5]);
  bs[2] = addc_cc(b[2], b[6]);
  bs[3] = addc_cc(b[3], b[7]);
  const u32 cb = addc(0u, 0u);
  mul_wide_4x4(zm, as, bs);
  // zm9 = zm + (ca ? bs : 0) * 2^128 + (cb ? as : 0) * 2^128 + (ca & cb) * 2^256
  const u32 ma = 0u - ca, mb = 0u - cb;
  u32 zm8 = ca & cb;
  zm[4] = add_cc(zm[4], bs[0] & ma);
  zm[5] = addc_cc(zm[5], bs[1] & ma);
  zm[6] = addc_cc(zm[6], bs[2] & ma);
  zm[7] = addc_cc(zm[7], bs[3] & ma);
  zm8 = addc(zm8, 0u);
  zm[4] = add_cc(zm[4], as[0] & mb);
  zm[5] = addc_cc(zm[5], as[1] & mb);
  zm[6] = addc_cc(zm[6], as[2] & mb);
  zm[7] = addc_cc(zm[7], as[3] & mb);
  zm8 = addc(zm8, 0u);
  karatsuba_combine(r, z0, z2, zm, zm8);
}

// t[0..7] = (t + a * b) mod 2^256   (low half only; 28 wide + 8 low products)
__device__ __forceinline__ void mad_low_8x8(u32 (&t)[8], const u32 (&a)[8], const u32 (&b)[8]) {
  u32 e[8], o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { e[i] = t[i]; o[i] = 0; }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    // positions i+j <= 7.  a_even: j = 0,2,4,6 ; a_odd: j = 1,3,5,7
    if ((i & 1) == 0) {
      // even positions i+j (j even): wide while i+j <= 6.
      constexpr int dummy = 0; (void)dummy;
      switch (i) {
        case 0: row_chain<4, 0>(e + 0, a, b[0]); row_chain<3, 1>(o + 1, a + 1, b[0]); break;   // odd pos 1,3,5 wide, 7 lo
        case 2: row_chain<3, 0>(e + 2, a, b[2]); row_chain<2, 1>(o + 3, a + 1, b[2]); break;   // even 2,4,6 ; odd 3,5 wide, 7 lo
        case 4: row_chain<2, 0>(e + 4, a, b[4]); row_chain<1, 1>(o + 5, a + 1, b[4]); break;
        case 6: row_chain<1, 0>(e + 6, a, b[6]); row_chain<0, 1>(o + 7, a + 1, b[6]); break;
      }
    } else {
      switch (i) {
        case 1: row_chain<3, 1>(o + 1, a, b[1]); row_chain<3, 0>(e + 2, a + 1, b[1]); break;   // odd pos 1,3,5 wide,7 lo ; even 2,4,6
        case 3: row_chain<2, 1>(o + 3, a, b[3]); row_chain<2, 0>(e + 4, a + 1, b[3]); break;
        case 5: row_chain<1, 1>(o + 5, a, b[5]); row_chain<1, 0>(e + 6, a + 1, b[5]); break;
        case 7: row_chain<0, 1>(o + 7, a, b[7]); break;
      }
    }
  }
  t[0] = e[0];
  t[1] = add_cc(e[1], o[1]);
#pragma unroll
  for (int i = 2; i < 7; ++i) t[i] = addc_cc(e[i], o[i]);
  t[7] = addc(e[7], o[7]);
}

// reduce_512 : field_arithmetic.cpp:250-330 on a 16-word product.
//   Mh = high*k ; t = (Mh_lo + (Mh_hi*k mod W)) mod W ; hc = Mh_hi != 0 ? reduce(t) : t ;
//   r = reduce((low + hc) mod W)
// EXACT = false uses the top-word reductions and reports undecided comparisons in `unc` (see cond_sub_top).
template <bool EXACT>
__device__ __forceinline__ void fr_reduce_512(u32 (&r)[8], const u32 (&prod)[16], u32 &unc) {
  const u32 kk[8] = {CUZK_K0, CUZK_K1, CUZK_K2, CUZK_K3, CUZK_K4, CUZK_K5, CUZK_K6, CUZK_K7};
  u32 high[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) high[i] = prod[8 + i];
  u32 m1[16];
#if CUZK_KARATSUBA
  if (EXACT) mul_wide_8x8(m1, high, kk);   // the exact path stays on the schoolbook product: an independent evaluation
  else mul_wide_by_k(m1, high);
#else
  mul_wide_8x8(m1, high, kk);
#endif
  u32 t[8], mh[8];
  u32 mh_or = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { t[i] = m1[i]; mh[i] = m1[8 + i]; mh_or |= m1[8 + i]; }
  mad_low_8x8(t, mh, kk);     // adds nothing when mh == 0, matching the reference's skip (:303)
  if (EXACT) {
    if (mh_or != 0) fr_reduce(t);
  } else {
#if CUZK_REDUCE_TABLE
    fr_reduce_table(t, unc, mh_or);
#else
    fr_reduce_fast(t, unc, mh_or);
#endif
  }
  r[0] = add_cc(prod[0], t[0]);
#pragma unroll
  for (int i = 1; i < 7; ++i) r[i] = addc_cc(prod[i], t[i]);
  r[7] = addc(prod[7], t[7]);
  if (EXACT) fr_reduce(r);
#if CUZK_REDUCE_TABLE
  else fr_reduce_table(r, unc, 1u);
#else
  else fr_reduce_fast(r, unc);
#endif
}

// multiply : field_arithmetic.cpp:221-238 (valid for arbitrary 256-bit operands)
template <bool EXACT>
__device__ __forceinline__ void fr_mul_t(u32 (&r)[8], const u32 (&a)[8], const u32 (&b)[8], u32 &unc) {
  u32 prod[16];
#if CUZK_KARATSUBA
  if (EXACT) mul_wide_8x8(prod, a, b);
  else mul_wide_8x8_karatsuba(prod, a, b);
#else
  mul_wide_8x8(prod, a, b);
#endif
  fr_reduce_512<EXACT>(r, prod, unc);
}
__device__ __forceinline__ void fr_mul(u32 (&r)[8], const u32 (&a)[8], const u32 (&b)[8]) {
  u32 unused = 0;
  fr_mul_t<true>(r, a, b, unused);
}

// r[0..15] = a * a using the symmetry of the product: the 28 off-diagonal products are accumulated once
// on the even/odd lanes, doubled, and the 8 diagonal squares are added by one carry-chained row
// (36 IMAD.WIDE instead of 64).  Exact for any 256-bit a.
__device__ __forceinline__ void sqr_wide_8(u32 (&r)[16], const u32 (&a)[8]) {
  u32 e[16], o[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { e[i] = 0; o[i] = 0; }
  // row i: a_i * a_j, j > i.  j - i odd -> odd position 2i+1+2t (accumulator o); j - i even -> even position (e).
  // After row i every accumulator is < 2^(32(i+9)): the chain whose top lane ends at word i+8 has no carry out,
  // the other one carries into word i+8 and stops there.
  row_chain<4, 0>(o + 1, a + 1, a[0]);   // (0,1)(0,3)(0,5)(0,7): positions 1,3,5,7 -> words 1..8
  row_chain<3, 0>(e + 2, a + 2, a[0]);   // (0,2)(0,4)(0,6): positions 2,4,6 -> words 2..7
  e[8] = addc(e[8], 0u);
  row_chain<3, 0>(o + 3, a + 2, a[1]);   // (1,2)(1,4)(1,6): positions 3,5,7 -> words 3..8
  o[9] = addc(o[9], 0u);
  row_chain<3, 0>(e + 4, a + 3, a[1]);   // (1,3)(1,5)(1,7): positions 4,6,8 -> words 4..9
  row_chain<3, 0>(o + 5, a + 3, a[2]);   // (2,3)(2,5)(2,7): positions 5,7,9 -> words 5..10
  row_chain<2, 0>(e + 6, a + 4, a[2]);   // (2,4)(2,6): positions 6,8 -> words 6..9
  e[10] = addc(e[10], 0u);
  row_chain<2, 0>(o + 7, a + 4, a[3]);   // (3,4)(3,6): positions 7,9 -> words 7..10
  o[11] = addc(o[11], 0u);
  row_chain<2, 0>(e + 8, a + 5, a[3]);   // (3,5)(3,7): positions 8,10 -> words 8..11
  row_chain<2, 0>(o + 9, a + 5, a[4]);   // (4,5)(4,7): positions 9,11 -> words 9..12
  row_chain<1, 0>(e + 10, a + 6, a[4]);  // (4,6): position 10 -> words 10,11
  e[12] = addc(e[12], 0u);
  row_chain<1, 0>(o + 11, a + 6, a[5]);  // (5,6): position 11 -> words 11,12
  o[13] = addc(o[13], 0u);
  row_chain<1, 0>(e + 12, a + 7, a[5]);  // (5,7): position 12 -> words 12,13
  row_chain<1, 0>(o + 13, a + 7, a[6]);  // (6,7): position 13 -> words 13,14
  // d = e + o (off-diagonal sum, < 2^511)
  u32 d[16];
  d[0] = 0;
  d[1] = o[1];
  d[2] = add_cc(e[2], o[2]);
#pragma unroll
  for (int i = 3; i < 15; ++i) d[i] = addc_cc(e[i], o[i]);
  d[15] = addc(e[15], o[15]);
  // doubled, then the diagonal a_i^2 at position 2i as one carry-chained row of eight lanes
#pragma unroll
  for (int i = 15; i >= 1; --i) r[i] = __funnelshift_l(d[i - 1], d[i], 1);
  r[0] = 0;
  r[0] = mad_lo_cc(a[0], a[0], r[0]);
  r[1] = madc_hi_cc(a[0], a[0], r[1]);
#pragma unroll
  for (int i = 1; i < 8; ++i) {
    r[2 * i] = madc_lo_cc(a[i], a[i], r[2 * i]);
    r[2 * i + 1] = (i == 7) ? madc_hi(a[i], a[i], r[2 * i + 1]) : madc_hi_cc(a[i], a[i], r[2 * i + 1]);
  }
}

// square : field_arithmetic.cpp:240-242 (= multiply(a, a), evaluated with the symmetric product)
template <bool EXACT>
__device__ __forceinline__ void fr_sqr_t(u32 (&r)[8], const u32 (&a)[8], u32 &unc) {
  u32 prod[16];
  sqr_wide_8(prod, a);
  fr_reduce_512<EXACT>(r, prod, unc);
}
__device__ __forceinline__ void fr_sqr(u32 (&r)[8], const u32 (&a)[8]) {
  u32 unused = 0;
  fr_sqr_t<true>(r, a, unused);
}

// power5 : field_arithmetic.cpp:332-338
template <bool EXACT>
__device__ __forceinline__ void fr_pow5_t(u32 (&r)[8], const u32 (&a)[8], u32 &unc) {
  u32 a2[8], a4[8];
  fr_sqr_t<EXACT>(a2, a, unc);
  fr_sqr_t<EXACT>(a4, a2, unc);
  fr_mul_t<EXACT>(r, a4, a, unc);
}
__device__ __forceinline__ void fr_pow5(u32 (&r)[8], const u32 (&a)[8]) {
  u32 unused = 0;
  fr_pow5_t<true>(r, a, unused);
}

}  // namespace cuzk
