// octet.cuh -- the cooperative Poseidon path: EIGHT LANES PER PERMUTATION ("octet"), lane m holds 32-bit word m of every
// field element.  Used for launches narrower than one wave of the one-thread-per-unit kernels (upper Merkle levels, small
// proof batches, 4096-hash calls), where the latency of one permutation -- not throughput -- is the bound: a lone warp of
// the one-thread kernel spends ~5 pipe cycles on every one of its 33.7 k IMAD.WIDE whatever the number of active lanes,
// so spreading one permutation over eight lanes cuts the instruction stream per permutation by ~8x and leaves the
// shuffle round trips as the critical path.
//
// It evaluates exactly the reference functions of fr.cuh / poseidon.cuh (src/poseidon/field_arithmetic.cpp:172-338,
// src/poseidon/poseidon.cpp:60-167); only the evaluation order inside a multi-word operation differs:
//
//   product      lane j multiplies the whole operand a (replicated, or the constant k) by its own word b_j: one serial
//                IMAD.WIDE chain -> a 9-word row.  "Transposed sum": lane m fetches word d of lane (m - d)'s row for
//                d = 0..8 (9 SHFL) and adds those with d <= m into column m, the others into column m + 8.
//   carries      a column sum is < 9 * 2^32; ONE neighbour shuffle brings the high part of the column below.  That add can
//                overflow again only if the low word is within 16 of 2^32 (~2^-28): the lane then sets `unc` instead of
//                rippling further.
//   reductions   the quotients of `reduce` (<= 5) are read off the top word alone (quot_top: x7 against the top words of
//                1p..5p; equality = undecidable -> `unc`), the multiple of p is subtracted lane-wise as a signed 64-bit
//                value and normalised by one more neighbour shuffle.
//   MDS + ARC    the linear-form row of poseidon.cuh (mds_row_fast) on word-distributed state; the nine wrap bits are
//                evaluated one per lane and collected with a ballot; the next round's constants are added before the row's
//                carry pass.
//
// Any `unc` in an octet means "this unit's fast evaluation is not trustworthy": the unit is evaluated again by one lane on
// the exact one-thread path (sponge_exact / permute_exact in poseidon.cuh), exactly like the one-thread kernels do.
//
// The file is plain C++ apart from the three communication primitives, so tests/cpp/octet_emul.cpp can run the same source
// on the host (eight threads in lockstep) against the oracle.
#pragma once
#include "fr_consts.cuh"

#ifdef CUZK_OCTET_HOST_EMUL
#define OCT_FN inline
#else
#define OCT_FN __device__ __forceinline__
#endif

namespace cuzk {
namespace oct {

typedef int32_t i32;
typedef int64_t i64;

#ifndef CUZK_OCTET_HOST_EMUL
// lane within the octet; value of x in lane `src` (0..7) of the own octet; 8-bit vote of the own octet
OCT_FN u32 lane8() { return threadIdx.x & 7u; }
OCT_FN u32 shfl(u32 x, u32 src) { return __shfl_sync(0xffffffffu, x, (int)src, 8); }
OCT_FN u32 ballot8(bool p) { return (__ballot_sync(0xffffffffu, p) >> (threadIdx.x & 24u)) & 0xffu; }
OCT_FN u32 umulhi32(u32 a, u32 b) { return __umulhi(a, b); }
OCT_FN u32 popc32(u32 a) { return (u32)__popc(a); }
#else
u32 lane8();
u32 shfl(u32 x, u32 src);
u32 ballot8(bool p);
inline u32 umulhi32(u32 a, u32 b) { return (u32)(((u64)a * (u64)b) >> 32); }
inline u32 popc32(u32 a) { return (u32)__builtin_popcount(a); }
#endif

// per-lane constants, set up once per kernel
struct Lane {
  u32 m;      // lane within the octet = word index
  u32 P;      // word m of p
  u32 NP;     // word m of W - p
  u32 tc;     // MDS constant of the wrap-bit term this lane evaluates (term m: row m / 3, column m % 3)
  u32 tj;     // its column
};

OCT_FN u32 pick8(u32 m, u32 v0, u32 v1, u32 v2, u32 v3, u32 v4, u32 v5, u32 v6, u32 v7) {
  const u32 a = (m & 1u) ? v1 : v0, b = (m & 1u) ? v3 : v2, c = (m & 1u) ? v5 : v4, d = (m & 1u) ? v7 : v6;
  const u32 e = (m & 2u) ? b : a, f = (m & 2u) ? d : c;
  return (m & 4u) ? f : e;
}

OCT_FN Lane make_lane() {
  Lane L;
  L.m = lane8();
  L.P = pick8(L.m, CUZK_P0, CUZK_P1, CUZK_P2, CUZK_P3, CUZK_P4, CUZK_P5, CUZK_P6, CUZK_P7);
  L.NP = pick8(L.m, CUZK_NP0, CUZK_NP1, CUZK_NP2, CUZK_NP3, CUZK_NP4, CUZK_NP5, CUZK_NP6, CUZK_NP7);
  L.tc = pick8(L.m, 7u, 23u, 8u, 26u, 5u, 4u, 15u, 20u);   // MDS = [[7,23,8],[26,5,4],[15,20,9]] (poseidon.cpp:46-58), row-major
  L.tj = pick8(L.m, 0u, 1u, 2u, 0u, 1u, 2u, 0u, 1u);
  return L;
}

// ---- building blocks ----------------------------------------------------------------------------------------------------

// quotient floor(x / p) (<= 5) of a 256-bit x from its top word; an x7 equal to the top word of a multiple of p cannot be
// decided here and raises `unc`
OCT_FN u32 quot_top(u32 x7, u32 &unc) {
  u32 e = 0;
#define CUZK_OCT_QSTEP(M)                                         \
  {                                                               \
    constexpr u32 T = mulp_limb(M, 7);                            \
    e += (x7 > T) ? 1u : 0u;                                      \
    unc |= (((x7 ^ T) >> CUZK_UNC_WIDEN) == 0u) ? 1u : 0u;        \
  }
  CUZK_OCT_QSTEP(1) CUZK_OCT_QSTEP(2) CUZK_OCT_QSTEP(3) CUZK_OCT_QSTEP(4) CUZK_OCT_QSTEP(5)
#undef CUZK_OCT_QSTEP
  return e;
}

// one carry pass over word-distributed lane values v_m = lo + 2^32 * c (c a small signed carry, |c| < 2^15 here):
// word m becomes lo_m + c_{m-1}; the carry out of lane 7 is dropped (arithmetic mod W).  A result outside [0, 2^32)
// would have to ripple on: `unc`.
OCT_FN u32 carry_pass(u32 lo, i32 c, u32 m, u32 &unc) {
  i32 cin = (i32)shfl((u32)c, (m - 1u) & 7u);
  cin = (m == 0u) ? 0 : cin;
  const i64 r = (i64)(u64)lo + (i64)cin;
  unc |= ((u64)r >> 32) != 0ull ? 1u : 0u;
  return (u32)r;
}

// top word (word 7) of the same normalisation, computed by every lane from lanes 7 and 6
OCT_FN u32 top_word(u32 lo, i32 c) { return shfl(lo, 7u) + shfl((u32)c, 6u); }

// row product: R[0..8] = a[0..7] * b
OCT_FN void row_mul(u32 (&R)[9], const u32 (&a)[8], u32 b) {
  u64 t = (u64)a[0] * (u64)b;
  R[0] = (u32)t;
#pragma unroll
  for (int i = 1; i < 8; ++i) {
    t = (u64)a[i] * (u64)b + (t >> 32);
    R[i] = (u32)t;
  }
  R[8] = (u32)(t >> 32);
}
// R[0..8] = k * b   (k = 2^256 mod p, immediates)
OCT_FN void row_mul_k(u32 (&R)[9], u32 b) {
  u64 t = (u64)k_limb(0) * (u64)b;
  R[0] = (u32)t;
#pragma unroll
  for (int i = 1; i < 8; ++i) {
    t = (u64)k_limb(i) * (u64)b + (t >> 32);
    R[i] = (u32)t;
  }
  R[8] = (u32)(t >> 32);
}
// R[0..7] = (k * b) mod 2^256
OCT_FN void row_mul_k_low(u32 (&R)[8], u32 b) {
  u64 t = (u64)k_limb(0) * (u64)b;
  R[0] = (u32)t;
#pragma unroll
  for (int i = 1; i < 7; ++i) {
    t = (u64)k_limb(i) * (u64)b + (t >> 32);
    R[i] = (u32)t;
  }
  R[7] = k_limb(7) * b + (u32)(t >> 32);
}

// transposed sum of the eight lanes' rows (lane j's row sits at word offset j): column m -> lo, column m + 8 -> hi
OCT_FN void tsum9(u64 &lo, u64 &hi, const u32 (&R)[9], u32 m) {
  u64 l = 0, all = 0;
#pragma unroll
  for (int d = 0; d < 9; ++d) {
    const u32 v = shfl(R[d], (m - (u32)d) & 7u);
    all += v;
    if (d == 0) l += v;
    else if (d < 8) l += ((u32)d <= m) ? v : 0u;
  }
  lo = l;
  hi = all - l;
}
// low columns only (product mod 2^256), added to `acc`
OCT_FN u64 tsum8_low(const u32 (&R)[8], u32 m, u32 acc) {
  u64 l = acc;
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    const u32 v = shfl(R[d], (m - (u32)d) & 7u);
    l += (d == 0 || (u32)d <= m) ? v : 0u;
  }
  return l;
}

// normalise the 16 column sums of a product: lane m gets words m (low) and m + 8 (high)
OCT_FN void norm16(u32 &low, u32 &high, u64 lo, u64 hi, u32 m, u32 &unc) {
  const u32 clo = (u32)(lo >> 32), chi = (u32)(hi >> 32);   // < 16
  const u32 pin = shfl(clo | (chi << 8), (m - 1u) & 7u);
  const u32 cin_lo = (m == 0u) ? 0u : (pin & 0xffu);
  const u32 cin_hi = (m == 0u) ? (pin & 0xffu) : (pin >> 8);   // word 8 receives the carry of column 7
  low = (u32)lo + cin_lo;
  high = (u32)hi + cin_hi;
  unc |= (low < cin_lo) ? 1u : 0u;
  unc |= (high < cin_hi) ? 1u : 0u;
}

// all eight words of a word-distributed element
OCT_FN void gather(u32 (&r)[8], u32 x) {
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = shfl(x, (u32)i);
}

// ---- multiply : field_arithmetic.cpp:221-238 + reduce_512 :250-330 (any 256-bit operands) ----------------------------------
//   prod = a*b = high*W + low ;  Mh = high*k = mh*W + ml ;  t = (ml + (mh*k mod W)) mod W ;  hc = mh != 0 ? t mod p : t ;
//   r = ((low + hc) mod W) mod p
// NS independent multiplications are evaluated side by side (the three S-boxes of a full round) so their shuffle
// latencies overlap.  a: replicated operand, b: this lane's word of the other operand.
template <int NS>
OCT_FN void mulred(u32 (&r)[NS], const u32 (&a)[NS][8], const u32 (&b)[NS], const Lane &L, u32 &unc) {
  const u32 m = L.m;
  u32 low[NS], high[NS], ml[NS], mh[NS];
#pragma unroll
  for (int e = 0; e < NS; ++e) {
    u32 R[9];
    u64 lo, hi;
    row_mul(R, a[e], b[e]);
    tsum9(lo, hi, R, m);
    norm16(low[e], high[e], lo, hi, m, unc);
  }
#pragma unroll
  for (int e = 0; e < NS; ++e) {
    u32 R[9];
    u64 lo, hi;
    row_mul_k(R, high[e]);
    tsum9(lo, hi, R, m);
    norm16(ml[e], mh[e], lo, hi, m, unc);
  }
  u32 e1[NS], tw[NS];
#pragma unroll
  for (int e = 0; e < NS; ++e) {
    const bool any_mh = ballot8(mh[e] != 0u) != 0u;   // the reference reduces t only when mh != 0 (:303)
    u32 R[8];
    row_mul_k_low(R, mh[e]);
    const u64 T = tsum8_low(R, m, ml[e]);
    const u32 tl = (u32)T;
    const i32 tc = (i32)(u32)(T >> 32);
    const u32 t7 = top_word(tl, tc);
    tw[e] = carry_pass(tl, tc, m, unc);
    u32 q = quot_top(t7, unc);
    e1[e] = any_mh ? q : 0u;
  }
  u32 e2[NS], uw[NS];
#pragma unroll
  for (int e = 0; e < NS; ++e) {
    const i64 y = (i64)(u64)low[e] + (i64)(u64)tw[e] - (i64)((u64)e1[e] * (u64)L.P);
    const u32 yl = (u32)(u64)y;
    const i32 yc = (i32)(y >> 32);
    const u32 u7 = top_word(yl, yc);
    uw[e] = carry_pass(yl, yc, m, unc);
    e2[e] = quot_top(u7, unc);
  }
#pragma unroll
  for (int e = 0; e < NS; ++e) {
    const i64 z = (i64)(u64)uw[e] - (i64)((u64)e2[e] * (u64)L.P);
    r[e] = carry_pass((u32)(u64)z, (i32)(z >> 32), m, unc);
  }
}

// x -> x^5 as the reference does: x2 = x*x, x4 = x2*x2, x5 = x4*x (field_arithmetic.cpp:332-338)
template <int NS>
OCT_FN void sbox(u32 (&x)[NS], const Lane &L, u32 &unc) {
  u32 xr[NS][8], x2r[NS][8], x2[NS], x4[NS], x5[NS];
#pragma unroll
  for (int e = 0; e < NS; ++e) gather(xr[e], x[e]);
  mulred<NS>(x2, xr, x, L, unc);
#pragma unroll
  for (int e = 0; e < NS; ++e) gather(x2r[e], x2[e]);
  mulred<NS>(x4, x2r, x2, L, unc);
  mulred<NS>(x5, xr, x4, L, unc);
#pragma unroll
  for (int e = 0; e < NS; ++e) x[e] = x5[e];
}

// add : field_arithmetic.cpp:172-182 for arbitrary 256-bit operands: (a + b) mod W, then the full reduce
template <int NS>
OCT_FN void add_reduce(u32 (&r)[NS], const u32 (&a)[NS], const u32 (&b)[NS], const Lane &L, u32 &unc) {
  u32 v[NS], e[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const u64 y = (u64)a[i] + (u64)b[i];
    const u32 yl = (u32)y;
    const i32 yc = (i32)(u32)(y >> 32);
    const u32 v7 = top_word(yl, yc);
    v[i] = carry_pass(yl, yc, L.m, unc);
    e[i] = quot_top(v7, unc);
  }
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const i64 z = (i64)(u64)v[i] - (i64)((u64)e[i] * (u64)L.P);
    r[i] = carry_pass((u32)(u64)z, (i32)(z >> 32), L.m, unc);
  }
}

// wrap bit of one MDS term C * s (see mds_wrap_bit in poseidon.cuh) from the two top words of s
OCT_FN u32 wrap_bit(u32 C, u32 s6, u32 s7, u32 &unc) {
  const u64 y = (u64)s6 * (u64)C;
  const u64 z = (u64)s7 * (u64)C + (y >> 32);
  const u32 h = (u32)(z >> 32), low7 = (u32)z;
  unc |= ((u32)y >= 0xFFFFFFE0u) ? 1u : 0u;
  const u32 fl = h * CUZK_K7 + ((h * 5u) >> 3);
  const u32 t = low7 + fl;
  unc |= (t == 0xFFFFFFFFu) ? 1u : 0u;
  return (t < fl) ? 1u : 0u;
}

// apply_mds_matrix (poseidon.cpp:148-167) in the linear form of mds_row_fast, followed -- when has_rc -- by the next round's
// add_round_constants (:128-134): rc[i] is this lane's word of the constant for state element i (0 above word 1).
OCT_FN void mds_arc(u32 (&s)[3], const u32 (&rc)[3], bool has_rc, const Lane &L, u32 &unc) {
  const u32 m = L.m;
  u64 Ls[3];
  Ls[0] = (u64)s[0] * 7u + (u64)s[1] * 23u + (u64)s[2] * 8u;
  Ls[1] = (u64)s[0] * 26u + (u64)s[1] * 5u + (u64)s[2] * 4u;
  Ls[2] = (u64)s[0] * 15u + (u64)s[1] * 20u + (u64)s[2] * 9u;
  u32 s7[3], s6[3], l7[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    s7[j] = shfl(s[j], 7u);
    s6[j] = shfl(s[j], 6u);
  }
  const u32 pk = (u32)(Ls[0] >> 32) | ((u32)(Ls[1] >> 32) << 8) | ((u32)(Ls[2] >> 32) << 16);   // each high part <= 46
  const u32 pk7 = shfl(pk, 7u), pk6 = shfl(pk, 6u);
#pragma unroll
  for (int i = 0; i < 3; ++i) l7[i] = shfl((u32)Ls[i], 7u);
  // wrap bits: term m here, term (2,2) everywhere
  const u32 a7 = (L.tj == 0u) ? s7[0] : (L.tj == 1u ? s7[1] : s7[2]);
  const u32 a6 = (L.tj == 0u) ? s6[0] : (L.tj == 1u ? s6[1] : s6[2]);
  const u32 wb = wrap_bit(L.tc, a6, a7, unc);
  const u32 w22 = wrap_bit(9u, s6[2], s7[2], unc);
  const u32 bal = ballot8(wb != 0u);
  u32 wsum[3];
  wsum[0] = popc32(bal & 0x07u);
  wsum[1] = popc32(bal & 0x38u);
  wsum[2] = popc32(bal & 0xC0u) + w22;
  u32 v[3], ge[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    // quotient estimate from the top of S = sum_j C_ij s_j and the wrap count (mds_row_fast)
    const u32 S7 = l7[i] + ((pk6 >> (8 * i)) & 0xffu);
    const u32 S8 = ((pk7 >> (8 * i)) & 0xffu) + (S7 < l7[i] ? 1u : 0u);
    const u32 a4 = (S8 << 28) | (S7 >> 4);
    const u32 lp = a4 - ((wsum[i] * (CUZK_K7 + 1u) + 15u) >> 4);
    const u32 qhat = umulhi32(lp, kQuotMagic) >> 25;
    const u32 q = qhat - 5u * wsum[i];
    const u64 y = Ls[i] + (u64)q * (u64)L.NP + (u64)rc[i];
    const u32 yl = (u32)y;
    const i32 yc = (i32)(u32)(y >> 32);
    const u32 v7 = top_word(yl, yc);
    v[i] = carry_pass(yl, yc, m, unc);
    ge[i] = (v7 > CUZK_P7) ? 1u : 0u;
    unc |= (((v7 ^ CUZK_P7) >> CUZK_UNC_WIDEN) == 0u) ? 1u : 0u;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const i64 z = (i64)(u64)v[i] - (i64)(u64)(ge[i] ? L.P : 0u);
    s[i] = carry_pass((u32)(u64)z, (i32)(z >> 32), m, unc);
    // a state + constant whose top word reaches p's may need the reference's subtraction (arc_fast in poseidon.cuh)
    if (has_rc) unc |= (m == 7u && (s[i] >> CUZK_UNC_WIDEN) >= (CUZK_P7 >> CUZK_UNC_WIDEN)) ? 1u : 0u;
  }
}

// this lane's word of round constant idx (all constants are < 2^64: words 0 and 1)
template <class RcTable>
OCT_FN u32 rc_word(const RcTable &rct, int idx, u32 m) {
  const u32 c0 = rct(idx, 0), c1 = rct(idx, 1);
  return (m == 0u) ? c0 : (m == 1u ? c1 : 0u);
}

// permutation : poseidon.cpp:60-87 on a word-distributed state (any 256-bit values on entry)
template <class RcTable>
OCT_FN void permute(u32 (&s)[3], const RcTable &rct, const Lane &L, u32 &unc) {
  {
    u32 rc[3], t[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) rc[i] = rc_word(rct, i, L.m);
    add_reduce<3>(t, s, rc, L, unc);
#pragma unroll
    for (int i = 0; i < 3; ++i) s[i] = t[i];
  }
#pragma unroll 1
  for (int round = 0; round < 64; ++round) {
    const bool full = (round < 4) || (round >= 60);
    if (full) {
      sbox<3>(s, L, unc);
    } else {
      u32 x[1] = {s[0]};
      sbox<1>(x, L, unc);
      s[0] = x[0];
    }
    u32 rc[3];
    const bool has_rc = round < 63;
#pragma unroll
    for (int i = 0; i < 3; ++i) rc[i] = has_rc ? rc_word(rct, 3 * (round + 1) + i, L.m) : 0u;
    mds_arc(s, rc, has_rc, L, unc);
  }
}

// hash_multiple / sponge over `width` inputs (poseidon.cpp:98-126): out = this lane's word of the digest.
// load(i) returns this lane's word of input i.  Returns the octet's `unc` vote: non-zero = evaluate this unit again exactly.
template <class RcTable, class Loader>
OCT_FN u32 sponge(u32 &out, u32 ds_lo, u32 ds_hi, int width, const RcTable &rct, const Lane &L, Loader load) {
  u32 unc = 0;
  u32 s[3];
  s[0] = (L.m == 0u) ? ds_lo : (L.m == 1u ? ds_hi : 0u);
  s[1] = 0u;
  s[2] = 0u;
#pragma unroll 1
  for (int i = 0; i < width; i += 2) {
    if (i + 1 < width) {
      u32 a[2] = {s[1], s[2]}, x[2] = {load(i), load(i + 1)}, r[2];
      add_reduce<2>(r, a, x, L, unc);
      s[1] = r[0];
      s[2] = r[1];
    } else {
      u32 a[1] = {s[1]}, x[1] = {load(i)}, r[1];
      add_reduce<1>(r, a, x, L, unc);
      s[1] = r[0];
    }
    permute(s, rct, L, unc);
  }
  out = s[1];
  return ballot8(unc != 0u);
}

}  // namespace oct
}  // namespace cuzk
