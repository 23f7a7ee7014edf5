// coop16.cuh -- the sixteen-lanes-per-permutation layout of the cooperative path (Wide16 of coop.cuh), as ONE non-templated
// piece of code.  It computes exactly what coop.cuh's algorithms compute with Y = Wide16 (tests/cpp/coop_emul.cpp checks both
// against the oracle); it exists because of how ptxas schedules it: in this form the eight shuffles that replicate x^2 before
// the second multiplication of an S-box are issued back to back, in the templated form they are interleaved one by one with the
// multiplies that consume them (registers recycled), which costs 14 % on the latency-bound launches this layout is for
// (93 us against 107 us per permutation, same box, profiles/r02_tuning_notes.md).  The kernels of the narrowest launches use
// this file; see coop.cuh for the description of the method.
#pragma once
#include "fr_consts.cuh"
#include "coop.cuh"


namespace cuzk {
namespace coop16 {

typedef int32_t i32;
typedef int64_t i64;

constexpr int kGroup = 16;   // lanes per permutation

#ifndef CUZK_COOP_HOST_EMUL
// lane within the group; value of x in lane `src` (0..15) of the own group; 16-bit vote of the own group
COOP_FN u32 lane16() { return threadIdx.x & 15u; }
COOP_FN u32 shfl(u32 x, u32 src) { return __shfl_sync(0xffffffffu, x, (int)src, 16); }
COOP_FN u32 ballot16(bool p) { return (__ballot_sync(0xffffffffu, p) >> (threadIdx.x & 16u)) & 0xffffu; }
COOP_FN u32 umulhi32(u32 a, u32 b) { return __umulhi(a, b); }
COOP_FN u32 popc32(u32 a) { return (u32)__popc(a); }
#else
inline u32 lane16() { return coop::emul_lane(); }
inline u32 shfl(u32 x, u32 src) { return coop::emul_shfl(x, src & 15u); }
inline u32 ballot16(bool p) { return coop::emul_ballot(p); }
inline u32 umulhi32(u32 a, u32 b) { return (u32)(((u64)a * (u64)b) >> 32); }
inline u32 popc32(u32 a) { return (u32)__builtin_popcount(a); }
#endif

// per-lane constants, set up once per kernel
struct Lane {
  u32 g;      // lane within the group
  bool low;   // g < 8: this lane holds a word of every element
  u32 P;      // word g of p            (0 in lanes 8..15)
  u32 NP;     // word g of W - p        (0 in lanes 8..15)
  u32 tc;     // MDS constant of the wrap-bit term this lane evaluates (term g: row g / 3, column g % 3; 0 = none)
  u32 tj;     // its column
  u32 prev8;  // source lane of an element carry pass: g - 1, except lanes 0 and 8 which read lane 15 (always zero)
  u32 arc;    // lane 7: 2^32 - (p's top word, low CUZK_UNC_WIDEN bits cleared); else 0   (see mds_arc)
};

// "this unit's fast evaluation cannot be trusted" in two accumulators: ovf collects bits (any non-zero = flagged), near is
// a running minimum of distances to an undecidable comparison (below 2^CUZK_UNC_WIDEN = flagged; 0 = truly undecidable)
struct Flags {
  u32 ovf = 0u;
  u32 near = 0xffffffffu;
};
COOP_FN bool flagged(const Flags &F) { return F.ovf != 0u || F.near < (1u << CUZK_UNC_WIDEN); }
COOP_FN u32 umin32(u32 a, u32 b) { return a < b ? a : b; }

COOP_FN u32 pick8(u32 m, u32 v0, u32 v1, u32 v2, u32 v3, u32 v4, u32 v5, u32 v6, u32 v7) {
  const u32 a = (m & 1u) ? v1 : v0, b = (m & 1u) ? v3 : v2, c = (m & 1u) ? v5 : v4, d = (m & 1u) ? v7 : v6;
  const u32 e = (m & 2u) ? b : a, f = (m & 2u) ? d : c;
  return (m & 4u) ? f : e;
}

COOP_FN Lane make_lane() {
  Lane L;
  L.g = lane16();
  L.low = L.g < 8u;
  const u32 m = L.g & 7u;
  L.P = L.low ? pick8(m, CUZK_P0, CUZK_P1, CUZK_P2, CUZK_P3, CUZK_P4, CUZK_P5, CUZK_P6, CUZK_P7) : 0u;
  L.NP = L.low ? pick8(m, CUZK_NP0, CUZK_NP1, CUZK_NP2, CUZK_NP3, CUZK_NP4, CUZK_NP5, CUZK_NP6, CUZK_NP7) : 0u;
  // MDS = [[7,23,8],[26,5,4],[15,20,9]] (poseidon.cpp:46-58), row-major: terms 0..8 in lanes 0..8
  L.tc = L.low ? pick8(m, 7u, 23u, 8u, 26u, 5u, 4u, 15u, 20u) : (L.g == 8u ? 9u : 0u);
  L.tj = L.low ? pick8(m, 0u, 1u, 2u, 0u, 1u, 2u, 0u, 1u) : 2u;
  L.prev8 = (L.g == 8u) ? 15u : ((L.g - 1u) & 15u);
  L.arc = (L.g == 7u) ? (0u - ((CUZK_P7 >> CUZK_UNC_WIDEN) << CUZK_UNC_WIDEN)) : 0u;
  return L;
}

// ---- building blocks ----------------------------------------------------------------------------------------------------

// quotient floor(x / p) (<= 5) of a 256-bit x from its top word x7.  The top words of 1p..5p are i * D - 1 with
// D = p7 + 1 (checked below), so floor(x / p) = floor(x7 / D) unless x7 + 1 is a multiple of D, where the lower words decide:
// that distance goes to F.near.  floor(y / D) by a multiply-high is exact for every y that is a multiple of D or not within
// one of the next multiple, which is all this needs.
constexpr u32 kTopD = CUZK_P7 + 1u;
constexpr u32 kTopDMagic = (u32)(((1ull << 61) + kTopD - 1u) / kTopD);   // ceil(2^61 / D)
static_assert(mulp_limb(1, 7) == 1u * kTopD - 1u && mulp_limb(2, 7) == 2u * kTopD - 1u && mulp_limb(3, 7) == 3u * kTopD - 1u &&
                  mulp_limb(4, 7) == 4u * kTopD - 1u && mulp_limb(5, 7) == 5u * kTopD - 1u,
              "top words of the multiples of p");
COOP_FN u32 quot_top(u32 x7, Flags &F) {
  const u32 y = x7 + 1u;                                  // wraps to 0 for x7 = 2^32 - 1: flagged (k = 0, rem = 0), harmless
  const u32 k = umulhi32(y, kTopDMagic) >> 29;
  const u32 rem = y - k * kTopD;
  F.near = umin32(F.near, rem);
  return k;
}

// One carry pass over lane values v_g = lo + 2^32 * c (c a small signed carry): word g becomes lo_g + c_{g-1}.  A result
// outside [0, 2^32) would have to ripple on: F.ovf.
//   carry_pass16: over all sixteen columns of a product (column 15 never carries out, so lane 0 receives zero)
//   carry_pass8 : over the eight words of an element; the carry out of word 7 is dropped (arithmetic mod W): lane 8 reads
//                 lane 15 instead of lane 7, so lanes 8..15 stay zero
COOP_FN u32 carry_pass16(u32 lo, u32 c, u32 g, Flags &F) {
  const u32 cin = shfl(c, (g - 1u) & 15u);
  const u64 r = (u64)lo + (u64)cin;
  F.ovf |= (u32)(r >> 32);
  return (u32)r;
}
COOP_FN u32 carry_pass8(u32 lo, i32 c, const Lane &L, Flags &F) {
  const i32 cin = (i32)shfl((u32)c, L.prev8);
  const i64 r = (i64)(u64)lo + (i64)cin;
  F.ovf |= (u32)((u64)r >> 32);
  return (u32)(u64)r;
}

// carry_pass8 for non-negative carries, exact whatever ripples (coop::carry_exact explains the two votes)
COOP_FN u32 carry_exact8(u32 lo, u32 c, const Lane &L) {
  const u64 r = (u64)lo + (u64)shfl(c, L.prev8);
  const u32 G = ballot16((u32)(r >> 32) != 0u) & 0xffu, P = ballot16((u32)r == 0xFFFFFFFFu) & 0xffu;
  const u32 cin = ((G | P) + G) ^ P;
  return (u32)r + (L.low ? ((cin >> L.g) & 1u) : 0u);
}

// Top word (word 7) of the normalised value, read BEFORE its carry pass completes: lo_7 + c_6.  That is the exact word 7
// whenever the carry pass of the same lane values raises no flag (by induction from lane 0 every carry c_g is then the true
// carry out of word g), so callers run carry_pass8 on the side for its flag and do not wait for its result.
COOP_FN u32 top_word(u32 lo, i32 c) { return shfl(lo, 7u) + shfl((u32)c, 6u); }

// ---- row products ----------------------------------------------------------------------------------------------------
// R[0..8] = a[0..7] * b as eight independent 32x32->64 products (even positions fill words 0..7, odd positions words 1..8
// without overlapping each other) and one carry chain that adds the two.
COOP_FN void merge_even_odd(u32 (&R)[9], const u64 (&ev)[4], const u64 (&od)[4]) {
  const u32 E[9] = {(u32)ev[0], (u32)(ev[0] >> 32), (u32)ev[1], (u32)(ev[1] >> 32), (u32)ev[2], (u32)(ev[2] >> 32),
                    (u32)ev[3], (u32)(ev[3] >> 32), 0u};
  const u32 O[9] = {0u, (u32)od[0], (u32)(od[0] >> 32), (u32)od[1], (u32)(od[1] >> 32), (u32)od[2], (u32)(od[2] >> 32),
                    (u32)od[3], (u32)(od[3] >> 32)};
  R[0] = E[0];
#ifdef __CUDA_ARCH__
  R[1] = add_cc(E[1], O[1]);
#pragma unroll
  for (int i = 2; i < 8; ++i) R[i] = addc_cc(E[i], O[i]);
  R[8] = addc(E[8], O[8]);
#else
  u64 c = 0;
  for (int i = 1; i < 9; ++i) {
    c += (u64)E[i] + (u64)O[i];
    R[i] = (u32)c;
    c >>= 32;
  }
#endif
}
COOP_FN void row_mul(u32 (&R)[9], const u32 (&a)[8], u32 b) {
  u64 ev[4], od[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ev[i] = (u64)a[2 * i] * (u64)b;
    od[i] = (u64)a[2 * i + 1] * (u64)b;
  }
  merge_even_odd(R, ev, od);
}
// R[0..8] = k * b   (k = 2^256 mod p, immediates)
COOP_FN void row_mul_k(u32 (&R)[9], u32 b) {
  u64 ev[4], od[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ev[i] = (u64)k_limb(2 * i) * (u64)b;
    od[i] = (u64)k_limb(2 * i + 1) * (u64)b;
  }
  merge_even_odd(R, ev, od);
}
// R[0..7] = (k * b) mod 2^256
COOP_FN void row_mul_k_low(u32 (&R)[8], u32 b) {
  u32 R9[9];
  u64 ev[4], od[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ev[i] = (u64)k_limb(2 * i) * (u64)b;
    od[i] = (i < 3) ? (u64)k_limb(2 * i + 1) * (u64)b : (u64)(k_limb(7) * b);   // word 8 is not needed
  }
  merge_even_odd(R9, ev, od);
#pragma unroll
  for (int i = 0; i < 8; ++i) R[i] = R9[i];
}

// sum of three 32-bit words (one 3-input add with two carries on the GPU)
COOP_FN u64 add3(u32 a, u32 b, u32 c) { return (u64)a + (u64)b + (u64)c; }

// transposed sum: column g of the product whose rows sit in lanes base .. base + 7 (row of lane base + j at word offset j);
// a lane with zsrc reads lane 0 (a zero row in products by k) for every word
template <int ND>
COOP_FN u64 tsum(const u32 (&R)[ND], u32 g, u32 base, bool zsrc = false) {
  u32 v[9];
#pragma unroll
  for (int d = 0; d < 9; ++d) v[d] = (d < ND) ? shfl(R[d < ND ? d : 0], zsrc ? 0u : ((g - (u32)d + base) & 15u)) : 0u;
  return add3(v[0], v[1], v[2]) + add3(v[3], v[4], v[5]) + add3(v[6], v[7], v[8]);
}

// all eight words of a word-distributed element
COOP_FN void gather(u32 (&r)[8], u32 x) {
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = shfl(x, (u32)i);
}

// ---- multiply : field_arithmetic.cpp:221-238 + reduce_512 :250-330 (any 256-bit operands) ----------------------------------
//   prod = a*b = high*W + low ;  Mh = high*k = mh*W + ml ;  t = (ml + (mh*k mod W)) mod W ;  hc = mh != 0 ? t mod p : t ;
//   r = ((low + hc) mod W) mod p
// NS independent multiplications are evaluated side by side (the three S-boxes of a full round) so their shuffle
// latencies overlap.  a: replicated operand, b: this lane's word of the other operand (zero in lanes 8..15).
template <int NS>
COOP_FN void mulred(u32 (&r)[NS], const u32 (&a)[NS][8], const u32 (&b)[NS], const Lane &L, Flags &F) {
  const u32 g = L.g;
  u32 w1[NS], w2[NS];   // words of prod (low | high by lane), then of Mh (ml | mh)
#pragma unroll
  for (int e = 0; e < NS; ++e) {
    u32 R[9];
    row_mul(R, a[e], b[e]);
    const u64 T = tsum<9>(R, g, 0u);
    w1[e] = carry_pass16((u32)T, (u32)(T >> 32), g, F);
  }
#pragma unroll
  for (int e = 0; e < NS; ++e) {
    u32 R[9];
    row_mul_k(R, L.low ? 0u : w1[e]);
    const u64 T = tsum<9>(R, g, 8u);
    w2[e] = carry_pass16((u32)T, (u32)(T >> 32), g, F);
  }
  // t = (ml + mh*k) mod W stays un-normalised (lane value T < 2^36): only its top word is needed, for the quotient e1;
  // y = low + t - e1*p likewise, for e2;  z = y - e2*p is normalised by the one carry pass of this tail.
  i64 y[NS];
#pragma unroll
  for (int e = 0; e < NS; ++e) {
    const bool any_mh = (ballot16(w2[e] != 0u) & 0xff00u) != 0u;   // the reference reduces t only when mh != 0 (:303)
    u32 R[8];
    row_mul_k_low(R, L.low ? 0u : w2[e]);
    const u64 T = tsum<8>(R, g, 8u, !L.low) + (u64)(L.low ? w2[e] : 0u);
    const u32 t7 = top_word((u32)T, (i32)(u32)(T >> 32));
    (void)carry_pass8((u32)T, (i32)(u32)(T >> 32), L, F);   // flag only: t itself is never normalised
    const u32 q = quot_top(t7, F);
    const u32 e1 = any_mh ? q : 0u;
    y[e] = (i64)(T + (u64)(L.low ? w1[e] : 0u)) - (i64)((u64)e1 * (u64)L.P);
  }
#pragma unroll
  for (int e = 0; e < NS; ++e) {
    const u32 u7 = top_word((u32)(u64)y[e], (i32)(y[e] >> 32));
    (void)carry_pass8((u32)(u64)y[e], (i32)(y[e] >> 32), L, F);   // flag only
    const u32 e2 = quot_top(u7, F);
    const i64 z = y[e] - (i64)((u64)e2 * (u64)L.P);
    r[e] = carry_pass8((u32)(u64)z, (i32)(z >> 32), L, F);
  }
}

// x -> x^5 as the reference does: x2 = x*x, x4 = x2*x2, x5 = x4*x (field_arithmetic.cpp:332-338)
template <int NS>
COOP_FN void sbox(u32 (&x)[NS], const Lane &L, Flags &F) {
  u32 xr[NS][8], x2r[NS][8], x2[NS], x4[NS], x5[NS];
#pragma unroll
  for (int e = 0; e < NS; ++e) gather(xr[e], x[e]);
  mulred<NS>(x2, xr, x, L, F);
#pragma unroll
  for (int e = 0; e < NS; ++e) gather(x2r[e], x2[e]);
  mulred<NS>(x4, x2r, x2, L, F);
  mulred<NS>(x5, xr, x4, L, F);
#pragma unroll
  for (int e = 0; e < NS; ++e) x[e] = x5[e];
}

// add : field_arithmetic.cpp:172-182 for arbitrary 256-bit operands: (a + b) mod W, then the full reduce
template <int NS>
COOP_FN void add_reduce(u32 (&r)[NS], const u32 (&a)[NS], const u32 (&b)[NS], const Lane &L, Flags &F) {
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const u64 y = (u64)a[i] + (u64)b[i];
    const u32 v7 = top_word((u32)y, (i32)(u32)(y >> 32));
    (void)carry_pass8((u32)y, (i32)(u32)(y >> 32), L, F);   // flag only
    const u32 e = quot_top(v7, F);
    const i64 z = (i64)y - (i64)((u64)e * (u64)L.P);
    r[i] = carry_pass8((u32)(u64)z, (i32)(z >> 32), L, F);
  }
}

// wrap bit of one MDS term C * s (see mds_wrap_bit in poseidon.cuh) from the two top words of s
COOP_FN u32 wrap_bit(u32 C, u32 s6, u32 s7, Flags &F) {
  const u64 y = (u64)s6 * (u64)C;
  const u64 z = (u64)s7 * (u64)C + (y >> 32);
  const u32 h = (u32)(z >> 32), low7 = (u32)z;
  F.ovf |= ((u32)y >= 0xFFFFFFE0u) ? 1u : 0u;
  const u32 fl = h * CUZK_K7 + ((h * 5u) >> 3);
  const u32 t = low7 + fl;
  F.ovf |= (t == 0xFFFFFFFFu) ? 1u : 0u;
  return (t < fl) ? 1u : 0u;
}

// apply_mds_matrix (poseidon.cpp:148-167) in the linear form of mds_row_fast, followed -- when has_rc -- by the next round's
// add_round_constants (:128-134): rc[i] is this lane's word of the constant for state element i (0 above word 1).
struct MdsShared {   // what the three rows of one MDS layer share
  u64 Ls[3];          // this lane's word of sum_j C_ij s_j, with its carry part
  u32 l7[3], pk7, pk6;   // the same for lanes 7 and 6 (carry parts packed)
#if CUZK_COOP_MDS_LOCAL_TOP
  u32 l6[3];
#endif
  u32 wsum[3];        // wrap counts of the rows
};
COOP_FN void mds_prepare(MdsShared &M, const u32 (&s)[3], const Lane &L, Flags &F) {
  M.Ls[0] = (u64)s[0] * 7u + (u64)s[1] * 23u + (u64)s[2] * 8u;
  M.Ls[1] = (u64)s[0] * 26u + (u64)s[1] * 5u + (u64)s[2] * 4u;
  M.Ls[2] = (u64)s[0] * 15u + (u64)s[1] * 20u + (u64)s[2] * 9u;
  u32 s7[3], s6[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    s7[j] = shfl(s[j], 7u);
    s6[j] = shfl(s[j], 6u);
  }
  const u32 pk = (u32)(M.Ls[0] >> 32) | ((u32)(M.Ls[1] >> 32) << 8) | ((u32)(M.Ls[2] >> 32) << 16);   // each high part <= 46
  M.pk7 = shfl(pk, 7u);
  M.pk6 = shfl(pk, 6u);
#pragma unroll
  for (int i = 0; i < 3; ++i) M.l7[i] = shfl((u32)M.Ls[i], 7u);
#if CUZK_COOP_MDS_LOCAL_TOP
#pragma unroll
  for (int i = 0; i < 3; ++i) M.l6[i] = shfl((u32)M.Ls[i], 6u);
#endif
  // wrap bits: term g in lane g (g = 0..8)
  const u32 a7 = (L.tj == 0u) ? s7[0] : (L.tj == 1u ? s7[1] : s7[2]);
  const u32 a6 = (L.tj == 0u) ? s6[0] : (L.tj == 1u ? s6[1] : s6[2]);
  const u32 wb = wrap_bit(L.tc, a6, a7, F);
  const u32 bal = ballot16(wb != 0u);
  M.wsum[0] = popc32(bal & 0x007u);
  M.wsum[1] = popc32(bal & 0x038u);
  M.wsum[2] = popc32(bal & 0x1C0u);
}
// row i of the layer (+ the next round's constant word rc when has_rc): this lane's word of the new state element i
template <int i>
COOP_FN u32 mds_row(const MdsShared &M, u32 rc, bool has_rc, const Lane &L, Flags &F) {
  // quotient estimate from the top of S = sum_j C_ij s_j and the wrap count (mds_row_fast)
  const u32 S7 = M.l7[i] + ((M.pk6 >> (8 * i)) & 0xffu);
  const u32 S8 = ((M.pk7 >> (8 * i)) & 0xffu) + (S7 < M.l7[i] ? 1u : 0u);
  const u32 a4 = (S8 << 28) | (S7 >> 4);
  const u32 lp = a4 - ((M.wsum[i] * (CUZK_K7 + 1u) + 15u) >> 4);
  const u32 qhat = umulhi32(lp, kQuotMagic) >> 25;
  const u32 q = qhat - 5u * M.wsum[i];
  const u64 y = M.Ls[i] + (u64)q * (u64)L.NP + (u64)rc;            // < 2^40: carries up to 2^8
#if CUZK_COOP_MDS_LOCAL_TOP
  // the top word of y from the two top lane values, computed by every lane itself (the constant has no words above 1)
  const u64 y7 = ((u64)((M.pk7 >> (8 * i)) & 0xffu) << 32 | (u64)M.l7[i]) + (u64)q * (u64)CUZK_NP7;
  const u64 y6 = ((u64)((M.pk6 >> (8 * i)) & 0xffu) << 32 | (u64)M.l6[i]) + (u64)q * (u64)CUZK_NP6;
  const u32 v7 = (u32)y7 + (u32)(y6 >> 32);
#else
  const u32 v7 = top_word((u32)y, (i32)(u32)(y >> 32));
#endif
  u32 r;
#if CUZK_COOP_MDS_EXACT_CARRY
  {   // v7 is the true top word unless word 6 carries out; a margin of one covers the ripple from below (coop.cuh, mds_arc)
    const u64 r6 = (u64)(u32)y + (u64)shfl((u32)(y >> 32), L.prev8);
    F.ovf |= (L.g == 6u && r6 >= 0xFFFFFFFFull) ? 1u : 0u;
  }
  const u32 ge = (v7 > CUZK_P7) ? 1u : 0u;                           // y < 2p: one conditional subtraction, decided by the top word
  F.near = umin32(F.near, v7 ^ CUZK_P7);
  const u64 z = y + (u64)(ge ? L.NP : 0u);                           // y - p as y + (W - p): mod W, lane values stay non-negative
  r = carry_exact8((u32)z, (u32)(z >> 32), L);
#else
  (void)carry_pass8((u32)y, (i32)(u32)(y >> 32), L, F);   // flag only
  const u32 ge = (v7 > CUZK_P7) ? 1u : 0u;                           // y < 2p: one conditional subtraction, decided by the top word
  F.near = umin32(F.near, v7 ^ CUZK_P7);
  const i64 z = (i64)y - (i64)(u64)(ge ? L.P : 0u);
  r = carry_pass8((u32)(u64)z, (i32)(z >> 32), L, F);
#endif
  // a state + constant whose top word reaches p's may need the reference's subtraction (arc_fast in poseidon.cuh):
  // lane 7 adds 2^32 - p7 and flags the carry
  if (has_rc) F.ovf |= (u32)(((u64)r + (u64)L.arc) >> 32);
  return r;
}
COOP_FN void mds_arc(u32 (&s)[3], const u32 (&rc)[3], bool has_rc, const Lane &L, Flags &F) {
  MdsShared M;
  mds_prepare(M, s, L, F);
  s[0] = mds_row<0>(M, rc[0], has_rc, L, F);
  s[1] = mds_row<1>(M, rc[1], has_rc, L, F);
  s[2] = mds_row<2>(M, rc[2], has_rc, L, F);
}

// this lane's word of round constant idx (all constants are < 2^64: words 0 and 1)
template <class RcTable>
COOP_FN u32 rc_word(const RcTable &rct, int idx, u32 g) {
  const u32 c0 = rct(idx, 0), c1 = rct(idx, 1);
  return (g == 0u) ? c0 : (g == 1u ? c1 : 0u);
}

// permutation : poseidon.cpp:60-87 on a word-distributed state (any 256-bit values on entry)
template <class RcTable>
COOP_FN void permute(u32 (&s)[3], const RcTable &rct, const Lane &L, Flags &F) {
  {
    u32 rc[3], t[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) rc[i] = rc_word(rct, i, L.g);
    add_reduce<3>(t, s, rc, L, F);
#pragma unroll
    for (int i = 0; i < 3; ++i) s[i] = t[i];
  }
#if CUZK_COOP_RC_PREFETCH
  u32 pick0 = (L.g == 0u) ? 0xffffffffu : 0u, pick1 = (L.g == 1u) ? 0xffffffffu : 0u;   // lanes 0 and 1 hold the two words of a constant
  coop::pin_register(pick0);
  coop::pin_register(pick1);
#endif
#if CUZK_COOP_PIPELINE_MDS
  MdsShared P;          // the layer whose rows 1 and 2 are still to be evaluated
  u32 prc[2] = {0u, 0u};
  bool pending = false;
#endif
#pragma unroll 1
  for (int round = 0; round < 64; ++round) {
    const bool full = (round < 4) || (round >= 60);
    const bool has_rc = round < 63;
#if CUZK_COOP_RC_PREFETCH
    // the next round's constants: both words of each, fetched by every lane with a uniform index BEFORE the S-box and
    // selected by lane after it, so the ~50-cycle constant loads (and no divergent branch around them) hide behind the S-box;
    // fetched after it they were 12 % of the kernel's time (profiles/r02_tuning_notes.md)
    u32 cw[3][2];
    {
      const int next = 3 * (has_rc ? round + 1 : round);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        cw[i][0] = rct(next + i, 0);
        cw[i][1] = rct(next + i, 1);
        coop::pin_register(cw[i][0]);
        coop::pin_register(cw[i][1]);
      }
    }
#endif
#if CUZK_COOP_PIPELINE_MDS
    // Partial rounds only pass s[0] through the S-box, so rows 1 and 2 of the previous layer are not needed before the next
    // layer: they are evaluated next to the following round's S-box (same basic block, independent chains), where a lone
    // warp has idle issue slots, instead of in front of it.
    if (full) {
      if (pending) {
        s[1] = mds_row<1>(P, prc[0], true, L, F);
        s[2] = mds_row<2>(P, prc[1], true, L, F);
        pending = false;
      }
      sbox<3>(s, L, F);
    } else if (pending) {
      u32 x[1] = {s[0]};
      const u32 t1 = mds_row<1>(P, prc[0], true, L, F);
      const u32 t2 = mds_row<2>(P, prc[1], true, L, F);
      sbox<1>(x, L, F);
      s[0] = x[0];
      s[1] = t1;
      s[2] = t2;
    } else {
      u32 x[1] = {s[0]};
      sbox<1>(x, L, F);
      s[0] = x[0];
    }
#else
    if (full) {
      sbox<3>(s, L, F);
    } else {
      u32 x[1] = {s[0]};
      sbox<1>(x, L, F);
      s[0] = x[0];
    }
#endif
    u32 rc[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#if CUZK_COOP_RC_PREFETCH
      rc[i] = ((cw[i][0] & pick0) | (cw[i][1] & pick1)) & (has_rc ? 0xffffffffu : 0u);   // masks, not selects: ptxas turns the selects into branches
#else
      rc[i] = has_rc ? rc_word(rct, 3 * (round + 1) + i, L.g) : 0u;
#endif
    }
#if CUZK_COOP_PIPELINE_MDS
    if (full) {
      mds_arc(s, rc, has_rc, L, F);
    } else {   // partial rounds are 4..59: a constant always follows
      mds_prepare(P, s, L, F);
      s[0] = mds_row<0>(P, rc[0], true, L, F);
      prc[0] = rc[1];
      prc[1] = rc[2];
      pending = true;
    }
#else
    mds_arc(s, rc, has_rc, L, F);
#endif
  }
}

// hash_multiple / sponge over `width` inputs (poseidon.cpp:98-126): out = this lane's word of the digest (lanes 0..7).
// load(i) returns this lane's word of input i (zero in lanes 8..15).  Returns the group's `unc` vote: non-zero = evaluate
// this unit again exactly.
template <class RcTable, class Loader>
COOP_FN u32 sponge(u32 &out, u32 ds_lo, u32 ds_hi, int width, const RcTable &rct, const Lane &L, Loader load) {
  Flags F;
  u32 s[3];
  s[0] = (L.g == 0u) ? ds_lo : (L.g == 1u ? ds_hi : 0u);
  s[1] = 0u;
  s[2] = 0u;
#pragma unroll 1
  for (int i = 0; i < width; i += 2) {
    if (i + 1 < width) {
      u32 a[2] = {s[1], s[2]}, x[2] = {load(i), load(i + 1)}, r[2];
      add_reduce<2>(r, a, x, L, F);
      s[1] = r[0];
      s[2] = r[1];
    } else {
      u32 a[1] = {s[1]}, x[1] = {load(i)}, r[1];
      add_reduce<1>(r, a, x, L, F);
      s[1] = r[0];
    }
    permute(s, rct, L, F);
  }
  out = s[1];
  return ballot16(flagged(F));
}

}  // namespace coop16
}  // namespace cuzk
