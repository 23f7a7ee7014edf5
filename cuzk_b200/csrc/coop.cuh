// coop.cuh -- the cooperative Poseidon path: ONE PERMUTATION SPREAD OVER A GROUP OF LANES, each lane holding one 32-bit word
// of every field element.  Used for launches narrower than one wave of the one-thread-per-unit kernels (upper Merkle levels,
// small proof batches, 4096-hash calls), where the latency of ONE permutation -- not throughput -- is the bound: a lone warp
// of the one-thread kernel pays ~6 pipe cycles for each of its 33.7 k IMAD.WIDE whatever the number of active lanes (186 us
// per permutation).  Spreading one permutation over a group cuts the instruction stream per permutation several times;
// shuffle round trips become the critical path (measured on B200: SHFL 5 cycles issue / ~26 latency, dependent IMAD.WIDE
// 11-12, tools/latbench.cu).
//
// Two layouts share every algorithm below (template parameter Y):
//   Wide16    sixteen lanes per permutation: lanes 0..7 hold the element words, lanes 8..15 the upper halves of 512-bit
//             products (zero otherwise).  Every lane owns ONE product column, so column sums need no masking: the shortest
//             instruction stream per permutation, i.e. the lowest latency; two permutations per warp.
//   Narrow8   eight lanes per permutation: every lane owns product columns m and m + 8.  About 1.3x the instructions per
//             warp but four permutations per warp: the better choice once a launch no longer fits the chip at one or two
//             warps per SM sub-partition.
//
// They evaluate exactly the reference functions of fr.cuh / poseidon.cuh (src/poseidon/field_arithmetic.cpp:172-338,
// src/poseidon/poseidon.cpp:60-167); only the evaluation order inside a multi-word operation differs:
//
//   product      the lane that holds word b_j multiplies the whole operand a (replicated, or the constant k) by it: eight
//                independent IMAD.WIDE and one carry chain -> a 9-word row at word offset j.  "Transposed sum": the lane of
//                column c fetches word d of the row at offset c - d for d = 0..8 (9 SHFL) and adds them.
//   carries      a column sum is < 9 * 2^32; ONE neighbour shuffle brings the high part of the column below.  That add can
//                overflow again only if the low word is within 16 of 2^32 (~2^-28): the lane then raises a flag instead of
//                rippling further.
//   reductions   the quotients of `reduce` (<= 5) are read off the top word alone (quot_top); the multiple of p is
//                subtracted lane-wise as a signed 64-bit value and normalised by one more neighbour shuffle.  Intermediate
//                values (t, low + hc) are never normalised: only their top words are formed, early.
//   MDS + ARC    the linear-form row of poseidon.cuh (mds_row_fast) on word-distributed state; the nine wrap bits are
//                evaluated one per lane and collected with a ballot; the next round's constants are added before the row's
//                carry pass.
//
// Any flag in a group means "this unit's fast evaluation is not trustworthy": the unit is evaluated again by one lane on
// the exact one-thread path (sponge_exact / permute_exact in poseidon.cuh), exactly like the one-thread kernels do.
//
// The file is plain C++ apart from the communication primitives, so tests/cpp/coop_emul.cpp runs the same source on the
// host (one thread per lane, in lockstep) against the oracle.
#pragma once
#include "fr_consts.cuh"

#ifndef CUZK_COOP_TAIL
#define CUZK_COOP_TAIL 2   // 2: two top-word rounds in the multiply tail; 3: one (both quotients from words fetched in one round; measured 4 % slower)
#endif

#ifndef CUZK_COOP_RC_PREFETCH
#define CUZK_COOP_RC_PREFETCH 1   // 1: round constants are fetched before the S-box and selected after it (see permute)
#endif
#ifndef CUZK_COOP_PIPELINE_MDS
#define CUZK_COOP_PIPELINE_MDS 1   // 1 (sixteen-lane file): rows 1 and 2 of a partial round's MDS layer run beside the next S-box
#endif
#ifndef CUZK_COOP_MDS_LOCAL_TOP
#define CUZK_COOP_MDS_LOCAL_TOP 0   // 1: the MDS layer's top word is computed locally from lanes 7 and 6 of the sums (see mds_arc)
#endif
#ifndef CUZK_COOP_MDS_EXACT_CARRY
#define CUZK_COOP_MDS_EXACT_CARRY 1   // 1: the carries of the MDS layer (up to 2^8 per word) are resolved exactly, see carry_exact
#endif

#ifdef CUZK_COOP_HOST_EMUL
#define COOP_FN inline
#else
#define COOP_FN __device__ __forceinline__
#endif

namespace cuzk {
namespace coop {

typedef int32_t i32;
typedef int64_t i64;

// ---- communication primitives for a group of W lanes (W = 8 or 16) ------------------------------------------------------
#ifndef CUZK_COOP_HOST_EMUL
template <int W> COOP_FN u32 lane_in_group() { return threadIdx.x & (u32)(W - 1); }
template <int W> COOP_FN u32 shfl(u32 x, u32 src) { return __shfl_sync(0xffffffffu, x, (int)src, W); }
template <int W> COOP_FN u32 ballot(bool p) {
  return (__ballot_sync(0xffffffffu, p) >> (threadIdx.x & (u32)(32 - W) & 31u)) & ((1u << W) - 1u);
}
COOP_FN u32 umulhi32(u32 a, u32 b) { return __umulhi(a, b); }
COOP_FN u32 popc32(u32 a) { return (u32)__popc(a); }
#else
u32 emul_lane();
u32 emul_shfl(u32 x, u32 src);
u32 emul_ballot(bool p);
template <int W> inline u32 lane_in_group() { return emul_lane(); }
template <int W> inline u32 shfl(u32 x, u32 src) { return emul_shfl(x, src & (u32)(W - 1)); }
template <int W> inline u32 ballot(bool p) { return emul_ballot(p); }
inline u32 umulhi32(u32 a, u32 b) { return (u32)(((u64)a * (u64)b) >> 32); }
inline u32 popc32(u32 a) { return (u32)__builtin_popcount(a); }
#endif

// per-lane constants, set up once per kernel
struct Lane {
  u32 g;        // lane within the group
  bool low;     // this lane holds a word of every element (Wide16: g < 8; Narrow8: always)
  u32 P;        // word g of p            (0 in lanes 8..15 of Wide16)
  u32 NP;       // word g of W - p
  u32 tc;       // MDS constant of the wrap-bit term this lane evaluates (row-major term index = lane; 0 = none)
  u32 tj;       // its column
  u32 prev;     // source lane of an element carry pass (see Y::carry_in)
  u32 nz;       // Narrow8: 0 in lane 0, all ones elsewhere (lane 0 receives no carry)
  u32 arc;      // lane 7: 2^32 - (p's top word, low CUZK_UNC_WIDEN bits cleared); else 0   (see mds_arc)
  u32 msk[8];   // Narrow8: msk[d] = all ones when d <= g (word d of the row at offset g - d belongs to column g, else g + 8)
  u32 hs;       // Narrow8: shift that selects the carry into this lane's high column
  // source lanes of the transposed sums: word d of the row at offset g - d comes from lane src_e[d] when the rows were made
  // from element words; Wide16 reads lane src_e[d] ^ 8 when they were made from high-half words
  u32 src_e[9];
};

// Keeps a per-lane constant in its register: without this ptxas re-derives the ~27 shuffle source indices from the lane id
// next to every shuffle (two extra ALU instructions each: +13 % instructions, +20 % time measured).
#ifndef CUZK_COOP_PIN
#define CUZK_COOP_PIN 1
#endif
COOP_FN void pin_register(u32 &x) {
#if defined(__CUDA_ARCH__) && CUZK_COOP_PIN
  asm volatile("" : "+r"(x));
#else
  (void)x;
#endif
}

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

// ---- layout-independent building blocks ---------------------------------------------------------------------------------

// quotient floor(x / p) (<= 5) of a 256-bit x from its top word x7.  The top words of 1p..5p are i * D - 1 with
// D = p7 + 1 (checked below), so floor(x / p) = floor(x7 / D) unless x7 + 1 is a multiple of D, where the lower words decide:
// that distance goes to F.near.  The multiply-high division is exact for every 32-bit input up to the flagged ones
// (checked exhaustively, profiles/r02_tuning_notes.md).
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
// The multiplier word goes through a 32-bit register barrier: it usually comes out of a 64-bit carry add and a select, which
// the compiler front end otherwise keeps as a 64-bit value and multiplies with mul.lo.s64 (an IMAD.WIDE plus high-part fix-up
// per product) instead of mul.wide.u32.
#ifndef CUZK_COOP_W32
#define CUZK_COOP_W32 3   // bit 0: barrier in the products by k, bit 1: in the small multiples of p, bit 2: in the first product
#endif
COOP_FN u32 word32(u32 b) {
#if defined(__CUDA_ARCH__) && (CUZK_COOP_W32 != 0)
  asm("" : "+r"(b));
#endif
  return b;
}
COOP_FN void row_mul(u32 (&R)[9], const u32 (&a)[8], u32 b) {
  if (CUZK_COOP_W32 & 4) b = word32(b);
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
  b = word32(b);
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
  b = word32(b);
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

// ---- the two layouts ----------------------------------------------------------------------------------------------------
// Interface of a layout Y:
//   kLanes                        lanes per permutation
//   make_lane()                   per-lane constants
//   product(ew, hw, R, hi, L, F)  R = this lane's row of a product; the rows sit at word offsets 0..7 in the lanes whose
//                                 multiplier word was an element word (hi = false) or a high-half word (hi = true).  Returns
//                                 this lane's normalised word of the low half (ew: an element word, zero where the lane holds
//                                 none) and of the high half (hw: zero where the lane holds none)
//   product_low(R, L)             the un-normalised low-half column of this lane (rows from high-half words, 8 words each)
//   carry_in(c, L)                the carry an element word receives from the word below (lane 0: none; dropped above word 7)
//   wrap_sums(...)                the MDS layer's three wrap counts from the nine wrap bits
struct Wide16 {
  static constexpr int kLanes = 16;
  static constexpr int kMinCtas = 0;    // launch bounds of the kernels of this layout: 0 = unspecified, registers are ptxas's own choice (measured best)
  static COOP_FN Lane make_lane() {
    Lane L = {};
    L.g = lane_in_group<16>();
    L.low = L.g < 8u;
    const u32 m = L.g & 7u;
    L.P = L.low ? pick8(m, CUZK_P0, CUZK_P1, CUZK_P2, CUZK_P3, CUZK_P4, CUZK_P5, CUZK_P6, CUZK_P7) : 0u;
    L.NP = L.low ? pick8(m, CUZK_NP0, CUZK_NP1, CUZK_NP2, CUZK_NP3, CUZK_NP4, CUZK_NP5, CUZK_NP6, CUZK_NP7) : 0u;
    // MDS = [[7,23,8],[26,5,4],[15,20,9]] (poseidon.cpp:46-58), row-major: terms 0..8 in lanes 0..8
    L.tc = L.low ? pick8(m, 7u, 23u, 8u, 26u, 5u, 4u, 15u, 20u) : (L.g == 8u ? 9u : 0u);
    L.tj = L.low ? pick8(m, 0u, 1u, 2u, 0u, 1u, 2u, 0u, 1u) : 2u;
    L.prev = (L.g == 8u) ? 15u : ((L.g - 1u) & 15u);   // lanes 0 and 8 read lane 15, which always holds zero
    L.arc = (L.g == 7u) ? (0u - ((CUZK_P7 >> CUZK_UNC_WIDEN) << CUZK_UNC_WIDEN)) : 0u;
#pragma unroll
    for (int d = 0; d < 9; ++d) {
      L.src_e[d] = (L.g - (u32)d) & 15u;
      pin_register(L.src_e[d]);
    }
    pin_register(L.prev);
    return L;
  }
  // column g of a product whose rows sit in lanes 0..7 (multiplier = element words) or 8..15 (multiplier = high-half words),
  // the row of lane base + j at word offset j; the other half of the group multiplied by zero and contributes zero rows
  static COOP_FN void product(u32 &ew, u32 &hw, const u32 (&R)[9], bool rows_high, const Lane &L, Flags &F) {
    u32 v[9];
#pragma unroll
    for (int d = 0; d < 9; ++d) v[d] = shfl<16>(R[d], rows_high ? (L.src_e[d] ^ 8u) : L.src_e[d]);   // ^ 8: one LOP3, off the critical path
    const u64 T = add3(v[0], v[1], v[2]) + add3(v[3], v[4], v[5]) + add3(v[6], v[7], v[8]);
    const u32 cin = shfl<16>((u32)(T >> 32), L.src_e[1]);   // lane g - 1; column 15 never carries out: lane 0 receives zero
    const u64 r = (u64)(u32)T + (u64)cin;
    F.ovf |= (u32)(r >> 32);
    ew = L.low ? (u32)r : 0u;
    hw = L.low ? 0u : (u32)r;
  }
  static COOP_FN u64 product_low(const u32 (&R)[8], const Lane &L) {
    u32 v[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) v[d] = shfl<16>(R[d], L.low ? (L.src_e[d] ^ 8u) : 0u);   // lanes 8..15 read lane 0, a zero row
    return add3(v[0], v[1], v[2]) + add3(v[3], v[4], v[5]) + add3(v[6], v[7], 0u);
  }
  static COOP_FN i32 carry_in(i32 c, const Lane &L) { return (i32)shfl<16>((u32)c, L.prev); }
  static COOP_FN void wrap_sums(u32 (&wsum)[3], const u32 (&s7)[3], const u32 (&s6)[3], const Lane &L, Flags &F) {
    const u32 a7 = (L.tj == 0u) ? s7[0] : (L.tj == 1u ? s7[1] : s7[2]);
    const u32 a6 = (L.tj == 0u) ? s6[0] : (L.tj == 1u ? s6[1] : s6[2]);
    const u32 bal = ballot<16>(wrap_bit(L.tc, a6, a7, F) != 0u);   // term g in lane g (g = 0..8)
    wsum[0] = popc32(bal & 0x007u);
    wsum[1] = popc32(bal & 0x038u);
    wsum[2] = popc32(bal & 0x1C0u);
  }
};

struct Narrow8 {
  static constexpr int kLanes = 8;
  static constexpr int kMinCtas = 2;    // two CTAs per SM is all 4736 units need; with that stated ptxas takes ~190 registers: 165 -> 162 us
  static COOP_FN Lane make_lane() {
    Lane L = {};
    L.g = lane_in_group<8>();
    L.low = true;
    const u32 m = L.g;
    L.P = pick8(m, CUZK_P0, CUZK_P1, CUZK_P2, CUZK_P3, CUZK_P4, CUZK_P5, CUZK_P6, CUZK_P7);
    L.NP = pick8(m, CUZK_NP0, CUZK_NP1, CUZK_NP2, CUZK_NP3, CUZK_NP4, CUZK_NP5, CUZK_NP6, CUZK_NP7);
    L.tc = pick8(m, 7u, 23u, 8u, 26u, 5u, 4u, 15u, 20u);   // terms 0..7 in lanes 0..7; term 8 (C = 9) is evaluated by every lane
    L.tj = pick8(m, 0u, 1u, 2u, 0u, 1u, 2u, 0u, 1u);
    L.prev = (m - 1u) & 7u;
    L.nz = (m == 0u) ? 0u : 0xffffffffu;
    L.hs = (m == 0u) ? 0u : 8u;
    L.arc = (m == 7u) ? (0u - ((CUZK_P7 >> CUZK_UNC_WIDEN) << CUZK_UNC_WIDEN)) : 0u;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      L.msk[d] = ((u32)d <= m) ? 0xffffffffu : 0u;
      pin_register(L.msk[d]);
    }
#pragma unroll
    for (int d = 0; d < 9; ++d) {
      L.src_e[d] = (m - (u32)d) & 7u;
      pin_register(L.src_e[d]);
    }
    return L;
  }
  // every lane holds a row (at word offset g) and owns the columns g and g + 8: word d of the row at offset g - d belongs to
  // column g when d <= g, to column g + 8 otherwise
  static COOP_FN void product(u32 &ew, u32 &hw, const u32 (&R)[9], bool, const Lane &L, Flags &F) {
    u32 v[9];
#pragma unroll
    for (int d = 0; d < 9; ++d) v[d] = shfl<8>(R[d], L.src_e[d]);
    const u64 all = add3(v[0], v[1], v[2]) + add3(v[3], v[4], v[5]) + add3(v[6], v[7], v[8]);
    const u64 lo = add3(v[0], v[1] & L.msk[1], v[2] & L.msk[2]) + add3(v[3] & L.msk[3], v[4] & L.msk[4], v[5] & L.msk[5]) +
                   add3(v[6] & L.msk[6], v[7] & L.msk[7], 0u);
    const u64 hi = all - lo;
    // one shuffle carries both column carries (each < 16); word 8 receives the carry of column 7
    const u32 pin = shfl<8>((u32)(lo >> 32) | ((u32)(hi >> 32) << 8), L.prev);
    const u64 rl = (u64)(u32)lo + (u64)(pin & 0xffu & L.nz);
    const u64 rh = (u64)(u32)hi + (u64)((pin >> L.hs) & 0xffu);
    F.ovf |= (u32)(rl >> 32) | (u32)(rh >> 32);
    ew = (u32)rl;
    hw = (u32)rh;
  }
  static COOP_FN u64 product_low(const u32 (&R)[8], const Lane &L) {
    u32 v[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) v[d] = shfl<8>(R[d], L.src_e[d]);
    return add3(v[0], v[1] & L.msk[1], v[2] & L.msk[2]) + add3(v[3] & L.msk[3], v[4] & L.msk[4], v[5] & L.msk[5]) +
           add3(v[6] & L.msk[6], v[7] & L.msk[7], 0u);
  }
  static COOP_FN i32 carry_in(i32 c, const Lane &L) { return (i32)(shfl<8>((u32)c, L.prev) & L.nz); }
  static COOP_FN void wrap_sums(u32 (&wsum)[3], const u32 (&s7)[3], const u32 (&s6)[3], const Lane &L, Flags &F) {
    const u32 a7 = (L.tj == 0u) ? s7[0] : (L.tj == 1u ? s7[1] : s7[2]);
    const u32 a6 = (L.tj == 0u) ? s6[0] : (L.tj == 1u ? s6[1] : s6[2]);
    const u32 bal = ballot<8>(wrap_bit(L.tc, a6, a7, F) != 0u);   // terms 0..7
    const u32 w22 = wrap_bit(9u, s6[2], s7[2], F);                 // term 8, on every lane
    wsum[0] = popc32(bal & 0x07u);
    wsum[1] = popc32(bal & 0x38u);
    wsum[2] = popc32(bal & 0xC0u) + w22;
  }
};

// ---- algorithms (shared by the layouts) ---------------------------------------------------------------------------------

// One carry pass over the eight words of an element given as lane values lo + 2^32 * c (c a small signed carry): word g
// becomes lo_g + c_{g-1}; the carry out of word 7 is dropped (arithmetic mod W).  A result outside [0, 2^32) would have to
// ripple on: F.ovf.
template <class Y>
COOP_FN u32 carry_pass(u32 lo, i32 c, const Lane &L, Flags &F) {
  const i64 r = (i64)(u64)lo + (i64)Y::carry_in(c, L);
  F.ovf |= (u32)((u64)r >> 32);
  return (u32)(u64)r;
}

// The same for non-negative carries, but exact whatever ripples: after the neighbour pass a word either generates a carry
// (>= 2^32), propagates one (all ones) or absorbs it; two votes give those sets for the whole element and one integer add
// plays the ripple (the carry into word g is bit g of ((G | P) + G) ^ P).  No flag.  Used where carries are large (the MDS
// layer: up to 2^8, i.e. one unit in ~10^4 would flag on a single pass).
template <class Y>
COOP_FN u32 carry_exact(u32 lo, u32 c, const Lane &L) {
  const u64 r = (u64)lo + (u64)(u32)Y::carry_in((i32)c, L);
  const u32 G = ballot<Y::kLanes>((u32)(r >> 32) != 0u) & 0xffu, P = ballot<Y::kLanes>((u32)r == 0xFFFFFFFFu) & 0xffu;
  const u32 cin = ((G | P) + G) ^ P;
  return (u32)r + (L.low ? ((cin >> L.g) & 1u) : 0u);
}

// all eight words of a word-distributed element
template <class Y>
COOP_FN void gather(u32 (&r)[8], u32 x) {
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = shfl<Y::kLanes>(x, (u32)i);
}

// multiply : field_arithmetic.cpp:221-238 + reduce_512 :250-330 (any 256-bit operands)
//   prod = a*b = high*W + low ;  Mh = high*k = mh*W + ml ;  t = (ml + (mh*k mod W)) mod W ;  hc = mh != 0 ? t mod p : t ;
//   r = ((low + hc) mod W) mod p
// NS independent multiplications are evaluated side by side (the three S-boxes of a full round) so their shuffle
// latencies overlap.  a: replicated operand, b: this lane's word of the other operand (zero where the lane holds none).
template <class Y, int NS>
COOP_FN void mulred(u32 (&r)[NS], const u32 (&a)[NS][8], const u32 (&b)[NS], const Lane &L, Flags &F) {
  constexpr int W = Y::kLanes;
  u32 low[NS], high[NS], ml[NS], mh[NS];
#pragma unroll
  for (int e = 0; e < NS; ++e) {
    u32 R[9];
    row_mul(R, a[e], b[e]);
    Y::product(low[e], high[e], R, false, L, F);
  }
#pragma unroll
  for (int e = 0; e < NS; ++e) {
    u32 R[9];
    row_mul_k(R, high[e]);
    Y::product(ml[e], mh[e], R, true, L, F);
  }
#if CUZK_COOP_TAIL == 2
  // Tail, two top-word rounds: t stays un-normalised, its top word gives e1; y = low + t - e1*p stays un-normalised, its top
  // word gives e2; z = y - e2*p is normalised by the one carry pass.  The carry passes over t and y run for their flags only.
  i64 y[NS];
#pragma unroll
  for (int e = 0; e < NS; ++e) {
    const bool any_mh = ballot<W>(mh[e] != 0u) != 0u;   // the reference reduces t only when mh != 0 (:303)
    u32 R[8];
    row_mul_k_low(R, mh[e]);
    const u64 T = Y::product_low(R, L) + (u64)ml[e];
    const u32 t7 = shfl<W>((u32)T, 7u) + shfl<W>((u32)(T >> 32), 6u);
    (void)carry_pass<Y>((u32)T, (i32)(u32)(T >> 32), L, F);   // flag only: t itself is never normalised
    const u32 q = quot_top(t7, F);
    const u32 e1 = word32(any_mh ? q : 0u);
    y[e] = (i64)(T + (u64)low[e]) - (i64)((u64)e1 * (u64)L.P);
  }
#pragma unroll
  for (int e = 0; e < NS; ++e) {
    const u32 u7 = shfl<W>((u32)(u64)y[e], 7u) + shfl<W>((u32)(y[e] >> 32), 6u);
    (void)carry_pass<Y>((u32)(u64)y[e], (i32)(y[e] >> 32), L, F);   // flag only
    const u32 e2 = word32(quot_top(u7, F));
    const i64 z = y[e] - (i64)((u64)e2 * (u64)L.P);
    r[e] = carry_pass<Y>((u32)(u64)z, (i32)(z >> 32), L, F);
  }
}
#else
  // Tail.  t = (ml + mh*k) mod W stays un-normalised (lane value T < 2^36): only its top word is needed, for the quotient
  // e1 of hc = t mod p.  u = (low + hc) mod W is not formed at all: every lane computes ITS OWN copy of the two top lane
  // values of y = low + t - e1*p from words it fetched in the same round as t's top word, reads off e2 = floor(u / p), and
  // z = low + t - (e1 + e2)*p is normalised by the one carry pass of the tail.  The carry passes over t and y run on the
  // side, for their flags only (they prove the top words exact: by induction from lane 0 every carry used is the true one).
  u32 lw7[NS], lw6[NS];
#pragma unroll
  for (int e = 0; e < NS; ++e) {
    lw7[e] = shfl<W>(low[e], 7u);   // known since the first product: off the critical path
    lw6[e] = shfl<W>(low[e], 6u);
  }
#pragma unroll
  for (int e = 0; e < NS; ++e) {
    const bool any_mh = ballot<W>(mh[e] != 0u) != 0u;   // the reference reduces t only when mh != 0 (:303)
    u32 R[8];
    row_mul_k_low(R, mh[e]);
    const u64 T = Y::product_low(R, L) + (u64)ml[e];
    const u32 tl = (u32)T, tc = (u32)(T >> 32);
    const u32 tl7 = shfl<W>(tl, 7u), tl6 = shfl<W>(tl, 6u), tc6 = shfl<W>(tc, 6u);
    (void)carry_pass<Y>(tl, (i32)tc, L, F);   // flag only: t itself is never normalised
    const u32 q1 = quot_top(tl7 + tc6, F);
    const u32 e1 = word32(any_mh ? q1 : 0u);
    // the two top lane values of y, locally
    const u32 y7lo = lw7[e] + tl7 - e1 * CUZK_P7;
    const i64 y6 = (i64)(((u64)tc6 << 32) | tl6) + (i64)(u64)lw6[e] - (i64)((u64)e1 * (u64)CUZK_P6);
    const u32 e2 = quot_top(y7lo + (u32)(y6 >> 32), F);
    const i64 base = (i64)(T + (u64)low[e]);
    const i64 y = base - (i64)((u64)e1 * (u64)L.P);
    (void)carry_pass<Y>((u32)(u64)y, (i32)(y >> 32), L, F);   // flag only: proves u's top word
    const i64 z = base - (i64)((u64)word32(e1 + e2) * (u64)L.P);
    r[e] = carry_pass<Y>((u32)(u64)z, (i32)(z >> 32), L, F);
  }
}
#endif

// x -> x^5 as the reference does: x2 = x*x, x4 = x2*x2, x5 = x4*x (field_arithmetic.cpp:332-338)
template <class Y, int NS>
COOP_FN void sbox(u32 (&x)[NS], const Lane &L, Flags &F) {
  u32 xr[NS][8], x2r[NS][8], x2[NS], x4[NS], x5[NS];
#pragma unroll
  for (int e = 0; e < NS; ++e) gather<Y>(xr[e], x[e]);
  mulred<Y, NS>(x2, xr, x, L, F);
#pragma unroll
  for (int e = 0; e < NS; ++e) gather<Y>(x2r[e], x2[e]);
  mulred<Y, NS>(x4, x2r, x2, L, F);
  mulred<Y, NS>(x5, xr, x4, L, F);
#pragma unroll
  for (int e = 0; e < NS; ++e) x[e] = x5[e];
}

// add : field_arithmetic.cpp:172-182 for arbitrary 256-bit operands: (a + b) mod W, then the full reduce
template <class Y, int NS>
COOP_FN void add_reduce(u32 (&r)[NS], const u32 (&a)[NS], const u32 (&b)[NS], const Lane &L, Flags &F) {
  constexpr int W = Y::kLanes;
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const u64 y = (u64)a[i] + (u64)b[i];
    const u32 v7 = shfl<W>((u32)y, 7u) + shfl<W>((u32)(y >> 32), 6u);   // top word, ahead of the carry pass
    (void)carry_pass<Y>((u32)y, (i32)(u32)(y >> 32), L, F);              // flag only
    const u32 e = word32(quot_top(v7, F));
    const i64 z = (i64)y - (i64)((u64)e * (u64)L.P);
    r[i] = carry_pass<Y>((u32)(u64)z, (i32)(z >> 32), L, F);
  }
}

// apply_mds_matrix (poseidon.cpp:148-167) in the linear form of mds_row_fast, followed -- when has_rc -- by the next round's
// add_round_constants (:128-134): rc[i] is this lane's word of the constant for state element i (0 above word 1).
template <class Y>
COOP_FN void mds_arc(u32 (&s)[3], const u32 (&rc)[3], bool has_rc, const Lane &L, Flags &F) {
  constexpr int W = Y::kLanes;
  u64 Ls[3];
  Ls[0] = (u64)s[0] * 7u + (u64)s[1] * 23u + (u64)s[2] * 8u;
  Ls[1] = (u64)s[0] * 26u + (u64)s[1] * 5u + (u64)s[2] * 4u;
  Ls[2] = (u64)s[0] * 15u + (u64)s[1] * 20u + (u64)s[2] * 9u;
  u32 s7[3], s6[3], l7[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    s7[j] = shfl<W>(s[j], 7u);
    s6[j] = shfl<W>(s[j], 6u);
  }
  const u32 pk = (u32)(Ls[0] >> 32) | ((u32)(Ls[1] >> 32) << 8) | ((u32)(Ls[2] >> 32) << 16);   // each high part <= 46
  const u32 pk7 = shfl<W>(pk, 7u), pk6 = shfl<W>(pk, 6u);
#pragma unroll
  for (int i = 0; i < 3; ++i) l7[i] = shfl<W>((u32)Ls[i], 7u);
#if CUZK_COOP_MDS_LOCAL_TOP
  u32 l6[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) l6[i] = shfl<W>((u32)Ls[i], 6u);
#endif
  u32 wsum[3];
  Y::wrap_sums(wsum, s7, s6, L, F);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    // quotient estimate from the top of S = sum_j C_ij s_j and the wrap count (mds_row_fast)
    const u32 S7 = l7[i] + ((pk6 >> (8 * i)) & 0xffu);
    const u32 S8 = ((pk7 >> (8 * i)) & 0xffu) + (S7 < l7[i] ? 1u : 0u);
    const u32 a4 = (S8 << 28) | (S7 >> 4);
    const u32 lp = a4 - ((wsum[i] * (CUZK_K7 + 1u) + 15u) >> 4);
    const u32 qhat = umulhi32(lp, kQuotMagic) >> 25;
    const u32 q = word32(qhat - 5u * wsum[i]);
    const u64 y = Ls[i] + (u64)q * (u64)L.NP + (u64)rc[i];            // < 2^40: carries up to 2^8
#if CUZK_COOP_MDS_LOCAL_TOP
    // the top word of y from the two top lane values, computed by every lane itself (the constant has no words above 1):
    // one communication round less than fetching them from lanes 7 and 6 after y is known
    const u64 y7 = ((u64)((pk7 >> (8 * i)) & 0xffu) << 32 | (u64)l7[i]) + (u64)q * (u64)CUZK_NP7;
    const u64 y6 = ((u64)((pk6 >> (8 * i)) & 0xffu) << 32 | (u64)l6[i]) + (u64)q * (u64)CUZK_NP6;
    const u32 v7 = (u32)y7 + (u32)(y6 >> 32);
#else
    const u32 v7 = shfl<W>((u32)y, 7u) + shfl<W>((u32)(y >> 32), 6u);
#endif
#if CUZK_COOP_MDS_EXACT_CARRY
    // v7 is the true top word unless word 6 carries out after its own neighbour pass; the ripple it can receive from below
    // is at most one, so a margin of one on word 6 covers every case: flag only there
    {
      const u64 r6 = (u64)(u32)y + (u64)(u32)Y::carry_in((i32)(u32)(y >> 32), L);
      F.ovf |= (L.g == 6u && r6 >= 0xFFFFFFFFull) ? 1u : 0u;
    }
    const u32 ge = (v7 > CUZK_P7) ? 1u : 0u;                           // y < 2p: one conditional subtraction, decided by the top word
    F.near = umin32(F.near, v7 ^ CUZK_P7);
    const u64 z = y + (u64)(ge ? L.NP : 0u);                           // y - p as y + (W - p): mod W, lane values stay non-negative
    s[i] = carry_exact<Y>((u32)z, (u32)(z >> 32), L);
#else
    (void)carry_pass<Y>((u32)y, (i32)(u32)(y >> 32), L, F);           // flag only
    const u32 ge = (v7 > CUZK_P7) ? 1u : 0u;                           // y < 2p: one conditional subtraction, decided by the top word
    F.near = umin32(F.near, v7 ^ CUZK_P7);
    const i64 z = (i64)y - (i64)(u64)(ge ? L.P : 0u);
    s[i] = carry_pass<Y>((u32)(u64)z, (i32)(z >> 32), L, F);
#endif
    // a state + constant whose top word reaches p's may need the reference's subtraction (arc_fast in poseidon.cuh):
    // lane 7 adds 2^32 - p7 and flags the carry
    if (has_rc) F.ovf |= (u32)(((u64)s[i] + (u64)L.arc) >> 32);
  }
}

// this lane's word of round constant idx (all constants are < 2^64: words 0 and 1)
template <class RcTable>
COOP_FN u32 rc_word(const RcTable &rct, int idx, u32 g) {
  const u32 c0 = rct(idx, 0), c1 = rct(idx, 1);
  return (g == 0u) ? c0 : (g == 1u ? c1 : 0u);
}

// permutation : poseidon.cpp:60-87 on a word-distributed state (any 256-bit values on entry)
template <class Y, class RcTable>
COOP_FN void permute(u32 (&s)[3], const RcTable &rct, const Lane &L, Flags &F) {
  {
    u32 rc[3], t[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) rc[i] = rc_word(rct, i, L.g);
    add_reduce<Y, 3>(t, s, rc, L, F);
#pragma unroll
    for (int i = 0; i < 3; ++i) s[i] = t[i];
  }
#if CUZK_COOP_RC_PREFETCH
  u32 pick0 = (L.g == 0u) ? 0xffffffffu : 0u, pick1 = (L.g == 1u) ? 0xffffffffu : 0u;   // lanes 0 and 1 hold the two words of a constant
  coop::pin_register(pick0);
  coop::pin_register(pick1);
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
    if (full) {
      sbox<Y, 3>(s, L, F);
    } else {
      u32 x[1] = {s[0]};
      sbox<Y, 1>(x, L, F);
      s[0] = x[0];
    }
    u32 rc[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#if CUZK_COOP_RC_PREFETCH
      rc[i] = ((cw[i][0] & pick0) | (cw[i][1] & pick1)) & (has_rc ? 0xffffffffu : 0u);   // masks, not selects: ptxas turns the selects into branches
#else
      rc[i] = has_rc ? rc_word(rct, 3 * (round + 1) + i, L.g) : 0u;
#endif
    }
    mds_arc<Y>(s, rc, has_rc, L, F);
  }
}

// hash_multiple / sponge over `width` inputs (poseidon.cpp:98-126): out = this lane's word of the digest.
// load(i) returns this lane's word of input i (zero where the lane holds none).  Returns the group's vote: non-zero =
// evaluate this unit again exactly.
template <class Y, class RcTable, class Loader>
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
      add_reduce<Y, 2>(r, a, x, L, F);
      s[1] = r[0];
      s[2] = r[1];
    } else {
      u32 a[1] = {s[1]}, x[1] = {load(i)}, r[1];
      add_reduce<Y, 1>(r, a, x, L, F);
      s[1] = r[0];
    }
    permute<Y>(s, rct, L, F);
  }
  out = s[1];
  return ballot<Y::kLanes>(flagged(F));
}

}  // namespace coop
}  // namespace cuzk
