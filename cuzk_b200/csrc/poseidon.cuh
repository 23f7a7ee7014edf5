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
// The sum is < p + 2^64, so it can only reach p when its top word equals p's top word: the exact
// compare-and-subtract runs on that (2^-30) branch only.
__device__ __forceinline__ void arc_canon(u32 (&s)[8], int idx) {
  const u32 c0 = c_rc[idx][0], c1 = c_rc[idx][1];
  s[0] = add_cc(s[0], c0);
  s[1] = addc_cc(s[1], c1);
#pragma unroll
  for (int i = 2; i < 7; ++i) s[i] = addc_cc(s[i], 0u);
  s[7] = addc(s[7], 0u);
  if (s[7] >= CUZK_P7) cond_sub_mp<1>(s);
}

// fast path: no branch; a sum whose top word reaches p's top word (2^-30) is left unreduced and reported in `unc`
__device__ __forceinline__ void arc_fast(u32 (&s)[8], int idx, u32 &unc) {
  const u32 c0 = c_rc[idx][0], c1 = c_rc[idx][1];
  s[0] = add_cc(s[0], c0);
  s[1] = addc_cc(s[1], c1);
#pragma unroll
  for (int i = 2; i < 7; ++i) s[i] = addc_cc(s[i], 0u);
  s[7] = addc(s[7], 0u);
  unc |= ((s[7] >> CUZK_UNC_WIDEN) >= (CUZK_P7 >> CUZK_UNC_WIDEN)) ? 1u : 0u;
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

// exact MDS layer, term by term as the reference evaluates it (also the fallback of mds_fast)
__device__ __forceinline__ void mds_exact(u32 (&s0)[8], u32 (&s1)[8], u32 (&s2)[8]) {
  u32 n0[8], n1[8], n2[8];
  mds_row<7, 23, 8>(n0, s0, s1, s2);
  mds_row<26, 5, 4>(n1, s0, s1, s2);
  mds_row<15, 20, 9>(n2, s0, s1, s2);
#pragma unroll
  for (int i = 0; i < 8; ++i) { s0[i] = n0[i]; s1[i] = n1[i]; s2[i] = n2[i]; }
}

// ---- fast MDS layer -----------------------------------------------------------------------------
// With x = C*s = h*W + low (W = 2^256, h <= 4), the reference term is
//     mul(C, s) = red(u),  u = (low + h*k) mod W = x - h*(W - k) - w*W,   w = [low + h*k >= W]
// and W - k = 5p, so  u == x - w*k (mod p)  and the three-term row is
//     n_i = (S_i - Wsum_i * k) mod p,   S_i = sum_j C_ij * s_j  (< 47p, nine words),  Wsum_i = sum_j w_ij.
// S_i needs no carry handling (each 64-bit lane holds a sum < 2^39), the reduction is one quotient
// estimate q^ in {q-1, q} from the top words, one multiply-add by (W - p), and one conditional subtract:
//     y = S_i - Wsum_i*k - q^*p  ==  S_low + (q^ - 5*Wsum_i) * (W - p)   (mod W),   0 <= y < 2p.
// The wrap bit w_ij = [low >= W - h*k] is decided from the top word of low; the two cases where the top
// word cannot decide (carry into word 7 ambiguous, or low_7 equal to the threshold word; ~2^-27 per term)
// set `unc` and the caller recomputes the layer with mds_exact.
// wrap bit of one term: adds w to wsum, ORs the "cannot decide" conditions into unc
template <u32 C>
__device__ __forceinline__ void mds_wrap_bit(u32 &wsum, u32 &unc, const u32 (&s)[8]) {
  const u64 y = (u64)s[6] * C;                       // word 6 product: its high half carries into word 7
  const u64 z = (u64)s[7] * C + (y >> 32);           // (h : low_7) unless the carry out of word 6 is ambiguous
  const u32 h = (u32)(z >> 32), low7 = (u32)z;
  unc |= ((u32)y >= 0xFFFFFFE0u) ? 1u : 0u;          // lower words could still push a carry into word 7
  const u32 fl = h * CUZK_K7 + ((h * 5u) >> 3);      // floor(h*k / 2^224) for h = 0..4
  const u32 t = add_cc(low7, fl);                    // low_7 > 0xFFFFFFFF - fl  <=>  carry out
  wsum = addc(wsum, 0u);
  unc |= (t == 0xFFFFFFFFu) ? 1u : 0u;               // low_7 equals the threshold word: lower words decide
}

template <u32 C0, u32 C1, u32 C2>
__device__ __forceinline__ void mds_row_fast(u32 (&n)[8], u32 &unc, const u32 (&s0)[8], const u32 (&s1)[8], const u32 (&s2)[8]) {
  u32 wsum = 0;
  mds_wrap_bit<C0>(wsum, unc, s0);
  mds_wrap_bit<C1>(wsum, unc, s1);
  mds_wrap_bit<C2>(wsum, unc, s2);
  // S = sum_j C_j * s_j on even/odd lanes (no carries: every lane < 3 * 26 * 2^32)
  u64 e[4], o[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    e[m] = (u64)s0[2 * m] * C0 + (u64)s1[2 * m] * C1 + (u64)s2[2 * m] * C2;
    o[m] = (u64)s0[2 * m + 1] * C0 + (u64)s1[2 * m + 1] * C1 + (u64)s2[2 * m + 1] * C2;
  }
  u32 S[9];
  S[0] = (u32)e[0];
  S[1] = add_cc((u32)(e[0] >> 32), (u32)o[0]);
  S[2] = addc_cc((u32)e[1], (u32)(o[0] >> 32));
  S[3] = addc_cc((u32)(e[1] >> 32), (u32)o[1]);
  S[4] = addc_cc((u32)e[2], (u32)(o[1] >> 32));
  S[5] = addc_cc((u32)(e[2] >> 32), (u32)o[2]);
  S[6] = addc_cc((u32)e[3], (u32)(o[2] >> 32));
  S[7] = addc_cc((u32)(e[3] >> 32), (u32)o[3]);
  S[8] = addc(0u, (u32)(o[3] >> 32));
  // quotient estimate: L' <= floor((S - wsum*k) / 2^228), q^ = floor(L' * floor(2^61/(p7+1)) / 2^57)
  const u32 a4 = (S[8] << 28) | (S[7] >> 4);
  const u32 lp = a4 - ((wsum * (CUZK_K7 + 1u) + 15u) >> 4);
  const u32 qhat = __umulhi(lp, kQuotMagic) >> 25;
  const u32 q = qhat - 5u * wsum;                    // >= 0: every wrap puts >= 1.89 W into S - wsum*k
  // y = S_low + q * (W - p)  (mod W)
  u32 ye[8], yo[9];
#pragma unroll
  for (int i = 0; i < 8; ++i) ye[i] = S[i];
#pragma unroll
  for (int i = 0; i < 9; ++i) yo[i] = 0;
  ye[0] = mad_lo_cc(q, np_limb(0), ye[0]);
  ye[1] = madc_hi_cc(q, np_limb(0), ye[1]);
#pragma unroll
  for (int m = 1; m < 4; ++m) {
    ye[2 * m] = madc_lo_cc(q, np_limb(2 * m), ye[2 * m]);
    ye[2 * m + 1] = madc_hi_cc(q, np_limb(2 * m), ye[2 * m + 1]);
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) mad_wide(yo[2 * m + 1], yo[2 * m + 2], q, np_limb(2 * m + 1));
  n[0] = ye[0];
  n[1] = add_cc(ye[1], yo[1]);
#pragma unroll
  for (int i = 2; i < 7; ++i) n[i] = addc_cc(ye[i], yo[i]);
  n[7] = addc(ye[7], yo[7]);
  cond_sub_top<1>(n, unc);
}

// ---- the same row on the FP64 pipe ---------------------------------------------------------------------------
// B200 issues DFMA at full rate (62 per SM per clock measured, csrc/microbench.cu variant 9) on a pipe the rest of the
// permutation never touches, while the 32x32->64 multiplier (IMAD.WIDE, ~28 per SM per clock) is the kernel's bound.
// Every product of this row has a factor below 2^6, so all of them are exact in double precision:
//   lane sums   S_m = C0*s0[m] + C1*s1[m] + C2*s2[m]          < 2^39
//   y_m         = S_m + q * (W - p)[m]                         < 2^40        (q < 2^6)
// Limbs enter as doubles through the 2^52 bias trick (no I2F), lanes leave the same way; the quotient estimate is
// evaluated in double with round-toward-zero so that it never exceeds floor(X / p), X = S - wsum * k, and is at most
// one below it (the dropped terms lower the estimate by < 1e-7 quotient units).
#ifndef CUZK_MDS_FP64
#define CUZK_MDS_FP64 1
#endif
#define CUZK_TWO52 4503599627370496.0
#ifndef CUZK_FP64_CVT
#define CUZK_FP64_CVT 0   // 0: 2^52 bias trick (register-pair moves + DADD); 1: I2F / F2I conversion instructions
#endif
__device__ __forceinline__ double u32_as_double(u32 x) {
#if CUZK_FP64_CVT
  return (double)x;
#else
  return __hiloint2double(0x43300000, (int)x) - CUZK_TWO52;
#endif
}

template <u32 C0, u32 C1, u32 C2>
__device__ __forceinline__ void mds_row_fp64(u32 (&n)[8], u32 &unc, const u32 (&s0)[8], const u32 (&s1)[8], const u32 (&s2)[8]) {
  u32 wsum = 0;
  mds_wrap_bit<C0>(wsum, unc, s0);
  mds_wrap_bit<C1>(wsum, unc, s1);
  mds_wrap_bit<C2>(wsum, unc, s2);
  double S[8];
#pragma unroll
  for (int m = 0; m < 8; ++m)
    S[m] = fma((double)C0, u32_as_double(s0[m]), fma((double)C1, u32_as_double(s1[m]), (double)C2 * u32_as_double(s2[m])));
  // quotient estimate: v <= S / 2^224, L <= X / 2^228, qhat = floor(L * c) with c <= 2^228 / p  =>  q - 1 <= qhat <= q
  const double wd = u32_as_double(wsum);
  const double v = __fma_rz(S[6], 0x1p-32, S[7]);
  const double L = __fma_rz(v, 0.0625, -(wd * 14722940.125));            // (K7 + 1) / 16 = 14722940.125 exactly
  const double qhat = __fma_rz(L, 0x1.5291d188b15fap-26, CUZK_TWO52) - CUZK_TWO52;   // 16 / (p7 + 1), rounded down
  const double q = fma(-5.0, wd, qhat);                                   // >= 0 (see mds_row_fast)
  // y = S + q * (W - p)  (mod W): lanes back to integers, one carry chain
  u32 lo[8], hi[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) {
#if CUZK_FP64_CVT
    const u64 y = __double2ull_rz(fma(q, (double)np_limb(m), S[m]));
    lo[m] = (u32)y;
    hi[m] = (u32)(y >> 32);
#else
    const double y = fma(q, (double)np_limb(m), S[m]) + CUZK_TWO52;
    lo[m] = (u32)__double2loint(y);
    hi[m] = (u32)__double2hiint(y) & 0x000FFFFFu;
#endif
  }
  n[0] = lo[0];
  n[1] = add_cc(lo[1], hi[0]);
#pragma unroll
  for (int i = 2; i < 7; ++i) n[i] = addc_cc(lo[i], hi[i - 1]);
  n[7] = addc(lo[7], hi[6]);
  cond_sub_top<1>(n, unc);
}

// fast MDS layer: undecided wrap bits / comparisons are reported in `unc`, the state is then meaningless and the caller
// recomputes its unit on the exact path
__device__ __forceinline__ void mds_fast(u32 (&s0)[8], u32 (&s1)[8], u32 (&s2)[8], u32 &unc) {
  u32 n0[8], n1[8], n2[8];
#if CUZK_MDS_FP64
  mds_row_fp64<7, 23, 8>(n0, unc, s0, s1, s2);
  mds_row_fp64<26, 5, 4>(n1, unc, s0, s1, s2);
  mds_row_fp64<15, 20, 9>(n2, unc, s0, s1, s2);
#else
  mds_row_fast<7, 23, 8>(n0, unc, s0, s1, s2);
  mds_row_fast<26, 5, 4>(n1, unc, s0, s1, s2);
  mds_row_fast<15, 20, 9>(n2, unc, s0, s1, s2);
#endif
#pragma unroll
  for (int i = 0; i < 8; ++i) { s0[i] = n0[i]; s1[i] = n1[i]; s2[i] = n2[i]; }
}

// one layer with its own fallback (test hook cuzk_debug_mds_layer only)
__device__ __forceinline__ void mds(u32 (&s0)[8], u32 (&s1)[8], u32 (&s2)[8]) {
  u32 t0[8], t1[8], t2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { t0[i] = s0[i]; t1[i] = s1[i]; t2[i] = s2[i]; }
  u32 unc = 0;
  mds_fast(t0, t1, t2, unc);
  if (unc != 0) {
    mds_exact(s0, s1, s2);
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { s0[i] = t0[i]; s1[i] = t1[i]; s2[i] = t2[i]; }
}

template <bool EXACT>
__device__ __forceinline__ void sbox(u32 (&s)[8], u32 &unc) {
  u32 r[8];
  fr_pow5_t<EXACT>(r, s, unc);
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = r[i];
}

// permutation : poseidon.cpp:60-87.  CANON = every state word is already < p on entry.
// EXACT = false is the production path: top-word reductions, linear-form MDS; any undecided comparison sets `unc` and the
// caller recomputes the unit with EXACT = true (the reference's evaluation order, step by step).
template <bool CANON, bool EXACT>
__device__ __forceinline__ void permute_t(u32 (&s0)[8], u32 (&s1)[8], u32 (&s2)[8], u32 &unc) {
#pragma unroll 1
  for (int round = 0; round < kRounds; ++round) {
    if (!CANON && round == 0) {
      arc_general(s0, 0);
      arc_general(s1, 1);
      arc_general(s2, 2);
    } else if (EXACT) {
      arc_canon(s0, 3 * round);
      arc_canon(s1, 3 * round + 1);
      arc_canon(s2, 3 * round + 2);
    } else {
      arc_fast(s0, 3 * round, unc);
      arc_fast(s1, 3 * round + 1, unc);
      arc_fast(s2, 3 * round + 2, unc);
    }
    const bool full = (round < kFullHalf) || (round >= kFullHalf + kPartial);
    const int nsbox = full ? 3 : 1;
    // one S-box body shared by all rounds: apply to s0, and in full rounds rotate the state so
    // the next element sits in s0 (three rotations restore the order).
#pragma unroll 1
    for (int i = 0; i < nsbox; ++i) {
      sbox<EXACT>(s0, unc);
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
    if (EXACT) mds_exact(s0, s1, s2);
    else mds_fast(s0, s1, s2, unc);
  }
}

// exact permutation, out of line: one copy per module, reached only from the rare fallback paths
__device__ unsigned long long g_exact_fallbacks;   // how many units took the exact path (cuzk_debug_fallback_count)
__device__ __noinline__ void permute_exact(u32 *st, int canon) {
  u32 s0[8], s1[8], s2[8], unused = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s0[i] = st[i]; s1[i] = st[8 + i]; s2[i] = st[16 + i]; }
  if (canon) permute_t<true, true>(s0, s1, s2, unused);
  else permute_t<false, true>(s0, s1, s2, unused);
#pragma unroll
  for (int i = 0; i < 8; ++i) { st[i] = s0[i]; st[8 + i] = s1[i]; st[16 + i] = s2[i]; }
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

// hash_multiple / device sponge over `width` inputs (poseidon.cpp:98-126; domain separator ds_hi:ds_lo):
// absorbs two per permutation, a final odd input alone; width == 0 -> zero permutations -> output 0.
// `load(x, i)` must be repeatable: when the fast pass reports an undecided comparison the whole sponge is evaluated
// again on the exact path.
// the whole sponge on the exact path (the reference's evaluation order, step by step): the fallback of every fast path
template <class Loader>
__device__ __forceinline__ void sponge_exact(u32 (&out)[8], u32 ds_lo, u32 ds_hi, int width, Loader load) {
  atomicAdd(&g_exact_fallbacks, 1ull);
  u32 st[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) st[i] = 0;
  st[0] = ds_lo;
  st[1] = ds_hi;
#pragma unroll 1
  for (int i = 0; i < width; i += 2) {
    u32 x[8], a[8];
    load(x, i);
#pragma unroll
    for (int w = 0; w < 8; ++w) a[w] = st[8 + w];
    absorb(a, x);
#pragma unroll
    for (int w = 0; w < 8; ++w) st[8 + w] = a[w];
    if (i + 1 < width) {
      load(x, i + 1);
#pragma unroll
      for (int w = 0; w < 8; ++w) a[w] = st[16 + w];
      absorb(a, x);
#pragma unroll
      for (int w = 0; w < 8; ++w) st[16 + w] = a[w];
    }
    permute_exact(st, 1);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) out[i] = st[8 + i];
}

template <class Loader>
__device__ __forceinline__ void sponge_n(u32 (&out)[8], u32 ds_lo, u32 ds_hi, int width, Loader load) {
  u32 unc = 0;
  {
    u32 s0[8], s1[8], s2[8];
    set_small(s0, ds_lo);
    s0[1] = ds_hi;
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
      permute_t<true, false>(s0, s1, s2, unc);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = s1[i];
  }
  if (unc != 0) sponge_exact(out, ds_lo, ds_hi, width, load);   // ~1e-6 per permutation on random data
}
template <class Loader>
__device__ __forceinline__ void sponge_n(u32 (&out)[8], u32 ds, int width, Loader load) {
  sponge_n(out, ds, 0u, width, load);
}

}  // namespace cuzk
