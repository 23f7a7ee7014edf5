// coop_kernels.cuh -- the __global__ kernels of the cooperative path (coop.cuh): a group of sixteen (layout Wide16) or eight
// (Narrow8) lanes per unit, for launches too narrow to fill the chip with one-thread-per-unit CTAs.  Same arguments and results
// as their one-thread twins in kernels.cuh; the launchers (host_path.cuh, cuzk_kernels.cu) choose by unit count (coop_kind()).
//
// Rules that keep the full-mask shuffles legal: every lane of a warp runs the same sequence of shuffles whatever its
// unit is (units past the end are clamped to the last one and only their stores are skipped); no loop exits early; the
// rare exact-path repairs synchronise their own group only.
#pragma once
#include "coop.cuh"
#include "coop16.cuh"

// launch bounds of the templated kernels: the CTA size and, per layout, how many CTAs per SM the register allocation must
// allow.  Stating "2" makes ptxas take ~190 registers and no spills, which helps the eight-lane kernels (165 -> 162 us) and hurts
// the sixteen-lane ones (86 -> 95 us: the scheduling lottery of profiles/r02_tuning_notes.md section 3 again), hence per layout.
#ifndef CUZK_COOP_MIN_BLOCKS
#define CUZK_COOP_MIN_BLOCKS 0   // > 0 overrides both layouts (tuning knob)
#endif
#if CUZK_COOP_MIN_BLOCKS > 0
#define CUZK_COOP_BOUNDS __launch_bounds__(kCoopBlock, CUZK_COOP_MIN_BLOCKS)
#else
#define CUZK_COOP_BOUNDS __launch_bounds__(kCoopBlock, Y::kMinCtas)
#endif
constexpr int kCoopBlock = 128;                       // 8 (Wide16) or 16 (Narrow8) units per CTA, one warp per SM sub-partition
using coop::Narrow8;
using coop::Wide16;

// which implementation serves a layout: Narrow8 the shared templates of coop.cuh, Wide16 the non-templated coop16.cuh (same
// arithmetic; ptxas schedules it 14 % faster, see the header of coop16.cuh)
template <class Y> struct Algo;
template <> struct Algo<Narrow8> {
  typedef coop::Lane Lane;
  typedef coop::Flags Flags;
  static __device__ __forceinline__ Lane make_lane() { return Narrow8::make_lane(); }
  template <class Rc, class Ld>
  static __device__ __forceinline__ u32 sponge(u32 &out, u32 ds_lo, u32 ds_hi, int width, const Rc &rct, const Lane &L, Ld load) {
    return coop::sponge<Narrow8>(out, ds_lo, ds_hi, width, rct, L, load);
  }
  template <class Rc>
  static __device__ __forceinline__ void permute(u32 (&s)[3], const Rc &rct, const Lane &L, Flags &F) { coop::permute<Narrow8>(s, rct, L, F); }
  static __device__ __forceinline__ bool flagged(const Flags &F) { return coop::flagged(F); }
};
template <> struct Algo<Wide16> {
  typedef coop16::Lane Lane;
  typedef coop16::Flags Flags;
  static __device__ __forceinline__ Lane make_lane() { return coop16::make_lane(); }
  template <class Rc, class Ld>
  static __device__ __forceinline__ u32 sponge(u32 &out, u32 ds_lo, u32 ds_hi, int width, const Rc &rct, const Lane &L, Ld load) {
    return coop16::sponge(out, ds_lo, ds_hi, width, rct, L, load);
  }
  template <class Rc>
  static __device__ __forceinline__ void permute(u32 (&s)[3], const Rc &rct, const Lane &L, Flags &F) { coop16::permute(s, rct, L, F); }
  static __device__ __forceinline__ bool flagged(const Flags &F) { return coop16::flagged(F); }
};

struct RcConst {
  __device__ __forceinline__ u32 operator()(int idx, int w) const { return c_rc[idx][w]; }
};

__device__ __forceinline__ const u32 *words_of(const uint4 *p) { return reinterpret_cast<const u32 *>(p); }
// this lane's word of the element at p (lanes 8..15 of a group hold zero)
template <class LaneT> __device__ __forceinline__ u32 ld_word(const LaneT &L, const uint4 *p) { return L.low ? words_of(p)[L.g] : 0u; }
template <class LaneT> __device__ __forceinline__ u32 ldg_word(const LaneT &L, const uint4 *p) { return L.low ? __ldg(words_of(p) + L.g) : 0u; }
template <class LaneT> __device__ __forceinline__ void st_word(const LaneT &L, uint4 *p, u32 w) {
  if (L.low) reinterpret_cast<u32 *>(p)[L.g] = w;
}
template <class Y> __device__ __forceinline__ size_t coop_unit() { return ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / Y::kLanes; }
template <class Y> __device__ __forceinline__ unsigned coop_group_mask() {
  return ((1u << Y::kLanes) - 1u) << (threadIdx.x & (unsigned)(32 - Y::kLanes) & 31u);
}

// sponge of one unit, result stored as 8 words at `dst` (lane g < 8 stores word g; after an `unc` vote lane 0 evaluates the
// unit again on the exact one-thread path).  loadw(i): this lane's word of input i; loadx(x, i): the whole input i.
template <class Y, class LW, class LX>
__device__ __forceinline__ void coop_sponge_store(uint4 *dst, bool active, u32 ds_lo, u32 ds_hi, int width, const typename Algo<Y>::Lane &L, LW loadw,
                                                  LX loadx) {
  u32 w;
  const u32 vote = Algo<Y>::sponge(w, ds_lo, ds_hi, width, RcConst(), L, loadw);
  if (!active) return;
  if (vote) {
    if (L.g == 0u) {
      u32 r[8];
      sponge_exact(r, ds_lo, ds_hi, width, loadx);
      store_fr(dst, r);
    }
  } else {
    st_word(L, dst, w);
  }
}

template <class Y>
__global__ void CUZK_COOP_BOUNDS coop_hash_pairs_kernel(const uint4 *__restrict__ l, const uint4 *__restrict__ r,
                                                                      uint4 *__restrict__ out, size_t n) {
  const typename Algo<Y>::Lane L = Algo<Y>::make_lane();
  size_t i = coop_unit<Y>();
  const bool active = i < n;
  if (!active) i = n - 1;
  coop_sponge_store<Y>(
      out + 2 * i, active, 2u, 0u, 2, L, [&](int j) { return ldg_word(L, (j == 0 ? l : r) + 2 * i); },
      [&](u32(&x)[8], int j) { load_fr(x, (j == 0 ? l : r) + 2 * i); });
}

template <class Y>
__global__ void CUZK_COOP_BOUNDS coop_hash_single_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t n) {
  const typename Algo<Y>::Lane L = Algo<Y>::make_lane();
  size_t i = coop_unit<Y>();
  const bool active = i < n;
  if (!active) i = n - 1;
  coop_sponge_store<Y>(
      out + 2 * i, active, 1u, 0u, 1, L, [&](int) { return ldg_word(L, in + 2 * i); }, [&](u32(&x)[8], int) { load_fr(x, in + 2 * i); });
}

template <class Y>
__global__ void CUZK_COOP_BOUNDS coop_sponge_kernel(const uint4 *__restrict__ in, int width, u32 ds_lo, u32 ds_hi,
                                                                  uint4 *__restrict__ out, size_t n) {
  const typename Algo<Y>::Lane L = Algo<Y>::make_lane();
  size_t i = coop_unit<Y>();
  const bool active = i < n;
  if (!active) i = n - 1;
  const uint4 *base = in + 2 * i * (size_t)width;
  coop_sponge_store<Y>(
      out + 2 * i, active, ds_lo, ds_hi, width, L, [&](int j) { return ldg_word(L, base + 2 * j); },
      [&](u32(&x)[8], int j) { load_fr(x, base + 2 * j); });
}

// batch_permutation: in place, caller-supplied (possibly non-canonical) states
template <class Y>
__global__ void CUZK_COOP_BOUNDS coop_permutation_kernel(uint4 *states, size_t n) {
  const typename Algo<Y>::Lane L = Algo<Y>::make_lane();
  size_t i = coop_unit<Y>();
  const bool active = i < n;
  if (!active) i = n - 1;
  uint4 *st = states + 6 * i;
  u32 s[3] = {ld_word(L, st), ld_word(L, st + 2), ld_word(L, st + 4)};
  typename Algo<Y>::Flags F;
  Algo<Y>::permute(s, RcConst(), L, F);
  const u32 vote = coop::ballot<Y::kLanes>(Algo<Y>::flagged(F));
  if (!active) return;
  if (vote) {
    if (L.g == 0u) {
      atomicAdd(&g_exact_fallbacks, 1ull);
      u32 *sw = reinterpret_cast<u32 *>(st);
      u32 full[24];
#pragma unroll
      for (int w = 0; w < 24; ++w) full[w] = sw[w];
      permute_exact(full, 0);
#pragma unroll
      for (int w = 0; w < 24; ++w) sw[w] = full[w];
    }
  } else {
    st_word(L, st, s[0]);
    st_word(L, st + 2, s[1]);
    st_word(L, st + 4, s[2]);
  }
}

// one Merkle level (see merkle_level_kernel): unit = (tree, node)
template <class Y>
__global__ void CUZK_COOP_BOUNDS coop_merkle_level_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t in_real,
                                                                        size_t out_count, int arity, const uint4 *__restrict__ pad_in,
                                                                        const uint4 *__restrict__ pad_out, size_t ntrees, size_t tree_stride) {
  const typename Algo<Y>::Lane L = Algo<Y>::make_lane();
  size_t t = coop_unit<Y>();
  const bool active = t < out_count * ntrees;
  if (!active) t = out_count * ntrees - 1;
  const size_t tree = t / out_count, i = t - tree * out_count;
  in += 2 * tree * tree_stride;
  out += 2 * tree * tree_stride;
  const size_t first = i * (size_t)arity;
  const bool padding = first >= in_real;   // no real child: the node is the padding constant of the output level
  u32 w;
  const u32 vote = Algo<Y>::sponge(w, 3u, 0u, arity, RcConst(), L, [&](int j) {
    const uint4 *src = (first + j < in_real) ? in + 2 * (first + j) : pad_in;
    return ld_word(L, src);
  });
  if (!active) return;
  if (padding) {
    st_word(L, out + 2 * i, ld_word(L, pad_out));
  } else if (vote) {
    if (L.g == 0u) {
      u32 r[8];
      sponge_exact(r, 3u, 0u, arity, [&](u32(&x)[8], int j) {
        const uint4 *src = (first + j < in_real) ? in + 2 * (first + j) : pad_in;
        load_fr_plain(x, src);
      });
      store_fr(out + 2 * i, r);
    }
  } else {
    st_word(L, out + 2 * i, w);
  }
}

// incremental update, one level (see merkle_update_level_kernel)
template <class Y>
__global__ void CUZK_COOP_BOUNDS coop_merkle_update_level_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out,
                                                                               const u64 *__restrict__ indices, size_t count, size_t n,
                                                                               u64 divisor, int arity) {
  const typename Algo<Y>::Lane L = Algo<Y>::make_lane();
  size_t q = coop_unit<Y>();
  bool active = q < count;
  if (!active) q = count - 1;
  u64 idx = indices[q];
  if (idx >= n) {   // out of range: skipped (the lanes stay in step on leaf 0)
    active = false;
    idx = 0;
  }
  const size_t node = idx / divisor;
  const uint4 *kids = in + 2 * node * (size_t)arity;
  coop_sponge_store<Y>(
      out + 2 * node, active, 3u, 0u, arity, L, [&](int j) { return ld_word(L, kids + 2 * j); },
      [&](u32(&x)[8], int j) { load_fr_plain(x, kids + 2 * j); });
}

// One sponge whose inputs may include the group's running value `cur` (word-distributed): on an `unc` vote the level is
// evaluated again by lane 0 on the exact path, through a shared-memory slot of the group.  load_full(x, j, full): whole
// input j, `full` = the eight words of `cur`.
template <class Y, class LW, class LXF>
__device__ __forceinline__ u32 coop_sponge_chained(u32 cur, u32 *slot, int width, const typename Algo<Y>::Lane &L, LW loadw, LXF load_full) {
  u32 w;
  const u32 vote = Algo<Y>::sponge(w, 3u, 0u, width, RcConst(), L, loadw);
  if (vote) {   // group-uniform, rare
    const unsigned mask = coop_group_mask<Y>();
    if (L.low) slot[L.g] = cur;
    __syncwarp(mask);
    if (L.g == 0u) {
      u32 full[8], r[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) full[k] = slot[k];
      sponge_exact(r, 3u, 0u, width, [&](u32(&x)[8], int j) { load_full(x, j, full); });
#pragma unroll
      for (int k = 0; k < 8; ++k) slot[k] = r[k];
    }
    __syncwarp(mask);
    w = L.low ? slot[L.g] : 0u;
    __syncwarp(mask);
  }
  return w;
}

// batch verification (see merkle_verify_kernel): one group per proof, levels in sequence
template <class Y>
__global__ void CUZK_COOP_BOUNDS coop_merkle_verify_kernel(const uint4 *__restrict__ leaves, const uint4 *__restrict__ sib,
                                                                         const u32 *__restrict__ pos, int nlv, int arity,
                                                                         const uint4 *__restrict__ root, uint4 root_lo, uint4 root_hi,
                                                                         uint8_t *__restrict__ results, size_t num_proofs) {
  __shared__ u32 s_fix[kCoopBlock / Y::kLanes][8];
  const typename Algo<Y>::Lane L = Algo<Y>::make_lane();
  size_t q = coop_unit<Y>();
  const bool active = q < num_proofs;
  if (!active) q = num_proofs - 1;
  u32 *slot = s_fix[threadIdx.x / Y::kLanes];
  u32 cur = ldg_word(L, leaves + 2 * q);
  bool ok = true;
#pragma unroll 1
  for (int l = 0; l < nlv; ++l) {
    u32 my = pos[q * (size_t)nlv + l];
    if (my >= (u32)arity) {   // malformed proof: keep the lanes in step on a clamped position, remember the verdict
      ok = false;
      my = 0;
    }
    const uint4 *sb = sib + 2 * (q * (size_t)nlv + l) * (size_t)(arity - 1);
    const u32 c = cur;
    cur = coop_sponge_chained<Y>(
        c, slot, arity, L,
        [&](int j) {
          const int k = (j < (int)my) ? j : j - 1;
          const u32 v = ldg_word(L, sb + 2 * (j == (int)my ? 0 : k));
          return (j == (int)my) ? c : v;
        },
        [&](u32(&x)[8], int j, const u32(&full)[8]) {
          if (j == (int)my) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = full[k];
          } else {
            load_fr(x, sb + 2 * (j < (int)my ? j : j - 1));
          }
        });
  }
  u32 rt;
  if (root) {
    rt = ldg_word(L, root);
  } else {
    rt = coop::pick8(L.g & 7u, root_lo.x, root_lo.y, root_lo.z, root_lo.w, root_hi.x, root_hi.y, root_hi.z, root_hi.w);
    rt = L.low ? rt : 0u;
  }
  const u32 differ = coop::ballot<Y::kLanes>(rt != cur);
  if (active && L.g == 0u) results[q] = (ok && differ == 0u) ? 1 : 0;
}

// padding chain (see padding_chain_kernel): one warp, both of its groups run the same chain, group 0 stores
__global__ void __launch_bounds__(32) coop_padding_chain_kernel(uint4 *pad, int arity, int start, int end) {
  __shared__ u32 s_fix[2][8];
  typedef Wide16 Y;
  const typename Algo<Y>::Lane L = Algo<Y>::make_lane();
  const u32 grp = threadIdx.x >> 4;
  u32 cur = (start > 0) ? ld_word(L, pad + 2 * (start - 1)) : 0u;
  for (int l = start; l < end; ++l) {
    const u32 c = cur;
    cur = coop_sponge_chained<Y>(
        c, s_fix[grp], arity, L, [&](int) { return c; },
        [&](u32(&x)[8], int, const u32(&full)[8]) {
#pragma unroll
          for (int k = 0; k < 8; ++k) x[k] = full[k];
        });
    if (grp == 0u) st_word(L, pad + 2 * l, cur);
  }
}
