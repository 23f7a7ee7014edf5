// kernels.cuh -- every __global__ kernel of libcuzk_b200.so (included once, by cuzk_kernels.cu).
//
// One thread evaluates one unit (field op, permutation chain, Merkle node or proof): the work is ~106 k instructions per
// permutation against <= 256 bytes of traffic, so the kernels are bound by the integer multiplier (IMAD.WIDE), not by HBM;
// see DESIGN.md section 4 for the roofline.
#pragma once
#include "poseidon.cuh"

using namespace cuzk;

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
#ifndef CUZK_BLOCK
#define CUZK_BLOCK 128
#endif
#ifndef CUZK_MIN_BLOCKS
#define CUZK_MIN_BLOCKS 6   // 80 registers: measured best on B200 (profiles/r01_tuning_notes.md)
#endif
constexpr int kBlock = CUZK_BLOCK;

// generate_round_constants : poseidon.cpp:33-44, evaluated with the reference's own multiply/add
__global__ void gen_round_constants_kernel(uint4 *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kRounds * 3) return;
  u32 base[8], mix[8], off[8], r[8], r2[8];
  set_small(base, (u32)(i + 1));
  set_small(mix, 0x89ABCDEFu);
  mix[1] = 0x01234567u;  // 0x123456789ABCDEF
  u64 o = (u64)i * 0x987654321ULL;
  set_small(off, (u32)o);
  off[1] = (u32)(o >> 32);
  fr_mul(r, base, mix);
  fr_add_general(r2, r, off);
  store_fr(out + 2 * i, r2);
}

template <int OP>
__global__ void __launch_bounds__(kBlock) fr_batch_kernel(const uint4 *__restrict__ a, const uint4 *__restrict__ b,
                                                           uint4 *__restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  fr_table_init();
  if (i >= n) return;
  if (OP == CUZK_FR_SUB) {
    const u64 *pa = reinterpret_cast<const u64 *>(a + 2 * i);
    const u64 *pb = reinterpret_cast<const u64 *>(b + 2 * i);
    u64 x[4] = {pa[0], pa[1], pa[2], pa[3]}, y[4] = {pb[0], pb[1], pb[2], pb[3]}, r[4];
    fr_sub_ref(r, x, y);
    u64 *po = reinterpret_cast<u64 *>(out + 2 * i);
    po[0] = r[0]; po[1] = r[1]; po[2] = r[2]; po[3] = r[3];
    return;
  }
  u32 x[8], y[8], r[8];
  load_fr(x, a + 2 * i);
  if (OP == CUZK_FR_ADD) {
    load_fr(y, b + 2 * i);
    fr_add_general(r, x, y);
  } else if (OP == CUZK_FR_MUL) {
    load_fr(y, b + 2 * i);
    fr_mul(r, x, y);
  } else if (OP == CUZK_FR_SQR) {
    fr_sqr(r, x);
  } else {
    fr_pow5(r, x);
  }
  store_fr(out + 2 * i, r);
}

// batch_hash_single: state [1, in, 0]
__global__ void __launch_bounds__(kBlock, CUZK_MIN_BLOCKS) hash_single_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  fr_table_init();
  if (i >= n) return;
  u32 r[8];
  sponge_n(r, 1u, 1, [&](u32(&x)[8], int) { load_fr(x, in + 2 * i); });
  store_fr(out + 2 * i, r);
}

// batch_hash_pairs: state [2, l, r]  -- the headline kernel
__global__ void __launch_bounds__(kBlock, CUZK_MIN_BLOCKS) hash_pairs_kernel(const uint4 *__restrict__ l, const uint4 *__restrict__ r,
                                                             uint4 *__restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  fr_table_init();
  if (i >= n) return;
  u32 h[8];
  sponge_n(h, 2u, 2, [&](u32(&x)[8], int j) { load_fr(x, (j == 0 ? l : r) + 2 * i); });
  store_fr(out + 2 * i, h);
}

// batch_permutation: in-place, caller-supplied (possibly non-canonical) states
__global__ void __launch_bounds__(kBlock, CUZK_MIN_BLOCKS) permutation_kernel(uint4 *states, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  fr_table_init();
  if (i >= n) return;
  u32 s0[8], s1[8], s2[8], unc = 0;
  load_fr_plain(s0, states + 6 * i);
  load_fr_plain(s1, states + 6 * i + 2);
  load_fr_plain(s2, states + 6 * i + 4);
  permute_t<false, false>(s0, s1, s2, unc);
  if (unc != 0) {   // undecided comparison on the fast path: evaluate again exactly from the untouched input
    atomicAdd(&g_exact_fallbacks, 1ull);
    u32 st[24];
    const u32 *src = reinterpret_cast<const u32 *>(states + 6 * i);
#pragma unroll
    for (int w = 0; w < 24; ++w) st[w] = src[w];
    permute_exact(st, 0);
#pragma unroll
    for (int w = 0; w < 8; ++w) { s0[w] = st[w]; s1[w] = st[8 + w]; s2[w] = st[16 + w]; }
  }
  store_fr(states + 6 * i, s0);
  store_fr(states + 6 * i + 2, s1);
  store_fr(states + 6 * i + 4, s2);
}

// test hook: one MDS layer on canonical states (mode 0 = production fast path with fallback, 1 = exact path only)
__global__ void __launch_bounds__(kBlock) debug_mds_kernel(uint4 *states, size_t n, int mode) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  fr_table_init();
  if (i >= n) return;
  u32 s0[8], s1[8], s2[8];
  load_fr_plain(s0, states + 6 * i);
  load_fr_plain(s1, states + 6 * i + 2);
  load_fr_plain(s2, states + 6 * i + 4);
  if (mode == 0) mds(s0, s1, s2);
  else mds_exact(s0, s1, s2);
  store_fr(states + 6 * i, s0);
  store_fr(states + 6 * i + 2, s1);
  store_fr(states + 6 * i + 4, s2);
}

// test hook: the FAST-PATH field operations on their own, with the "undecided comparison" flag they raise.
// op 0 = reduce (any 256-bit a), 1 = multiply, 2 = square, 3 = power5.  Soundness property checked by the tests:
// flags[i] == 0  =>  out[i] equals the reference operation bit for bit.
__global__ void __launch_bounds__(kBlock) debug_fast_ops_kernel(int op, const uint4 *__restrict__ a, const uint4 *__restrict__ b,
                                                                 uint4 *__restrict__ out, u32 *__restrict__ flags, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  fr_table_init();
  if (i >= n) return;
  u32 x[8], y[8], r[8], unc = 0;
  load_fr(x, a + 2 * i);
  if (op == 0) {
#pragma unroll
    for (int w = 0; w < 8; ++w) r[w] = x[w];
    fr_reduce_fast(r, unc);
  } else if (op == 1) {
    load_fr(y, b + 2 * i);
    fr_mul_t<false>(r, x, y, unc);
  } else if (op == 2) {
    fr_sqr_t<false>(r, x, unc);
  } else {
    fr_pow5_t<false>(r, x, unc);
  }
  store_fr(out + 2 * i, r);
  flags[i] = unc;
}

// generic sponge: out[i] = sponge(in[i*width ..], ds)
__global__ void __launch_bounds__(kBlock, CUZK_MIN_BLOCKS) sponge_kernel(const uint4 *__restrict__ in, int width, u32 ds_lo, u32 ds_hi,
                                                         uint4 *__restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  fr_table_init();
  if (i >= n) return;
  const uint4 *base = in + 2 * i * (size_t)width;
  u32 r[8];
  sponge_n(r, ds_lo, ds_hi, width, [&](u32(&x)[8], int j) { load_fr(x, base + 2 * j); });
  store_fr(out + 2 * i, r);
}

// padding chain for one arity: pad[0] = hash_multiple(arity zeros), pad[l+1] = hash_multiple(arity x pad[l])
// computes levels [start, end); level start-1 must already be in pad[] when start > 0
__global__ void padding_chain_kernel(uint4 *pad, int arity, int start, int end) {
  fr_table_init();
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  u32 cur[8];
  if (start > 0) load_fr_plain(cur, pad + 2 * (start - 1));
  else set_small(cur, 0);
  for (int l = start; l < end; ++l) {
    u32 outv[8];
    const u32(&c)[8] = cur;
    sponge_n(outv, 3u, arity, [&](u32(&x)[8], int) {
#pragma unroll
      for (int w = 0; w < 8; ++w) x[w] = c[w];
    });
    store_fr(pad + 2 * l, outv);
#pragma unroll
    for (int w = 0; w < 8; ++w) cur[w] = outv[w];
  }
}

// level 0: copy the n leaves and append padding E_0 up to `padded`.  Forest form: tree t reads leaves + t * n and writes
// out + t * out_stride (elements); a single tree is ntrees = 1.
__global__ void merkle_pad_leaves_kernel(const uint4 *__restrict__ leaves, size_t n, size_t padded,
                                         const uint4 *__restrict__ pad, uint4 *__restrict__ out, size_t ntrees, size_t out_stride) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= padded * ntrees) return;
  const size_t tree = t / padded, i = t - tree * padded;
  const uint4 *src = (i < n) ? (leaves + 2 * (tree * n + i)) : pad;
  uint4 *dst = out + 2 * (tree * out_stride + i);
  dst[0] = src[0];
  dst[1] = src[1];
}

// one level: out[i] = hash_multiple(in[i*arity .. i*arity+arity-1]).  Only the first `in_real` inputs exist in
// memory; children beyond them are the padding constant of the input level (pad_in), and output nodes with no real
// child are the padding constant of the output level (pad_out) -- never hashed.
// build_level_kernel : merkle_tree_cuda.cu:45-64 / build_tree_bottom_up : merkle_tree.cpp:66-97
__global__ void __launch_bounds__(kBlock, CUZK_MIN_BLOCKS) merkle_level_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out,
                                                               size_t in_real, size_t out_count, int arity,
                                                               const uint4 *__restrict__ pad_in, const uint4 *__restrict__ pad_out,
                                                               size_t ntrees, size_t tree_stride) {
  // forest form: `ntrees` trees of identical shape, tree t at in/out + t * tree_stride elements; thread = (tree, node)
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  fr_table_init();
  if (t >= out_count * ntrees) return;
  const size_t tree = t / out_count, i = t - tree * out_count;
  in += 2 * tree * tree_stride;
  out += 2 * tree * tree_stride;
  const size_t first = i * (size_t)arity;
  if (first >= in_real) {
    out[2 * i] = pad_out[0];
    out[2 * i + 1] = pad_out[1];
    return;
  }
  u32 r[8];
  sponge_n(r, 3u, arity, [&](u32(&x)[8], int j) {
    const uint4 *src = (first + j < in_real) ? in + 2 * (first + j) : pad_in;
    load_fr_plain(x, src);
  });
  store_fr(out + 2 * i, r);
}

// incremental update, step 0: write the new leaf values (level 0).  Entry s of the batch, sorted by leaf index (stable, so
// equal indices keep their batch order): only the LAST entry of a run of equal indices writes -- the value a serial loop of
// update_leaf calls would leave behind (merkle_tree.cpp:294-301).  Indices >= n are skipped and counted in *oob (the
// reference throws std::out_of_range there, :296-298).  pos == nullptr: the batch is unsorted and holds a single entry.
__global__ void merkle_write_leaves_kernel(uint4 *__restrict__ level0, const u64 *__restrict__ sorted_idx, const u32 *__restrict__ pos,
                                           const uint4 *__restrict__ values, size_t count, size_t n, unsigned long long *oob) {
  size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= count) return;
  const u64 idx = sorted_idx[s];
  if (s + 1 < count && sorted_idx[s + 1] == idx) return;
  if (idx >= n) {
    atomicAdd(oob, 1ull);
    return;
  }
  const size_t q = pos ? pos[s] : s;
  level0[2 * idx] = values[2 * q];
  level0[2 * idx + 1] = values[2 * q + 1];
}
__global__ void iota_u32_kernel(u32 *out, size_t count) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = (u32)i;
}
// incremental update, one level: thread q re-hashes the level-l ancestor of leaf indices[q] from its children.
// Updates that share an ancestor compute the same value from the same children and store it twice (benign); indices >= n
// are skipped.  in = level l-1, out = level l.
__global__ void __launch_bounds__(kBlock, CUZK_MIN_BLOCKS) merkle_update_level_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out,
                                                                      const u64 *__restrict__ indices, size_t count, size_t n, u64 divisor,
                                                                      int arity) {
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  fr_table_init();
  if (q >= count) return;
  if (indices[q] >= n) return;
  const size_t node = indices[q] / divisor;          // ancestor index at the output level (divisor = arity^l)
  const uint4 *kids = in + 2 * node * (size_t)arity;
  u32 r[8];
  sponge_n(r, 3u, arity, [&](u32(&x)[8], int j) { load_fr_plain(x, kids + 2 * j); });
  store_fr(out + 2 * node, r);
}

// proofs from level arrays: one thread per (proof, level)
__global__ void merkle_prove_kernel(const uint4 *__restrict__ levels, size_t n, size_t padded, int arity, int nlv,
                                    const u64 *__restrict__ indices, size_t num_proofs, uint4 *__restrict__ sib,
                                    u32 *__restrict__ pos) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_proofs * (size_t)nlv) return;
  size_t q = t / nlv;
  int l = (int)(t % nlv);
  u64 idx = indices[q];
  if (idx >= n) {
    pos[t] = 0xFFFFFFFFu;
    return;
  }
  size_t off = 0, p = padded;
  for (int k = 0; k < l; ++k) {
    off += p;
    p /= arity;
    idx /= arity;
  }
  u32 my = (u32)(idx % arity);
  size_t base = off + (idx - my);
  pos[t] = my;
  uint4 *dst = sib + 2 * t * (size_t)(arity - 1);
  int w = 0;
  for (int c = 0; c < arity; ++c) {
    if (c == (int)my) continue;
    dst[2 * w] = levels[2 * (base + c)];
    dst[2 * w + 1] = levels[2 * (base + c) + 1];
    ++w;
  }
}

// verify: one thread per proof.  batch_verify_proofs_kernel : merkle_tree_cuda.cu:67-118 / verify_proof : merkle_tree.cpp:214-254
__global__ void __launch_bounds__(kBlock, CUZK_MIN_BLOCKS) merkle_verify_kernel(const uint4 *__restrict__ leaves, const uint4 *__restrict__ sib,
                                                                const u32 *__restrict__ pos, int nlv, int arity,
                                                                const uint4 *__restrict__ root, uint4 root_lo, uint4 root_hi,
                                                                uint8_t *__restrict__ results, size_t num_proofs) {
  // the expected root comes from device memory (`root`) or, for host-buffer calls, by value (root == nullptr)
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  fr_table_init();
  if (q >= num_proofs) return;
  u32 cur[8];
  load_fr(cur, leaves + 2 * q);
  bool ok = true;
#pragma unroll 1
  for (int l = 0; l < nlv; ++l) {
    const u32 my = pos[q * (size_t)nlv + l];
    if (my >= (u32)arity) { ok = false; break; }
    const uint4 *sb = sib + 2 * (q * (size_t)nlv + l) * (size_t)(arity - 1);
    u32 r[8];
    const u32(&c)[8] = cur;
    sponge_n(r, 3u, arity, [&](u32(&x)[8], int j) {
      if (j == (int)my) {
#pragma unroll
        for (int w = 0; w < 8; ++w) x[w] = c[w];
      } else {
        load_fr(x, sb + 2 * (j < (int)my ? j : j - 1));
      }
    });
#pragma unroll
    for (int w = 0; w < 8; ++w) cur[w] = r[w];
  }
  if (ok) {
    u32 rt[8];
    if (root) {
      load_fr(rt, root);
    } else {
      rt[0] = root_lo.x; rt[1] = root_lo.y; rt[2] = root_lo.z; rt[3] = root_lo.w;
      rt[4] = root_hi.x; rt[5] = root_hi.y; rt[6] = root_hi.z; rt[7] = root_hi.w;
    }
    u32 diff = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) diff |= rt[w] ^ cur[w];
    ok = diff == 0;
  }
  results[q] = ok ? 1 : 0;
}

// ---- synthetic inputs ----
__device__ __forceinline__ u64 splitmix64_dev(u64 seed, u64 idx) {
  u64 z = seed * 0xD1342543DE82EF95ULL + (idx + 1) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__global__ void synth_elements_kernel(u64 *out, size_t n, u64 seed, u64 start, int canonical) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  fr_table_init();
  if (i >= n) return;
  u64 v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = splitmix64_dev(seed, 4 * (start + i) + j);
  if (canonical) v[3] &= 0x0FFFFFFFFFFFFFFFULL;
  reinterpret_cast<ulonglong4 *>(out)[i] = make_ulonglong4(v[0], v[1], v[2], v[3]);
}
__global__ void synth_u64_leaves_kernel(u64 *out, size_t n, u64 seed, u64 start) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  fr_table_init();
  if (i >= n) return;
  reinterpret_cast<ulonglong4 *>(out)[i] = make_ulonglong4(splitmix64_dev(seed, start + i), 0, 0, 0);
}

#include "coop_kernels.cuh"
