// cuzk_kernels.cu -- sm_100a kernels and the extern "C" layer of libcuzk_b200.so (include/cuzk_b200.h).
//
// One thread evaluates one unit (field op, permutation chain, Merkle node or proof): the work is
// ~130 k integer instructions per permutation against <= 256 bytes of traffic, so the kernels are bound
// by the integer pipes (IMAD.WIDE + carry-chain IADD3), not by HBM; see DESIGN.md for the roofline.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cuzk_b200.h"
#include "poseidon.cuh"

using namespace cuzk;

// ------------------------------------------------------------------------------------------------
// host-side state
// ------------------------------------------------------------------------------------------------
namespace {

thread_local std::string g_err;
std::mutex g_mu;
int g_refcount = 0;
int g_device = -1;
int g_sm_count = 148;
std::atomic<uint64_t> g_launches{0};
uint64_t g_host_rc[kRounds * 3 * 4];
uint64_t g_host_mds[9 * 4];

// padding constants E_l per arity: E_0 = empty_hash(arity), E_{l+1} = hash_multiple(arity x E_l)
constexpr int kMaxPadLevels = 41;
uint64_t *g_d_pad[9] = {nullptr};           // device, kMaxPadLevels x 4 u64 per arity (per cuzk_init .. cuzk_shutdown)
uint64_t g_h_pad[9][kMaxPadLevels][4];      // host copy; constants of the hash function, kept for the life of the process
int g_h_pad_levels[9] = {0};                // how many levels the host copy holds
int g_pad_levels[9] = {0};                  // how many levels the device copy holds
std::mutex g_pad_mu;

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char *what) {
  return fail(CUZK_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                          \
  do {                                                    \
    cudaError_t e__ = (call);                             \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

inline cudaStream_t S(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// grid for n one-thread units at `block` threads: whole waves of SM-count multiples
inline unsigned grid_for(size_t n, unsigned block) {
  size_t g = (n + block - 1) / block;
  if (g == 0) g = 1;
  return (unsigned)g;
}

}  // namespace

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
  if (i >= n) return;
  u32 r[8];
  sponge_n(r, 1u, 1, [&](u32(&x)[8], int) { load_fr(x, in + 2 * i); });
  store_fr(out + 2 * i, r);
}

// batch_hash_pairs: state [2, l, r]  -- the headline kernel
__global__ void __launch_bounds__(kBlock, CUZK_MIN_BLOCKS) hash_pairs_kernel(const uint4 *__restrict__ l, const uint4 *__restrict__ r,
                                                             uint4 *__restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 h[8];
  sponge_n(h, 2u, 2, [&](u32(&x)[8], int j) { load_fr(x, (j == 0 ? l : r) + 2 * i); });
  store_fr(out + 2 * i, h);
}

// batch_permutation: in-place, caller-supplied (possibly non-canonical) states
__global__ void __launch_bounds__(kBlock, CUZK_MIN_BLOCKS) permutation_kernel(uint4 *states, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
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
  if (i >= n) return;
  const uint4 *base = in + 2 * i * (size_t)width;
  u32 r[8];
  sponge_n(r, ds_lo, ds_hi, width, [&](u32(&x)[8], int j) { load_fr(x, base + 2 * j); });
  store_fr(out + 2 * i, r);
}

// padding chain for one arity: pad[0] = hash_multiple(arity zeros), pad[l+1] = hash_multiple(arity x pad[l])
// computes levels [start, end); level start-1 must already be in pad[] when start > 0
__global__ void padding_chain_kernel(uint4 *pad, int arity, int start, int end) {
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

// two fused levels: thread i hashes `arity` groups of `arity` inputs into its own shared-memory slots and then hashes
// those into out[i]; the middle level never reaches HBM unless `mid_out` is given (full-tree builds keep every level).
// Same padding rules as merkle_level_kernel (pad_in / pad_mid / pad_out are consecutive padding constants).
__global__ void __launch_bounds__(kBlock, CUZK_MIN_BLOCKS) merkle_fused2_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ mid_out,
                                                                uint4 *__restrict__ out, size_t in_real, size_t out_count,
                                                                int arity, const uint4 *__restrict__ pad, size_t ntrees,
                                                                size_t tree_stride) {
  extern __shared__ uint4 smem[];                    // [2 * arity][kBlock] uint4: slot-major, so a warp's accesses never conflict
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= out_count * ntrees) return;
  const size_t tree = t / out_count, i = t - tree * out_count;   // forest form, see merkle_level_kernel
  in += 2 * tree * tree_stride;
  out += 2 * tree * tree_stride;
  if (mid_out) mid_out += 2 * tree * tree_stride;
  const uint4 *pad_in = pad, *pad_mid = pad + 2, *pad_out = pad + 4;
  const size_t span = (size_t)arity * arity;
  if (i * span >= in_real) {
    out[2 * i] = pad_out[0];
    out[2 * i + 1] = pad_out[1];
    if (mid_out) {
      for (int g = 0; g < arity; ++g) {
        mid_out[2 * (i * arity + g)] = pad_mid[0];
        mid_out[2 * (i * arity + g) + 1] = pad_mid[1];
      }
    }
    return;
  }
  uint4 *mine = smem + threadIdx.x;                  // slot s of this thread lives at mine[s * kBlock]
#pragma unroll 1
  for (int g = 0; g < arity; ++g) {
    const size_t first = i * span + (size_t)g * arity;
    uint4 lo, hi;
    if (first >= in_real) {
      lo = pad_mid[0];
      hi = pad_mid[1];
    } else {
      u32 r[8];
      sponge_n(r, 3u, arity, [&](u32(&x)[8], int j) {
        const uint4 *src = (first + j < in_real) ? in + 2 * (first + j) : pad_in;
        load_fr_plain(x, src);
      });
      lo = make_uint4(r[0], r[1], r[2], r[3]);
      hi = make_uint4(r[4], r[5], r[6], r[7]);
    }
    mine[(2 * g) * kBlock] = lo;
    mine[(2 * g + 1) * kBlock] = hi;
    if (mid_out) {
      mid_out[2 * (i * arity + g)] = lo;
      mid_out[2 * (i * arity + g) + 1] = hi;
    }
  }
  u32 r[8];
  sponge_n(r, 3u, arity, [&](u32(&x)[8], int j) {
    const uint4 a = mine[(2 * j) * kBlock], b = mine[(2 * j + 1) * kBlock];
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
    x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  });
  store_fr(out + 2 * i, r);
}

// incremental update, step 0: write the new leaf values (level 0)
__global__ void merkle_write_leaves_kernel(uint4 *__restrict__ level0, const u64 *__restrict__ indices, const uint4 *__restrict__ values,
                                           size_t count) {
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= count) return;
  const u64 idx = indices[q];
  level0[2 * idx] = values[2 * q];
  level0[2 * idx + 1] = values[2 * q + 1];
}
// incremental update, one level: thread q re-hashes the level-`shift_level` ancestor of leaf indices[q] from its children.
// Updates that share an ancestor compute the same value and store it twice (benign).  in = level l-1, out = level l.
__global__ void __launch_bounds__(kBlock, CUZK_MIN_BLOCKS) merkle_update_level_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out,
                                                                      const u64 *__restrict__ indices, size_t count, u64 divisor,
                                                                      int arity) {
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= count) return;
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
  if (i >= n) return;
  u64 v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = splitmix64_dev(seed, 4 * (start + i) + j);
  if (canonical) v[3] &= 0x0FFFFFFFFFFFFFFFULL;
  reinterpret_cast<ulonglong4 *>(out)[i] = make_ulonglong4(v[0], v[1], v[2], v[3]);
}
__global__ void synth_u64_leaves_kernel(u64 *out, size_t n, u64 seed, u64 start) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  reinterpret_cast<ulonglong4 *>(out)[i] = make_ulonglong4(splitmix64_dev(seed, start + i), 0, 0, 0);
}

// ------------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------------
namespace {

// Host-buffer calls (mem == CUZK_MEM_HOST) stage through library-owned device buffers that are kept between calls
// (the reference mallocs and frees on every call, poseidon_cuda.cu:374-408) and are cut into chunks that alternate
// between two internal streams, so the H2D copy of chunk c+1 and the D2H copy of chunk c-1 overlap the kernel of chunk c.
constexpr int kPipeStreams = 2;
constexpr int kPipeSlots = 4;                       // up to 3 inputs + 1 output per stream
constexpr size_t kHashChunk = 148 * CUZK_MIN_BLOCKS * CUZK_BLOCK;   // one resident wave of one-thread-per-hash CTAs
constexpr size_t kCheapChunk = 1 << 20;             // element-wise field ops
constexpr int kWsSlots = 6;

struct HostPath {
  cudaStream_t stream[kPipeStreams] = {nullptr, nullptr};
  void *buf[kPipeStreams][kPipeSlots] = {};
  size_t cap[kPipeStreams][kPipeSlots] = {};
  void *ws[kWsSlots] = {};
  size_t ws_cap[kWsSlots] = {};
  bool ready = false;
} g_hp;
std::mutex g_hp_mu;   // host-buffer calls serialise on the staging buffers

int hp_reserve(void *&p, size_t &cap, size_t bytes) {
  if (bytes <= cap) return CUZK_OK;
  if (p) {
    CK(cudaDeviceSynchronize());
    CK(cudaFree(p));
    p = nullptr;
    cap = 0;
  }
  size_t want = bytes + bytes / 4;
  CK(cudaMalloc(&p, want));
  cap = want;
  return CUZK_OK;
}
int ws_get(int slot, size_t bytes, void **out) {
  int rc = hp_reserve(g_hp.ws[slot], g_hp.ws_cap[slot], bytes ? bytes : 1);
  *out = g_hp.ws[slot];
  return rc;
}
int hp_start() {
  if (g_hp.ready) return CUZK_OK;
  for (int i = 0; i < kPipeStreams; ++i) CK(cudaStreamCreateWithFlags(&g_hp.stream[i], cudaStreamNonBlocking));
  g_hp.ready = true;
  return CUZK_OK;
}
void hp_stop() {
  for (int i = 0; i < kPipeStreams; ++i) {
    for (int j = 0; j < kPipeSlots; ++j) {
      if (g_hp.buf[i][j]) cudaFree(g_hp.buf[i][j]);
      g_hp.buf[i][j] = nullptr;
      g_hp.cap[i][j] = 0;
    }
    if (g_hp.stream[i]) cudaStreamDestroy(g_hp.stream[i]);
    g_hp.stream[i] = nullptr;
  }
  for (int j = 0; j < kWsSlots; ++j) {
    if (g_hp.ws[j]) cudaFree(g_hp.ws[j]);
    g_hp.ws[j] = nullptr;
    g_hp.ws_cap[j] = 0;
  }
  g_hp.ready = false;
}
void pin_stop();

int require_init() {
  if (g_refcount <= 0) return fail(CUZK_ERR_CUDA, "cuzk_b200: library not initialised (call cuzk_init)");
  return CUZK_OK;
}

int check_launch(const char *what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, what);
  return CUZK_OK;
}

// ---- parallel host copies for pageable caller memory ----------------------------------------------------------------
// cudaMemcpyAsync on pageable memory (a std::vector, which is what the reference's API hands us) is staged by the driver
// through one thread at well under PCIe speed.  For such buffers the pipeline below stages through its own pinned bounce
// buffers and fills / drains them with a few worker threads, so the DMA runs at pinned-memory speed while the copy of the
// next chunk overlaps the kernel of the current one.
class CopyPool {
 public:
  void copy(void *dst, const void *src, size_t bytes) {
    if (bytes < (1u << 20) || !start()) {
      memcpy(dst, src, bytes);
      return;
    }
    const int parts = (int)workers_.size() + 1;
    const size_t slice = ((bytes / parts) + 4095) & ~(size_t)4095;
    {
      std::lock_guard<std::mutex> lk(mu_);
      dst_ = static_cast<char *>(dst);
      src_ = static_cast<const char *>(src);
      bytes_ = bytes;
      slice_ = slice;
      pending_ = (int)workers_.size();
      ++generation_;
    }
    cv_.notify_all();
    run_slice(parts - 1);   // the caller takes the last slice
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto &t : workers_) t.join();
  }

 private:
  bool start() {
    if (started_) return !workers_.empty();
    started_ = true;
    unsigned hw = std::thread::hardware_concurrency();
    int n = (int)std::min<unsigned>(3, hw > 2 ? hw / 2 - 1 : 0);   // three helpers + the caller saturate one socket's copy rate
    for (int i = 0; i < n; ++i) workers_.emplace_back([this, i] { loop(i); });
    return !workers_.empty();
  }
  void run_slice(int part) {
    const size_t off = slice_ * (size_t)part;
    if (off < bytes_) memcpy(dst_ + off, src_ + off, std::min(slice_, bytes_ - off));
  }
  void loop(int idx) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
        if (stop_) return;
        seen = generation_;
      }
      run_slice(idx);
      std::lock_guard<std::mutex> lk(mu_);
      if (--pending_ == 0) done_cv_.notify_one();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  char *dst_ = nullptr;
  const char *src_ = nullptr;
  size_t bytes_ = 0, slice_ = 0;
  int pending_ = 0;
  uint64_t generation_ = 0;
  bool stop_ = false, started_ = false;
};
CopyPool g_copy_pool;

struct PinnedStage {   // per stream: pinned twins of the device staging buffers + "results are in the bounce buffer" event
  void *buf[kPipeSlots] = {};
  size_t cap[kPipeSlots] = {};
  cudaEvent_t done = nullptr;
} g_pin[kPipeStreams];

int pin_reserve(void *&p, size_t &cap, size_t bytes) {
  if (bytes <= cap) return CUZK_OK;
  if (p) CK(cudaFreeHost(p));
  p = nullptr;
  cap = 0;
  CK(cudaHostAlloc(&p, bytes, cudaHostAllocDefault));
  cap = bytes;
  return CUZK_OK;
}
void pin_stop() {
  for (auto &st : g_pin) {
    for (int j = 0; j < kPipeSlots; ++j) {
      if (st.buf[j]) cudaFreeHost(st.buf[j]);
      st.buf[j] = nullptr;
      st.cap[j] = 0;
    }
    if (st.done) cudaEventDestroy(st.done);
    st.done = nullptr;
  }
}
bool is_pageable(const void *p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return attr.type == cudaMemoryTypeUnregistered;
}

// bulk copies between caller host memory and device memory on stream `st`; pageable memory goes through the pinned
// bounce buffers in 8 MiB pieces filled / drained by the copy pool.  Both return with the copy complete or enqueued such
// that `host` may be reused (upload) / read (download) by the caller.  Call with g_hp_mu held.
constexpr size_t kBulkPiece = (size_t)8 << 20;
int bulk_upload(void *dev, const void *host, size_t bytes, cudaStream_t st) {
  if (bytes < ((size_t)4 << 20) || !is_pageable(host)) {
    CK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, st));
    return CUZK_OK;
  }
  int rc;
  for (int s = 0; s < kPipeStreams; ++s) {
    if ((rc = pin_reserve(g_pin[s].buf[0], g_pin[s].cap[0], kBulkPiece))) return rc;
    if (!g_pin[s].done) CK(cudaEventCreateWithFlags(&g_pin[s].done, cudaEventDisableTiming));
  }
  bool used[kPipeStreams] = {};
  size_t at = 0;
  for (int c = 0; at < bytes; ++c) {
    const int s = c % kPipeStreams;
    const size_t m = std::min(kBulkPiece, bytes - at);
    if (used[s]) CK(cudaEventSynchronize(g_pin[s].done));   // the piece that used this bounce buffer has left it
    g_copy_pool.copy(g_pin[s].buf[0], static_cast<const char *>(host) + at, m);
    CK(cudaMemcpyAsync(static_cast<char *>(dev) + at, g_pin[s].buf[0], m, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(g_pin[s].done, st));
    used[s] = true;
    at += m;
  }
  for (int s = 0; s < kPipeStreams; ++s)
    if (used[s]) CK(cudaEventSynchronize(g_pin[s].done));
  return CUZK_OK;
}
int bulk_download(void *host, const void *dev, size_t bytes, cudaStream_t st) {
  if (bytes < ((size_t)4 << 20) || !is_pageable(host)) {
    CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return CUZK_OK;
  }
  int rc;
  for (int s = 0; s < kPipeStreams; ++s) {
    if ((rc = pin_reserve(g_pin[s].buf[0], g_pin[s].cap[0], kBulkPiece))) return rc;
    if (!g_pin[s].done) CK(cudaEventCreateWithFlags(&g_pin[s].done, cudaEventDisableTiming));
  }
  size_t pending_at[kPipeStreams] = {}, pending_m[kPipeStreams] = {};
  auto drain = [&](int s) -> int {
    if (!pending_m[s]) return CUZK_OK;
    CK(cudaEventSynchronize(g_pin[s].done));
    g_copy_pool.copy(static_cast<char *>(host) + pending_at[s], g_pin[s].buf[0], pending_m[s]);
    pending_m[s] = 0;
    return CUZK_OK;
  };
  size_t at = 0;
  for (int c = 0; at < bytes; ++c) {
    const int s = c % kPipeStreams;
    const size_t m = std::min(kBulkPiece, bytes - at);
    if ((rc = drain(s))) return rc;
    CK(cudaMemcpyAsync(g_pin[s].buf[0], static_cast<const char *>(dev) + at, m, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(g_pin[s].done, st));
    pending_at[s] = at;
    pending_m[s] = m;
    at += m;
  }
  for (int s = 0; s < kPipeStreams; ++s)
    if ((rc = drain(s))) return rc;
  return CUZK_OK;
}

// Chunked, double-buffered host->device->host pass.  `nin` input arrays of `in_bytes[k]` bytes per unit, one output
// array of `out_bytes` per unit (out may alias in[0] for in-place ops).  launch(stream, d_in[], d_out, m) enqueues the
// kernel(s) for m units.  Returns after every result byte is in `out`.
template <class Launch>
int host_pipeline(size_t n, size_t chunk, int nin, const void *const *in, const size_t *in_bytes, void *out, size_t out_bytes,
                  bool out_aliases_in0, Launch launch) {
  std::lock_guard<std::mutex> lk(g_hp_mu);
  int rc = hp_start();
  if (rc) return rc;
  if (chunk > n) chunk = n;
  size_t unit_bytes = out_bytes;
  for (int k = 0; k < nin; ++k) unit_bytes += in_bytes[k];
  // pageable caller memory and enough of it: go through the pinned bounce buffers (bounded to 16 MiB per array and slot)
  bool staged = n * unit_bytes >= ((size_t)4 << 20) && (is_pageable(out) || is_pageable(in[0]));
  if (staged) {
    size_t widest = out_bytes;
    for (int k = 0; k < nin; ++k) widest = std::max(widest, in_bytes[k]);
    chunk = std::max<size_t>(1, std::min(chunk, ((size_t)16 << 20) / widest));
  }
  for (int s = 0; s < kPipeStreams; ++s) {
    for (int k = 0; k < nin; ++k)
      if ((rc = hp_reserve(g_hp.buf[s][k], g_hp.cap[s][k], chunk * in_bytes[k]))) return rc;
    if (!out_aliases_in0 && (rc = hp_reserve(g_hp.buf[s][kPipeSlots - 1], g_hp.cap[s][kPipeSlots - 1], chunk * out_bytes))) return rc;
    if (staged) {
      for (int k = 0; k < nin; ++k)
        if ((rc = pin_reserve(g_pin[s].buf[k], g_pin[s].cap[k], chunk * in_bytes[k]))) return rc;
      if ((rc = pin_reserve(g_pin[s].buf[kPipeSlots - 1], g_pin[s].cap[kPipeSlots - 1], chunk * out_bytes))) return rc;
      if (!g_pin[s].done) CK(cudaEventCreateWithFlags(&g_pin[s].done, cudaEventDisableTiming));
    }
  }
  size_t done = 0;
  size_t drain_at[kPipeStreams] = {}, drain_m[kPipeStreams] = {};   // staged mode: the chunk whose results sit in each bounce buffer
  auto drain = [&](int s) -> int {
    if (drain_m[s] == 0) return CUZK_OK;
    CK(cudaEventSynchronize(g_pin[s].done));
    g_copy_pool.copy(static_cast<char *>(out) + drain_at[s] * out_bytes, g_pin[s].buf[kPipeSlots - 1], drain_m[s] * out_bytes);
    drain_m[s] = 0;
    return CUZK_OK;
  };
  for (int c = 0; done < n; ++c) {
    const int s = c % kPipeStreams;
    const size_t m = (n - done < chunk) ? n - done : chunk;
    cudaStream_t st = g_hp.stream[s];
    void *d_in[kPipeSlots] = {};
    if (staged && (rc = drain(s))) return rc;   // the slot's previous results leave before its buffers are reused
    for (int k = 0; k < nin; ++k) {
      d_in[k] = g_hp.buf[s][k];
      const char *src = static_cast<const char *>(in[k]) + done * in_bytes[k];
      if (staged) {
        g_copy_pool.copy(g_pin[s].buf[k], src, m * in_bytes[k]);
        src = static_cast<const char *>(g_pin[s].buf[k]);
      }
      CK(cudaMemcpyAsync(d_in[k], src, m * in_bytes[k], cudaMemcpyHostToDevice, st));
    }
    void *d_out = out_aliases_in0 ? d_in[0] : g_hp.buf[s][kPipeSlots - 1];
    if ((rc = launch(st, d_in, d_out, m))) return rc;
    if (staged) {
      CK(cudaMemcpyAsync(g_pin[s].buf[kPipeSlots - 1], d_out, m * out_bytes, cudaMemcpyDeviceToHost, st));
      CK(cudaEventRecord(g_pin[s].done, st));
      drain_at[s] = done;
      drain_m[s] = m;
    } else {
      CK(cudaMemcpyAsync(static_cast<char *>(out) + done * out_bytes, d_out, m * out_bytes, cudaMemcpyDeviceToHost, st));
    }
    done += m;
  }
  if (staged) {
    for (int s = 0; s < kPipeStreams; ++s)
      if ((rc = drain(s))) return rc;
  }
  for (int s = 0; s < kPipeStreams; ++s) CK(cudaStreamSynchronize(g_hp.stream[s]));
  return CUZK_OK;
}

// makes the padding constants E_0 .. E_{need-1} of `arity` available on the device.  The chain is sequential (one thread,
// ceil(arity/2) permutations per level), so it is computed only as far as trees need it, extended on demand, and the
// values -- constants of the hash function -- are cached on the host across cuzk_shutdown / cuzk_init cycles.
int ensure_padding(unsigned arity, int need = 2) {
  if (need > kMaxPadLevels) return fail(CUZK_ERR_INVALID, "tree too tall");
  std::lock_guard<std::mutex> lk(g_pad_mu);
  if (g_d_pad[arity] && g_pad_levels[arity] >= need) return CUZK_OK;
  if (!g_d_pad[arity]) {
    uint64_t *d = nullptr;
    CK(cudaMalloc(&d, (size_t)kMaxPadLevels * 32));
    g_d_pad[arity] = d;
    g_pad_levels[arity] = 0;
  }
  uint64_t *d = g_d_pad[arity];
  if (g_pad_levels[arity] < g_h_pad_levels[arity]) {   // bring the device copy up to what the host already knows
    CK(cudaMemcpy(d, g_h_pad[arity], (size_t)g_h_pad_levels[arity] * 32, cudaMemcpyHostToDevice));
    g_pad_levels[arity] = g_h_pad_levels[arity];
  }
  if (g_pad_levels[arity] >= need) return CUZK_OK;
  const int start = g_pad_levels[arity], end = std::min(kMaxPadLevels, std::max(need, start + 4));
  padding_chain_kernel<<<1, 32>>>(reinterpret_cast<uint4 *>(d), (int)arity, start, end);
  int rc = check_launch("padding_chain_kernel");
  if (rc) return rc;
  CK(cudaMemcpy(g_h_pad[arity][start], d + 4 * start, (size_t)(end - start) * 32, cudaMemcpyDeviceToHost));
  g_pad_levels[arity] = g_h_pad_levels[arity] = end;
  return CUZK_OK;
}

int check_arity(unsigned arity) {
  if (arity < 2 || arity > 8) return fail(CUZK_ERR_INVALID, "arity must be between 2 and 8");
  return CUZK_OK;
}

inline size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }

// device-pointer implementations ------------------------------------------------------------------
int fr_batch_dev(int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n, cudaStream_t st) {
  if (n == 0) return CUZK_OK;
  const uint4 *pa = reinterpret_cast<const uint4 *>(a), *pb = reinterpret_cast<const uint4 *>(b);
  uint4 *po = reinterpret_cast<uint4 *>(out);
  unsigned g = grid_for(n, kBlock);
  switch (op) {
    case CUZK_FR_ADD: fr_batch_kernel<CUZK_FR_ADD><<<g, kBlock, 0, st>>>(pa, pb, po, n); break;
    case CUZK_FR_SUB: fr_batch_kernel<CUZK_FR_SUB><<<g, kBlock, 0, st>>>(pa, pb, po, n); break;
    case CUZK_FR_MUL: fr_batch_kernel<CUZK_FR_MUL><<<g, kBlock, 0, st>>>(pa, pb, po, n); break;
    case CUZK_FR_SQR: fr_batch_kernel<CUZK_FR_SQR><<<g, kBlock, 0, st>>>(pa, pb, po, n); break;
    case CUZK_FR_POW5: fr_batch_kernel<CUZK_FR_POW5><<<g, kBlock, 0, st>>>(pa, pb, po, n); break;
    default: return fail(CUZK_ERR_INVALID, "unknown field op");
  }
  return check_launch("fr_batch_kernel");
}

// Fuse two levels into one launch (merkle_fused2_kernel), or run them as two launches?  Both do the same hashing and a
// node hash is ~190 k instructions against 288 bytes, so the saved middle-level traffic is worth nothing; what differs is
// the tail.  A fused thread hashes arity + 1 nodes back to back, so the last, partly filled wave of CTAs costs
// (arity + 1) node times, while per-level launches quantise in single node times.  Measured on B200 (tools/fuse_probe.py,
// profiles/r01_fuse_probe.json): per-level launches win at every shard size (2^23 leaves, arity 8: 31.8 ms against
// 45.2 ms fused; 2^26: 214 ms against 237 ms), so they are the default; the fused kernel stays selectable and
// parity-tested (cuzk_debug_set_fuse).
int g_fuse_mode = 0;   // 0: one launch per level, 1: fuse pairs of levels (cuzk_debug_set_fuse)
inline bool fuse_two_levels(size_t /*mid_nodes*/, size_t out_nodes, unsigned /*arity*/) { return g_fuse_mode != 0 && out_nodes != 0; }

int launch_level(const uint4 *in, uint4 *out, size_t in_real, size_t out_count, unsigned arity, const uint4 *pad_in, cudaStream_t st,
                 size_t ntrees = 1, size_t tree_stride = 0) {
  merkle_level_kernel<<<grid_for(out_count * ntrees, kBlock), kBlock, 0, st>>>(in, out, in_real, out_count, (int)arity, pad_in, pad_in + 2,
                                                                               ntrees, tree_stride);
  return check_launch("merkle_level_kernel");
}
int launch_fused2(const uint4 *in, uint4 *mid, uint4 *out, size_t in_real, size_t out_count, unsigned arity, const uint4 *pad_in,
                  cudaStream_t st, size_t ntrees = 1, size_t tree_stride = 0) {
  const size_t smem = (size_t)2 * arity * kBlock * sizeof(uint4);
  merkle_fused2_kernel<<<grid_for(out_count * ntrees, kBlock), kBlock, smem, st>>>(in, mid, out, in_real, out_count, (int)arity, pad_in,
                                                                                   ntrees, tree_stride);
  return check_launch("merkle_fused2_kernel");
}

// builds `ntrees` trees of n leaves each in one pass: one launch per level (or per two levels) for the whole forest.
// leaves: ntrees x n elements; levels_out: ntrees flat level-major trees of total_nodes(n) elements each.
int merkle_build_dev(const uint64_t *leaves, size_t n, unsigned arity, uint64_t *levels_out, cudaStream_t st, size_t ntrees = 1) {
  int rc = ensure_padding(arity, (int)cuzk_merkle_num_levels(n, arity) + 1);
  if (rc) return rc;
  const uint4 *pad = reinterpret_cast<const uint4 *>(g_d_pad[arity]);
  size_t padded = cuzk_merkle_padded_leaves(n, arity);
  const size_t stride = cuzk_merkle_total_nodes(n, arity);
  uint4 *cur = reinterpret_cast<uint4 *>(levels_out);
  merkle_pad_leaves_kernel<<<grid_for(padded * ntrees, 256), 256, 0, st>>>(reinterpret_cast<const uint4 *>(leaves), n, padded, pad, cur,
                                                                          ntrees, stride);
  if ((rc = check_launch("merkle_pad_leaves_kernel"))) return rc;
  size_t p = padded, real = n;
  int level = 0;
  while (p > 1) {
    const size_t q = p / arity;
    if (q > 1 && fuse_two_levels(q * ntrees, (q / arity) * ntrees, arity)) {
      const size_t q2 = q / arity;
      if ((rc = launch_fused2(cur, cur + 2 * p, cur + 2 * p + 2 * q, real, q2, arity, pad + 2 * level, st, ntrees, stride))) return rc;
      cur += 2 * p + 2 * q;
      real = ceil_div(ceil_div(real, arity), arity);
      p = q2;
      level += 2;
    } else {
      if ((rc = launch_level(cur, cur + 2 * p, real, q, arity, pad + 2 * level, st, ntrees, stride))) return rc;
      cur += 2 * p;
      real = ceil_div(real, arity);
      p = q;
      level += 1;
    }
  }
  return CUZK_OK;
}

// roots of `count` consecutive subtrees of arity^height (virtual) leaves whose first n leaves are in memory
int subtree_roots_one_stream(const uint64_t *leaves, size_t n, unsigned arity, unsigned height, size_t count, uint64_t *roots_out,
                             cudaStream_t st) {
  int rc = ensure_padding(arity, (int)height + 2);
  if (rc) return rc;
  const uint4 *pad = reinterpret_cast<const uint4 *>(g_d_pad[arity]);
  if (height == 0) {
    merkle_pad_leaves_kernel<<<grid_for(count, 256), 256, 0, st>>>(reinterpret_cast<const uint4 *>(leaves), n, count, pad,
                                                                  reinterpret_cast<uint4 *>(roots_out), 1, 0);
    return check_launch("merkle_pad_leaves_kernel");
  }
  // stream-ordered scratch for the levels between the leaves and the roots; only nodes with a real leaf below them are stored
  void *scratch[2] = {nullptr, nullptr};
  auto release = [&]() {
    for (void *s : scratch)
      if (s) cudaFreeAsync(s, st);
  };
  const uint4 *cur = reinterpret_cast<const uint4 *>(leaves);
  size_t real = n;
  unsigned level = 0;
  int flip = 0;
  while (level < height) {
    const unsigned step =
        (height - level >= 2 && fuse_two_levels(ceil_div(real, arity), ceil_div(real, (size_t)arity * arity), arity)) ? 2 : 1;
    const bool last = level + step == height;
    size_t out_real = ceil_div(real, arity);
    if (step == 2) out_real = ceil_div(out_real, arity);
    const size_t out_count = last ? count : out_real;
    uint4 *dst;
    if (last) {
      dst = reinterpret_cast<uint4 *>(roots_out);
    } else {
      if (scratch[flip]) { cudaFreeAsync(scratch[flip], st); scratch[flip] = nullptr; }
      cudaError_t e = cudaMallocAsync(&scratch[flip], (out_count ? out_count : 1) * 32, st);
      if (e != cudaSuccess) { release(); return cuda_fail(e, "cudaMallocAsync(scratch)"); }
      dst = reinterpret_cast<uint4 *>(scratch[flip]);
      flip ^= 1;
    }
    if (out_count) {
      rc = (step == 2) ? launch_fused2(cur, nullptr, dst, real, out_count, arity, pad + 2 * level, st)
                       : launch_level(cur, dst, real, out_count, arity, pad + 2 * level, st);
      if (rc) { release(); return rc; }
    }
    cur = dst;
    real = out_real;
    level += step;
  }
  release();
  return CUZK_OK;
}

// The upper levels of a subtree are narrow: a level with fewer nodes than the chip has thread slots takes one node-hash
// latency (about 0.18 ms per permutation) however few nodes it has.  With several subtrees per call, groups of subtrees run
// on separate internal streams, so the narrow levels of one group hide behind the wide levels of the next.
constexpr int kSubtreeStreams = 4;
cudaStream_t g_sub_stream[kSubtreeStreams] = {};
cudaEvent_t g_sub_fork = nullptr, g_sub_join[kSubtreeStreams] = {};
std::mutex g_sub_mu;   // the internal streams and events are shared by all callers

int subtree_streams_start() {
  if (g_sub_fork) return CUZK_OK;
  for (int i = 0; i < kSubtreeStreams; ++i) {
    CK(cudaStreamCreateWithFlags(&g_sub_stream[i], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&g_sub_join[i], cudaEventDisableTiming));
  }
  CK(cudaEventCreateWithFlags(&g_sub_fork, cudaEventDisableTiming));
  return CUZK_OK;
}
void subtree_streams_stop() {
  for (int i = 0; i < kSubtreeStreams; ++i) {
    if (g_sub_stream[i]) cudaStreamDestroy(g_sub_stream[i]);
    if (g_sub_join[i]) cudaEventDestroy(g_sub_join[i]);
    g_sub_stream[i] = nullptr;
    g_sub_join[i] = nullptr;
  }
  if (g_sub_fork) cudaEventDestroy(g_sub_fork);
  g_sub_fork = nullptr;
}

int subtree_roots_dev(const uint64_t *leaves, size_t n, unsigned arity, unsigned height, size_t count, uint64_t *roots_out,
                      cudaStream_t st) {
  size_t span = 1;
  for (unsigned i = 0; i < height; ++i) span *= arity;
  const size_t real_subtrees = ceil_div(n, span);   // subtrees with at least one real leaf; the rest are padding constants
  const size_t groups = std::min<size_t>(real_subtrees, kSubtreeStreams);
  if (height < 3 || groups < 2 || g_sub_fork == nullptr) return subtree_roots_one_stream(leaves, n, arity, height, count, roots_out, st);
  std::lock_guard<std::mutex> lk(g_sub_mu);
  CK(cudaEventRecord(g_sub_fork, st));
  int rc = CUZK_OK;
  for (size_t g = 0; g < groups; ++g) {
    const size_t lo = real_subtrees * g / groups;
    const size_t hi = (g + 1 == groups) ? count : real_subtrees * (g + 1) / groups;   // the last group also writes the padding roots
    const size_t first_leaf = lo * span;
    const size_t n_g = std::min(n - first_leaf, (hi - lo) * span);
    cudaStream_t sg = g_sub_stream[g];
    CK(cudaStreamWaitEvent(sg, g_sub_fork, 0));
    const int r = subtree_roots_one_stream(leaves + 4 * first_leaf, n_g, arity, height, hi - lo, roots_out + 4 * lo, sg);
    if (r && !rc) rc = r;
    CK(cudaEventRecord(g_sub_join[g], sg));
    CK(cudaStreamWaitEvent(st, g_sub_join[g], 0));
  }
  return rc;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// extern "C"
// ------------------------------------------------------------------------------------------------
extern "C" {

const char *cuzk_last_error(void) { return g_err.c_str(); }
const char *cuzk_version(void) { return "cuzk_b200 0.2 (sm_100a)"; }
uint64_t cuzk_launch_count(void) { return g_launches.load(); }
int cuzk_debug_set_fuse(int mode) {
  const int old = g_fuse_mode;
  g_fuse_mode = mode > 0 ? 1 : 0;
  return old;
}
uint64_t cuzk_debug_fallback_count(void) {
  unsigned long long v = 0;
  if (cudaMemcpyFromSymbol(&v, g_exact_fallbacks, sizeof v) != cudaSuccess) return ~0ull;
  return v;
}

int cuzk_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int cuzk_is_initialized(void) { return g_refcount > 0 ? 1 : 0; }

int cuzk_device_info(int device, cuzk_device_info_t *out) {
  if (!out) return fail(CUZK_ERR_INVALID, "null pointer");
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  memset(out, 0, sizeof *out);
  strncpy(out->name, prop.name, sizeof out->name - 1);
  out->cc_major = prop.major;
  out->cc_minor = prop.minor;
  out->sm_count = prop.multiProcessorCount;
  out->max_threads_per_block = prop.maxThreadsPerBlock;
  out->total_mem_bytes = prop.totalGlobalMem;
  return CUZK_OK;
}

int cuzk_init(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_refcount > 0) {
    if (device != g_device) return fail(CUZK_ERR_INVALID, "cuzk_init: already initialised on another device");
    ++g_refcount;
    return CUZK_OK;
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return fail(CUZK_ERR_CUDA, "cuzk_init: no CUDA device (this library has no CPU fallback)");
  if (device < 0 || device >= count) return fail(CUZK_ERR_INVALID, "cuzk_init: bad device index");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(CUZK_ERR_CUDA, "cuzk_init: sm_100a (Blackwell B200) device required");
  g_sm_count = prop.multiProcessorCount;
  // keep stream-ordered scratch in the pool between calls
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  // round constants: generate on the device with the reference formula, check the <2^64 fast-path assumption
  void *d = nullptr;
  CK(cudaMalloc(&d, sizeof g_host_rc));
  gen_round_constants_kernel<<<2, 96>>>(reinterpret_cast<uint4 *>(d));
  int rc = check_launch("gen_round_constants_kernel");
  if (rc) { cudaFree(d); return rc; }
  e = cudaMemcpy(g_host_rc, d, sizeof g_host_rc, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(round constants)");
  u32 packed[kRounds * 3][2];
  for (int i = 0; i < kRounds * 3; ++i) {
    if (g_host_rc[4 * i + 1] | g_host_rc[4 * i + 2] | g_host_rc[4 * i + 3])
      return fail(CUZK_ERR_CONSTANTS, "round constant does not fit 64 bits");
    packed[i][0] = (u32)g_host_rc[4 * i];
    packed[i][1] = (u32)(g_host_rc[4 * i] >> 32);
  }
  CK(cudaMemcpyToSymbol(c_rc, packed, sizeof packed));
  static const uint64_t m[9] = {7, 23, 8, 26, 5, 4, 15, 20, 9};
  memset(g_host_mds, 0, sizeof g_host_mds);
  for (int i = 0; i < 9; ++i) g_host_mds[4 * i] = m[i];
  if ((rc = subtree_streams_start())) return rc;
  g_device = device;
  g_refcount = 1;
  return CUZK_OK;
}

int cuzk_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_refcount <= 0) return CUZK_OK;
  if (--g_refcount == 0) {
    cudaDeviceSynchronize();
    for (int a = 0; a < 9; ++a) {
      if (g_d_pad[a]) cudaFree(g_d_pad[a]);
      g_d_pad[a] = nullptr;
      g_pad_levels[a] = 0;
    }
    {
      std::lock_guard<std::mutex> lk2(g_hp_mu);
      hp_stop();
      pin_stop();
      subtree_streams_stop();
    }
    g_device = -1;
  }
  return CUZK_OK;
}

int cuzk_poseidon_constants(uint64_t *rc_out, uint64_t *mds_out) {
  int rc = require_init();
  if (rc) return rc;
  if (rc_out) memcpy(rc_out, g_host_rc, sizeof g_host_rc);
  if (mds_out) memcpy(mds_out, g_host_mds, sizeof g_host_mds);
  return CUZK_OK;
}

int cuzk_fr_batch(int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n, int mem, void *stream) {
  int rc = require_init();
  if (rc) return rc;
  if (op < 0 || op > CUZK_FR_POW5) return fail(CUZK_ERR_INVALID, "unknown field op");
  if (n == 0) return CUZK_OK;
  bool binary = op <= CUZK_FR_MUL;
  if (!a || !out || (binary && !b)) return fail(CUZK_ERR_INVALID, "null pointer");
  if (mem == CUZK_MEM_DEVICE) return fr_batch_dev(op, a, b, out, n, S(stream));
  const void *in[2] = {a, b};
  const size_t in_bytes[2] = {32, 32};
  return host_pipeline(n, kCheapChunk, binary ? 2 : 1, in, in_bytes, out, 32, false,
                       [&](cudaStream_t st, void **d_in, void *d_out, size_t m) {
                         return fr_batch_dev(op, static_cast<const uint64_t *>(d_in[0]), static_cast<const uint64_t *>(d_in[1]),
                                             static_cast<uint64_t *>(d_out), m, st);
                       });
}

int cuzk_poseidon_hash_single(const uint64_t *in, uint64_t *out, size_t n, int mem, void *stream) {
  int rc = require_init();
  if (rc) return rc;
  if (n == 0) return CUZK_OK;
  if (!in || !out) return fail(CUZK_ERR_INVALID, "null pointer");
  auto run = [](cudaStream_t st, const void *din, void *dout, size_t m) {
    hash_single_kernel<<<grid_for(m, kBlock), kBlock, 0, st>>>(static_cast<const uint4 *>(din), static_cast<uint4 *>(dout), m);
    return check_launch("hash_single_kernel");
  };
  if (mem == CUZK_MEM_DEVICE) return run(S(stream), in, out, n);
  const void *ins[1] = {in};
  const size_t in_bytes[1] = {32};
  return host_pipeline(n, kHashChunk, 1, ins, in_bytes, out, 32, false,
                       [&](cudaStream_t st, void **d_in, void *d_out, size_t m) { return run(st, d_in[0], d_out, m); });
}

int cuzk_poseidon_hash_pairs(const uint64_t *left, const uint64_t *right, uint64_t *out, size_t n, int mem, void *stream) {
  int rc = require_init();
  if (rc) return rc;
  if (n == 0) return CUZK_OK;
  if (!left || !right || !out) return fail(CUZK_ERR_INVALID, "null pointer");
  auto run = [](cudaStream_t st, const void *dl, const void *dr, void *dout, size_t m) {
    hash_pairs_kernel<<<grid_for(m, kBlock), kBlock, 0, st>>>(static_cast<const uint4 *>(dl), static_cast<const uint4 *>(dr),
                                                             static_cast<uint4 *>(dout), m);
    return check_launch("hash_pairs_kernel");
  };
  if (mem == CUZK_MEM_DEVICE) return run(S(stream), left, right, out, n);
  const void *ins[2] = {left, right};
  const size_t in_bytes[2] = {32, 32};
  return host_pipeline(n, kHashChunk, 2, ins, in_bytes, out, 32, false,
                       [&](cudaStream_t st, void **d_in, void *d_out, size_t m) { return run(st, d_in[0], d_in[1], d_out, m); });
}

int cuzk_poseidon_permutation(uint64_t *states, size_t n, int mem, void *stream) {
  int rc = require_init();
  if (rc) return rc;
  if (n == 0) return CUZK_OK;
  if (!states) return fail(CUZK_ERR_INVALID, "null pointer");
  auto run = [](cudaStream_t st, void *d, size_t m) {
    permutation_kernel<<<grid_for(m, kBlock), kBlock, 0, st>>>(static_cast<uint4 *>(d), m);
    return check_launch("permutation_kernel");
  };
  if (mem == CUZK_MEM_DEVICE) return run(S(stream), states, n);
  const void *ins[1] = {states};
  const size_t in_bytes[1] = {96};
  return host_pipeline(n, kHashChunk, 1, ins, in_bytes, states, 96, true,
                       [&](cudaStream_t st, void **d_in, void *, size_t m) { return run(st, d_in[0], m); });
}

int cuzk_debug_mds_layer(uint64_t *states, size_t n, int mode, void *stream) {
  int rc = require_init();
  if (rc) return rc;
  if (n == 0) return CUZK_OK;
  debug_mds_kernel<<<grid_for(n, kBlock), kBlock, 0, S(stream)>>>(reinterpret_cast<uint4 *>(states), n, mode);
  return check_launch("debug_mds_kernel");
}

int cuzk_debug_fast_ops(int op, const uint64_t *a, const uint64_t *b, uint64_t *out, uint32_t *flags, size_t n, void *stream) {
  int rc = require_init();
  if (rc) return rc;
  if (op < 0 || op > 3) return fail(CUZK_ERR_INVALID, "unknown fast op");
  if (n == 0) return CUZK_OK;
  if (!a || !out || !flags || (op == 1 && !b)) return fail(CUZK_ERR_INVALID, "null pointer");
  debug_fast_ops_kernel<<<grid_for(n, kBlock), kBlock, 0, S(stream)>>>(op, reinterpret_cast<const uint4 *>(a), reinterpret_cast<const uint4 *>(b),
                                                                      reinterpret_cast<uint4 *>(out), flags, n);
  return check_launch("debug_fast_ops_kernel");
}

int cuzk_poseidon_sponge(const uint64_t *in, size_t width, uint64_t ds, uint64_t *out, size_t n, int mem, void *stream) {
  int rc = require_init();
  if (rc) return rc;
  if (width > 64) return fail(CUZK_ERR_INVALID, "sponge width must be <= 64");
  if (n == 0) return CUZK_OK;
  if (!out || (width && !in)) return fail(CUZK_ERR_INVALID, "null pointer");
  auto run = [&](cudaStream_t st, const void *din, void *dout, size_t m) {
    sponge_kernel<<<grid_for(m, kBlock), kBlock, 0, st>>>(static_cast<const uint4 *>(din), (int)width, (u32)ds, (u32)(ds >> 32),
                                                         static_cast<uint4 *>(dout), m);
    return check_launch("sponge_kernel");
  };
  if (mem == CUZK_MEM_DEVICE) return run(S(stream), in, out, n);
  if (width == 0) {  // zero permutations: every output is the zero element (poseidon.cpp:103-126)
    memset(out, 0, n * 32);
    return CUZK_OK;
  }
  const void *ins[1] = {in};
  const size_t in_bytes[1] = {32 * width};
  return host_pipeline(n, kHashChunk, 1, ins, in_bytes, out, 32, false,
                       [&](cudaStream_t st, void **d_in, void *d_out, size_t m) { return run(st, d_in[0], d_out, m); });
}

// ---- Merkle geometry ----
size_t cuzk_merkle_padded_leaves(size_t n, unsigned arity) {
  if (arity < 2) return 0;
  size_t p = 1;
  while (p < n) p *= arity;
  return p;
}
size_t cuzk_merkle_num_levels(size_t n, unsigned arity) {
  if (n == 0 || arity < 2) return 0;
  size_t p = cuzk_merkle_padded_leaves(n, arity), lv = 1;
  while (p > 1) { p /= arity; ++lv; }
  return lv;
}
size_t cuzk_merkle_total_nodes(size_t n, unsigned arity) {
  if (n == 0 || arity < 2) return 0;
  size_t p = cuzk_merkle_padded_leaves(n, arity), tot = p;
  while (p > 1) { p /= arity; tot += p; }
  return tot;
}
size_t cuzk_merkle_tree_height(size_t leaf_count, unsigned arity) {
  if (leaf_count <= 1) return 1;
  return (size_t)std::ceil(std::log((double)leaf_count) / std::log((double)arity)) + 1;
}

int cuzk_merkle_empty_hash(unsigned arity, uint64_t out[4]) { return cuzk_merkle_padding_root(arity, 0, out); }

int cuzk_merkle_padding_root(unsigned arity, unsigned height, uint64_t out[4]) {
  int rc = require_init();
  if (rc) return rc;
  if ((rc = check_arity(arity))) return rc;
  if (height >= (unsigned)kMaxPadLevels) return fail(CUZK_ERR_INVALID, "padding level too high");
  if ((rc = ensure_padding(arity, (int)height + 1))) return rc;
  memcpy(out, g_h_pad[arity][height], 32);
  return CUZK_OK;
}

int cuzk_merkle_build(const uint64_t *leaves, size_t n, unsigned arity, uint64_t *levels_out, int mem, void *stream) {
  int rc = require_init();
  if (rc) return rc;
  if ((rc = check_arity(arity))) return rc;
  if (n == 0) return fail(CUZK_ERR_INVALID, "cuzk_merkle_build: n must be >= 1");
  if (!leaves || !levels_out) return fail(CUZK_ERR_INVALID, "null pointer");
  cudaStream_t st = S(stream);
  if (mem == CUZK_MEM_DEVICE) return merkle_build_dev(leaves, n, arity, levels_out, st);
  std::lock_guard<std::mutex> lk(g_hp_mu);
  size_t tot = cuzk_merkle_total_nodes(n, arity);
  void *dl, *dv;
  if ((rc = ws_get(0, n * 32, &dl)) || (rc = ws_get(1, tot * 32, &dv))) return rc;
  if ((rc = bulk_upload(dl, leaves, n * 32, st))) return rc;
  rc = merkle_build_dev(static_cast<uint64_t *>(dl), n, arity, static_cast<uint64_t *>(dv), st);
  if (rc) return rc;
  return bulk_download(levels_out, dv, tot * 32, st);
}

int cuzk_merkle_build_batch(const uint64_t *leaves, size_t n, size_t num_trees, unsigned arity, uint64_t *levels_out, int mem,
                            void *stream) {
  int rc = require_init();
  if (rc) return rc;
  if ((rc = check_arity(arity))) return rc;
  if (num_trees == 0) return CUZK_OK;
  if (n == 0) return fail(CUZK_ERR_INVALID, "cuzk_merkle_build_batch: n must be >= 1");
  if (!leaves || !levels_out) return fail(CUZK_ERR_INVALID, "null pointer");
  cudaStream_t st = S(stream);
  if (mem == CUZK_MEM_DEVICE) return merkle_build_dev(leaves, n, arity, levels_out, st, num_trees);
  std::lock_guard<std::mutex> lk(g_hp_mu);
  const size_t tot = cuzk_merkle_total_nodes(n, arity) * num_trees;
  void *dl, *dv;
  if ((rc = ws_get(0, n * num_trees * 32, &dl)) || (rc = ws_get(1, tot * 32, &dv))) return rc;
  if ((rc = bulk_upload(dl, leaves, n * num_trees * 32, st))) return rc;
  rc = merkle_build_dev(static_cast<uint64_t *>(dl), n, arity, static_cast<uint64_t *>(dv), st, num_trees);
  if (rc) return rc;
  return bulk_download(levels_out, dv, tot * 32, st);
}

int cuzk_merkle_subtree_roots(const uint64_t *leaves, size_t n, unsigned arity, unsigned height, size_t count,
                              uint64_t *roots_out, int mem, void *stream) {
  int rc = require_init();
  if (rc) return rc;
  if ((rc = check_arity(arity))) return rc;
  if (count == 0) return CUZK_OK;
  if (mem != CUZK_MEM_DEVICE) return fail(CUZK_ERR_INVALID, "cuzk_merkle_subtree_roots: device pointers only");
  if (!roots_out || (n && !leaves)) return fail(CUZK_ERR_INVALID, "null pointer");
  size_t span = 1;
  for (unsigned i = 0; i < height; ++i) {
    if (span > (~(size_t)0) / arity) return fail(CUZK_ERR_INVALID, "subtree too tall");
    span *= arity;
  }
  if (n > count * span) return fail(CUZK_ERR_INVALID, "more leaves than the subtrees hold");
  return subtree_roots_dev(leaves, n, arity, height, count, roots_out, S(stream));
}

int cuzk_merkle_top_root(const uint64_t *nodes, size_t count, unsigned arity, uint64_t *root_out, int mem, void *stream) {
  int rc = require_init();
  if (rc) return rc;
  if ((rc = check_arity(arity))) return rc;
  if (count == 0 || !nodes || !root_out) return fail(CUZK_ERR_INVALID, "bad arguments");
  size_t p = 1;
  unsigned h = 0;
  while (p < count) { p *= arity; ++h; }
  if (p != count) return fail(CUZK_ERR_INVALID, "count must be a power of arity");
  cudaStream_t st = S(stream);
  if (mem == CUZK_MEM_DEVICE) return subtree_roots_dev(nodes, count, arity, h, 1, root_out, st);
  std::lock_guard<std::mutex> lk(g_hp_mu);
  void *dn, *dr;
  if ((rc = ws_get(0, count * 32, &dn)) || (rc = ws_get(1, 32, &dr))) return rc;
  CK(cudaMemcpyAsync(dn, nodes, count * 32, cudaMemcpyHostToDevice, st));
  rc = subtree_roots_dev(static_cast<uint64_t *>(dn), count, arity, h, 1, static_cast<uint64_t *>(dr), st);
  if (rc) return rc;
  CK(cudaMemcpyAsync(root_out, dr, 32, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return CUZK_OK;
}

int cuzk_merkle_prove_batch(const uint64_t *levels, size_t n, unsigned arity, const uint64_t *indices, size_t num_proofs,
                            uint64_t *siblings_out, uint32_t *positions_out, int mem, void *stream) {
  int rc = require_init();
  if (rc) return rc;
  if ((rc = check_arity(arity))) return rc;
  if (num_proofs == 0) return CUZK_OK;
  if (n == 0) return fail(CUZK_ERR_INVALID, "empty tree");
  size_t nlv = cuzk_merkle_num_levels(n, arity) - 1;
  if (nlv == 0) return CUZK_OK;  // single-leaf tree: proofs have no levels
  if (!levels || !indices || !siblings_out || !positions_out) return fail(CUZK_ERR_INVALID, "null pointer");
  size_t padded = cuzk_merkle_padded_leaves(n, arity);
  cudaStream_t st = S(stream);
  size_t threads = num_proofs * nlv;
  if (mem == CUZK_MEM_DEVICE) {
    merkle_prove_kernel<<<grid_for(threads, 256), 256, 0, st>>>(reinterpret_cast<const uint4 *>(levels), n, padded, (int)arity, (int)nlv,
                                                                indices, num_proofs, reinterpret_cast<uint4 *>(siblings_out), positions_out);
    return check_launch("merkle_prove_kernel");
  }
  std::lock_guard<std::mutex> lk(g_hp_mu);
  size_t tot = cuzk_merkle_total_nodes(n, arity);
  void *dl, *di, *ds, *dp;
  if ((rc = ws_get(0, tot * 32, &dl)) || (rc = ws_get(1, num_proofs * 8, &di)) || (rc = ws_get(2, threads * (arity - 1) * 32, &ds)) ||
      (rc = ws_get(3, threads * 4, &dp)))
    return rc;
  CK(cudaMemcpyAsync(dl, levels, tot * 32, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(di, indices, num_proofs * 8, cudaMemcpyHostToDevice, st));
  merkle_prove_kernel<<<grid_for(threads, 256), 256, 0, st>>>(static_cast<uint4 *>(dl), n, padded, (int)arity, (int)nlv,
                                                              static_cast<u64 *>(di), num_proofs, static_cast<uint4 *>(ds),
                                                              static_cast<u32 *>(dp));
  if ((rc = check_launch("merkle_prove_kernel"))) return rc;
  if ((rc = bulk_download(siblings_out, ds, threads * (arity - 1) * 32, st))) return rc;
  return bulk_download(positions_out, dp, threads * 4, st);
}

int cuzk_merkle_verify_batch(const uint64_t *leaf_values, const uint64_t *siblings, const uint32_t *positions, size_t levels,
                             unsigned arity, const uint64_t *root, uint8_t *results_out, size_t num_proofs, int mem, void *stream) {
  int rc = require_init();
  if (rc) return rc;
  if ((rc = check_arity(arity))) return rc;
  if (num_proofs == 0) return CUZK_OK;
  if (!leaf_values || !root || !results_out || (levels && (!siblings || !positions))) return fail(CUZK_ERR_INVALID, "null pointer");
  cudaStream_t st = S(stream);
  if (mem == CUZK_MEM_DEVICE) {
    merkle_verify_kernel<<<grid_for(num_proofs, kBlock), kBlock, 0, st>>>(reinterpret_cast<const uint4 *>(leaf_values),
                                                                         reinterpret_cast<const uint4 *>(siblings), positions, (int)levels,
                                                                         (int)arity, reinterpret_cast<const uint4 *>(root), uint4{}, uint4{},
                                                                         results_out, num_proofs);
    return check_launch("merkle_verify_kernel");
  }
  uint4 root_lo, root_hi;   // host-buffer call: the 32-byte root travels as a kernel argument
  memcpy(&root_lo, root, 16);
  memcpy(&root_hi, root + 2, 16);
  const void *ins[3] = {leaf_values, siblings, positions};
  const size_t in_bytes[3] = {32, levels * (arity - 1) * 32, levels * 4};
  // proofs are independent: chunk them like hashes (each costs levels x ceil(arity/2) permutations)
  return host_pipeline(num_proofs, kHashChunk, levels ? 3 : 1, ins, in_bytes, results_out, 1, false,
                       [&](cudaStream_t s2, void **d_in, void *d_out, size_t m) {
                         merkle_verify_kernel<<<grid_for(m, kBlock), kBlock, 0, s2>>>(
                             static_cast<const uint4 *>(d_in[0]), static_cast<const uint4 *>(d_in[1]), static_cast<const u32 *>(d_in[2]),
                             (int)levels, (int)arity, nullptr, root_lo, root_hi, static_cast<uint8_t *>(d_out), m);
                         return check_launch("merkle_verify_kernel");
                       });
}

// ---- device-resident tree handle ----
struct cuzk_tree {
  uint64_t *levels = nullptr;   // device, level-major, cuzk_merkle_total_nodes elements
  size_t n = 0, padded = 0, total = 0, nlevels = 0;
  unsigned arity = 0;
  int device = 0;
};

int cuzk_tree_build(const uint64_t *leaves, size_t n, unsigned arity, int mem, void *stream, cuzk_tree_t **out) {
  int rc = require_init();
  if (rc) return rc;
  if ((rc = check_arity(arity))) return rc;
  if (!out) return fail(CUZK_ERR_INVALID, "null pointer");
  *out = nullptr;
  if (n == 0 || !leaves) return fail(CUZK_ERR_INVALID, "cuzk_tree_build: needs at least one leaf");
  cuzk_tree *t = new (std::nothrow) cuzk_tree;
  if (!t) return fail(CUZK_ERR_INVALID, "out of host memory");
  t->n = n;
  t->arity = arity;
  t->padded = cuzk_merkle_padded_leaves(n, arity);
  t->total = cuzk_merkle_total_nodes(n, arity);
  t->nlevels = cuzk_merkle_num_levels(n, arity);
  t->device = g_device;
  cudaStream_t st = S(stream);
  cudaError_t e = cudaMalloc(&t->levels, t->total * 32);
  if (e != cudaSuccess) { delete t; return cuda_fail(e, "cudaMalloc(tree levels)"); }
  if (mem == CUZK_MEM_DEVICE) {
    rc = merkle_build_dev(leaves, n, arity, t->levels, st);
  } else {
    std::lock_guard<std::mutex> lk(g_hp_mu);
    void *dl;
    if (!(rc = ws_get(0, n * 32, &dl)) && !(rc = bulk_upload(dl, leaves, n * 32, st))) {
      rc = merkle_build_dev(static_cast<uint64_t *>(dl), n, arity, t->levels, st);
      if (!rc && (e = cudaStreamSynchronize(st)) != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize");
    }
  }
  if (rc) {
    cudaFree(t->levels);
    delete t;
    return rc;
  }
  *out = t;
  return CUZK_OK;
}

int cuzk_tree_free(cuzk_tree_t *t) {
  if (!t) return CUZK_OK;
  cudaError_t e = cudaFree(t->levels);
  delete t;
  if (e != cudaSuccess) return cuda_fail(e, "cudaFree(tree levels)");
  return CUZK_OK;
}

size_t cuzk_tree_leaf_count(const cuzk_tree_t *t) { return t ? t->n : 0; }
size_t cuzk_tree_num_levels(const cuzk_tree_t *t) { return t ? t->nlevels : 0; }
size_t cuzk_tree_total_nodes(const cuzk_tree_t *t) { return t ? t->total : 0; }
unsigned cuzk_tree_arity(const cuzk_tree_t *t) { return t ? t->arity : 0; }
const uint64_t *cuzk_tree_device_levels(const cuzk_tree_t *t) { return t ? t->levels : nullptr; }

int cuzk_tree_root(const cuzk_tree_t *t, uint64_t *root_out, int mem, void *stream) {
  if (!t || !root_out) return fail(CUZK_ERR_INVALID, "null pointer");
  cudaStream_t st = S(stream);
  const uint64_t *src = t->levels + (t->total - 1) * 4;
  if (mem == CUZK_MEM_DEVICE) {
    CK(cudaMemcpyAsync(root_out, src, 32, cudaMemcpyDeviceToDevice, st));
    return CUZK_OK;
  }
  CK(cudaMemcpyAsync(root_out, src, 32, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return CUZK_OK;
}

int cuzk_tree_levels(const cuzk_tree_t *t, uint64_t *levels_out, int mem, void *stream) {
  if (!t || !levels_out) return fail(CUZK_ERR_INVALID, "null pointer");
  cudaStream_t st = S(stream);
  if (mem == CUZK_MEM_DEVICE) {
    CK(cudaMemcpyAsync(levels_out, t->levels, t->total * 32, cudaMemcpyDeviceToDevice, st));
    return CUZK_OK;
  }
  std::lock_guard<std::mutex> lk(g_hp_mu);
  return bulk_download(levels_out, t->levels, t->total * 32, st);
}

int cuzk_tree_level(const cuzk_tree_t *t, size_t level, uint64_t *level_out, int mem, void *stream) {
  if (!t || !level_out) return fail(CUZK_ERR_INVALID, "null pointer");
  if (level >= t->nlevels) return fail(CUZK_ERR_INVALID, "cuzk_tree_level: no such level");
  size_t off = 0, width = t->padded;
  for (size_t l = 0; l < level; ++l) {
    off += width;
    width /= t->arity;
  }
  cudaStream_t st = S(stream);
  if (mem == CUZK_MEM_DEVICE) {
    CK(cudaMemcpyAsync(level_out, t->levels + 4 * off, width * 32, cudaMemcpyDeviceToDevice, st));
    return CUZK_OK;
  }
  std::lock_guard<std::mutex> lk(g_hp_mu);
  return bulk_download(level_out, t->levels + 4 * off, width * 32, st);
}

int cuzk_tree_prove_batch(const cuzk_tree_t *t, const uint64_t *indices, size_t num_proofs, uint64_t *siblings_out,
                          uint32_t *positions_out, int mem, void *stream) {
  int rc = require_init();
  if (rc) return rc;
  if (!t) return fail(CUZK_ERR_INVALID, "null tree");
  if (num_proofs == 0 || t->nlevels <= 1) return CUZK_OK;
  if (!indices || !siblings_out || !positions_out) return fail(CUZK_ERR_INVALID, "null pointer");
  const size_t nlv = t->nlevels - 1, threads = num_proofs * nlv, sib_bytes = threads * (t->arity - 1) * 32;
  cudaStream_t st = S(stream);
  auto launch = [&](const u64 *di, uint4 *ds, u32 *dp) {
    merkle_prove_kernel<<<grid_for(threads, 256), 256, 0, st>>>(reinterpret_cast<const uint4 *>(t->levels), t->n, t->padded, (int)t->arity,
                                                                (int)nlv, di, num_proofs, ds, dp);
    return check_launch("merkle_prove_kernel");
  };
  if (mem == CUZK_MEM_DEVICE) return launch(indices, reinterpret_cast<uint4 *>(siblings_out), positions_out);
  std::lock_guard<std::mutex> lk(g_hp_mu);
  void *di, *ds, *dp;
  if ((rc = ws_get(1, num_proofs * 8, &di)) || (rc = ws_get(2, sib_bytes, &ds)) || (rc = ws_get(3, threads * 4, &dp))) return rc;
  CK(cudaMemcpyAsync(di, indices, num_proofs * 8, cudaMemcpyHostToDevice, st));
  if ((rc = launch(static_cast<u64 *>(di), static_cast<uint4 *>(ds), static_cast<u32 *>(dp)))) return rc;
  if ((rc = bulk_download(siblings_out, ds, sib_bytes, st))) return rc;
  return bulk_download(positions_out, dp, threads * 4, st);
}

int cuzk_tree_verify_batch(const cuzk_tree_t *t, const uint64_t *leaf_values, const uint64_t *siblings, const uint32_t *positions,
                           uint8_t *results_out, size_t num_proofs, int mem, void *stream) {
  if (!t) return fail(CUZK_ERR_INVALID, "null tree");
  const size_t nlv = t->nlevels - 1;
  if (mem == CUZK_MEM_DEVICE)
    return cuzk_merkle_verify_batch(leaf_values, siblings, positions, nlv, t->arity, t->levels + (t->total - 1) * 4, results_out, num_proofs,
                                    mem, stream);
  uint64_t root[4];
  int rc = cuzk_tree_root(t, root, CUZK_MEM_HOST, stream);
  if (rc) return rc;
  return cuzk_merkle_verify_batch(leaf_values, siblings, positions, nlv, t->arity, root, results_out, num_proofs, mem, stream);
}

int cuzk_tree_update_leaves(cuzk_tree_t *t, const uint64_t *indices, const uint64_t *values, size_t count, int mem, void *stream) {
  int rc = require_init();
  if (rc) return rc;
  if (!t) return fail(CUZK_ERR_INVALID, "null tree");
  if (count == 0) return CUZK_OK;
  if (!indices || !values) return fail(CUZK_ERR_INVALID, "null pointer");
  cudaStream_t st = S(stream);
  const u64 *di = indices;
  const uint4 *dv = reinterpret_cast<const uint4 *>(values);
  std::unique_lock<std::mutex> lk(g_hp_mu, std::defer_lock);
  if (mem != CUZK_MEM_DEVICE) {
    for (size_t q = 0; q < count; ++q)   // NaryMerkleTree::update_leaf throws std::out_of_range here (merkle_tree.cpp:296-298)
      if (indices[q] >= t->n) return fail(CUZK_ERR_INVALID, "cuzk_tree_update_leaves: leaf index out of range");
    lk.lock();
    void *wi, *wv;
    if ((rc = ws_get(1, count * 8, &wi)) || (rc = ws_get(2, count * 32, &wv))) return rc;
    CK(cudaMemcpyAsync(wi, indices, count * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(wv, values, count * 32, cudaMemcpyHostToDevice, st));
    di = static_cast<const u64 *>(wi);
    dv = static_cast<const uint4 *>(wv);
  }
  uint4 *cur = reinterpret_cast<uint4 *>(t->levels);
  merkle_write_leaves_kernel<<<grid_for(count, 256), 256, 0, st>>>(cur, di, dv, count);
  if ((rc = check_launch("merkle_write_leaves_kernel"))) return rc;
  size_t p = t->padded;
  u64 divisor = 1;
  while (p > 1) {
    divisor *= t->arity;
    merkle_update_level_kernel<<<grid_for(count, kBlock), kBlock, 0, st>>>(cur, cur + 2 * p, di, count, divisor, (int)t->arity);
    if ((rc = check_launch("merkle_update_level_kernel"))) return rc;
    cur += 2 * p;
    p /= t->arity;
  }
  if (mem != CUZK_MEM_DEVICE) CK(cudaStreamSynchronize(st));
  return CUZK_OK;
}

int cuzk_synth_elements(uint64_t *out, size_t n, uint64_t seed, uint64_t start, int canonical, void *stream) {
  if (n == 0) return CUZK_OK;
  synth_elements_kernel<<<grid_for(n, 256), 256, 0, S(stream)>>>(out, n, seed, start, canonical);
  return check_launch("synth_elements_kernel");
}
int cuzk_synth_u64_leaves(uint64_t *out, size_t n, uint64_t seed, uint64_t start, void *stream) {
  if (n == 0) return CUZK_OK;
  synth_u64_leaves_kernel<<<grid_for(n, 256), 256, 0, S(stream)>>>(out, n, seed, start);
  return check_launch("synth_u64_leaves_kernel");
}

}  // extern "C"
