// cuzk_kernels.cu -- the extern "C" layer of libcuzk_b200.so (include/cuzk_b200.h): one translation unit made of
//   fr.cuh        BN254 Fr "reference arithmetic" on 8 x 32-bit limbs
//   poseidon.cuh  the permutation (fast path + exact fallback, FP64-pipe MDS layer) and the sponge
//   kernels.cuh   every __global__ kernel
//   host_path.cuh staging pipeline, copy pool, padding constants, Merkle level scheduling
//   this file     library state and the C entry points
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

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
std::mutex g_mu;                            // cuzk_init / cuzk_shutdown
std::atomic<uint64_t> g_launches{0};
uint64_t g_host_rc[kRounds * 3 * 4];        // constants of the hash function: the same on every device
uint64_t g_host_mds[9 * 4];

// padding constants E_l per arity: E_0 = empty_hash(arity), E_{l+1} = hash_multiple(arity x E_l).  Host copy: constants of
// the hash function, kept for the life of the process and shared by all devices (each device context has its own upload).
constexpr int kMaxPadLevels = 41;
uint64_t g_h_pad[9][kMaxPadLevels][4];
int g_h_pad_levels[9] = {0};                // how many levels the host copy holds
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

// grid for n one-thread units at `block` threads
inline unsigned grid_for(size_t n, unsigned block) {
  size_t g = (n + block - 1) / block;
  if (g == 0) g = 1;
  return (unsigned)g;
}

}  // namespace

#include "kernels.cuh"
#include "host_path.cuh"

// ------------------------------------------------------------------------------------------------
// extern "C"
// ------------------------------------------------------------------------------------------------
namespace {

// round constants onto the current device: generated there with the reference formula (poseidon.cpp:33-44), checked against
// the <2^64 assumption of the fast paths, uploaded to __constant__ memory
int upload_constants() {
  void *d = nullptr;
  CK(cudaMalloc(&d, sizeof g_host_rc));
  gen_round_constants_kernel<<<2, 96>>>(reinterpret_cast<uint4 *>(d));
  int rc = check_launch("gen_round_constants_kernel");
  if (rc) { cudaFree(d); return rc; }
  uint64_t host_rc[kRounds * 3 * 4];
  cudaError_t e = cudaMemcpy(host_rc, d, sizeof host_rc, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy(round constants)");
  u32 packed[kRounds * 3][2];
  for (int i = 0; i < kRounds * 3; ++i) {
    if (host_rc[4 * i + 1] | host_rc[4 * i + 2] | host_rc[4 * i + 3])
      return fail(CUZK_ERR_CONSTANTS, "round constant does not fit 64 bits");
    packed[i][0] = (u32)host_rc[4 * i];
    packed[i][1] = (u32)(host_rc[4 * i] >> 32);
  }
  CK(cudaMemcpyToSymbol(c_rc, packed, sizeof packed));
  memcpy(g_host_rc, host_rc, sizeof g_host_rc);
  static const uint64_t m[9] = {7, 23, 8, 26, 5, 4, 15, 20, 9};
  memset(g_host_mds, 0, sizeof g_host_mds);
  for (int i = 0; i < 9; ++i) g_host_mds[4 * i] = m[i];
  return CUZK_OK;
}

}  // namespace

extern "C" {

const char *cuzk_last_error(void) { return g_err.c_str(); }
const char *cuzk_version(void) { return "cuzk_b200 0.3 (sm_100a)"; }
uint64_t cuzk_launch_count(void) { return g_launches.load(); }
size_t cuzk_debug_set_coop_max(size_t units) { return g_coop_max.exchange(units); }
size_t cuzk_debug_set_coop_wide_max(size_t units) { return g_coop_wide_max.exchange(units); }
size_t cuzk_debug_set_direct_max(size_t bytes) { return g_direct_max.exchange(bytes); }
void cuzk_debug_set_build_plan(int groups, int streams, size_t group_coop_max) {
  if (groups >= 1) g_build_groups.store(groups);
  if (streams >= 1) g_build_streams.store(streams > kSubtreeStreams ? kSubtreeStreams : streams);
  g_group_coop_max.store(group_coop_max);
}
uint64_t cuzk_debug_fallback_count(void) {
  CtxGuard guard;
  if (!guard.get()) return ~0ull;
  unsigned long long v = 0;
  if (cudaMemcpyFromSymbol(&v, g_exact_fallbacks, sizeof v) != cudaSuccess) return ~0ull;
  return v;
}

int cuzk_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int cuzk_is_initialized(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (int d = 0; d < kMaxDevices; ++d)
    if (g_ctx[d].refcount > 0) return 1;
  return 0;
}
int cuzk_is_initialized_on(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  return device >= 0 && device < kMaxDevices && g_ctx[device].refcount > 0 ? 1 : 0;
}

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
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return fail(CUZK_ERR_CUDA, "cuzk_init: no CUDA device (this library has no CPU fallback)");
  if (device < 0 || device >= count || device >= kMaxDevices) return fail(CUZK_ERR_INVALID, "cuzk_init: bad device index");
  CK(cudaSetDevice(device));   // the calling thread's current device from here on, as with the reference's cudaSetDevice(0)
  Ctx &c = g_ctx[device];
  if (c.refcount > 0) {
    ++c.refcount;
    return CUZK_OK;
  }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(CUZK_ERR_CUDA, "cuzk_init: sm_100a (Blackwell B200) device required");
  c.sm_count = prop.multiProcessorCount;
  // keep stream-ordered scratch in the pool between calls
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  int rc = upload_constants();
  if (rc) return rc;
  if ((rc = subtree_streams_start(c))) return rc;
  c.device = device;
  c.refcount = 1;
  return CUZK_OK;
}

// releases the calling thread's current device (or the only initialised one)
int cuzk_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  CtxGuard guard;
  Ctx *c = guard.get();
  if (!c) return CUZK_OK;
  if (--c->refcount == 0) {
    cudaDeviceSynchronize();
    for (int a = 0; a < 9; ++a) {
      if (c->d_pad[a]) cudaFree(c->d_pad[a]);
      c->d_pad[a] = nullptr;
      c->pad_levels[a] = 0;
    }
    {
      std::lock_guard<std::mutex> lk2(c->hp_mu);
      hp_stop(*c);
      subtree_streams_stop(*c);
    }
    c->device = -1;
  }
  return CUZK_OK;
}

int cuzk_poseidon_constants(uint64_t *rc_out, uint64_t *mds_out) {
  CUZK_CTX(c);
  (void)c;
  if (rc_out) memcpy(rc_out, g_host_rc, sizeof g_host_rc);
  if (mds_out) memcpy(mds_out, g_host_mds, sizeof g_host_mds);
  return CUZK_OK;
}

int cuzk_fr_batch(int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n, int mem, void *stream) {
  CUZK_CTX(c);
  if (op < 0 || op > CUZK_FR_POW5) return fail(CUZK_ERR_INVALID, "unknown field op");
  if (n == 0) return CUZK_OK;
  bool binary = op <= CUZK_FR_MUL;
  if (!a || !out || (binary && !b)) return fail(CUZK_ERR_INVALID, "null pointer");
  if (mem == CUZK_MEM_DEVICE) return fr_batch_dev(op, a, b, out, n, S(stream));
  const void *in[2] = {a, b};
  const size_t in_bytes[2] = {32, 32};
  return host_pipeline(c, n, kCheapChunk, binary ? 2 : 1, in, in_bytes, out, 32, false,
                       [&](cudaStream_t st, void **d_in, void *d_out, size_t m) {
                         return fr_batch_dev(op, static_cast<const uint64_t *>(d_in[0]), static_cast<const uint64_t *>(d_in[1]),
                                             static_cast<uint64_t *>(d_out), m, st);
                       });
}

// chunk size of a host-buffer hashing call: batches the cooperative kernels serve go as one chunk (they are latency-bound:
// cutting them would only add launches)
static size_t hash_chunk(size_t n) { return use_coop(n) ? n : kHashChunk; }

int cuzk_poseidon_hash_single(const uint64_t *in, uint64_t *out, size_t n, int mem, void *stream) {
  CUZK_CTX(c);
  if (n == 0) return CUZK_OK;
  if (!in || !out) return fail(CUZK_ERR_INVALID, "null pointer");
  auto run = [](cudaStream_t st, const void *din, void *dout, size_t m) {
    if (const CoopKind kind = coop_kind(m)) {
      CUZK_COOP_LAUNCH(kind, coop_hash_single_kernel, m, st, static_cast<const uint4 *>(din), static_cast<uint4 *>(dout), m);
      return check_launch("coop_hash_single_kernel");
    }
    hash_single_kernel<<<grid_for(m, kBlock), kBlock, 0, st>>>(static_cast<const uint4 *>(din), static_cast<uint4 *>(dout), m);
    return check_launch("hash_single_kernel");
  };
  if (mem == CUZK_MEM_DEVICE) return run(S(stream), in, out, n);
  const void *ins[1] = {in};
  const size_t in_bytes[1] = {32};
  return host_pipeline(c, n, hash_chunk(n), 1, ins, in_bytes, out, 32, false,
                       [&](cudaStream_t st, void **d_in, void *d_out, size_t m) { return run(st, d_in[0], d_out, m); });
}

int cuzk_poseidon_hash_pairs(const uint64_t *left, const uint64_t *right, uint64_t *out, size_t n, int mem, void *stream) {
  CUZK_CTX(c);
  if (n == 0) return CUZK_OK;
  if (!left || !right || !out) return fail(CUZK_ERR_INVALID, "null pointer");
  auto run = [](cudaStream_t st, const void *dl, const void *dr, void *dout, size_t m) {
    if (const CoopKind kind = coop_kind(m)) {
      CUZK_COOP_LAUNCH(kind, coop_hash_pairs_kernel, m, st, static_cast<const uint4 *>(dl), static_cast<const uint4 *>(dr),
                       static_cast<uint4 *>(dout), m);
      return check_launch("coop_hash_pairs_kernel");
    }
    hash_pairs_kernel<<<grid_for(m, kBlock), kBlock, 0, st>>>(static_cast<const uint4 *>(dl), static_cast<const uint4 *>(dr),
                                                             static_cast<uint4 *>(dout), m);
    return check_launch("hash_pairs_kernel");
  };
  if (mem == CUZK_MEM_DEVICE) return run(S(stream), left, right, out, n);
  const void *ins[2] = {left, right};
  const size_t in_bytes[2] = {32, 32};
  return host_pipeline(c, n, hash_chunk(n), 2, ins, in_bytes, out, 32, false,
                       [&](cudaStream_t st, void **d_in, void *d_out, size_t m) { return run(st, d_in[0], d_in[1], d_out, m); });
}

int cuzk_poseidon_permutation(uint64_t *states, size_t n, int mem, void *stream) {
  CUZK_CTX(c);
  if (n == 0) return CUZK_OK;
  if (!states) return fail(CUZK_ERR_INVALID, "null pointer");
  auto run = [](cudaStream_t st, void *d, size_t m) {
    if (const CoopKind kind = coop_kind(m)) {
      CUZK_COOP_LAUNCH(kind, coop_permutation_kernel, m, st, static_cast<uint4 *>(d), m);
      return check_launch("coop_permutation_kernel");
    }
    permutation_kernel<<<grid_for(m, kBlock), kBlock, 0, st>>>(static_cast<uint4 *>(d), m);
    return check_launch("permutation_kernel");
  };
  if (mem == CUZK_MEM_DEVICE) return run(S(stream), states, n);
  const void *ins[1] = {states};
  const size_t in_bytes[1] = {96};
  return host_pipeline(c, n, hash_chunk(n), 1, ins, in_bytes, states, 96, true,
                       [&](cudaStream_t st, void **d_in, void *, size_t m) { return run(st, d_in[0], m); });
}

int cuzk_debug_mds_layer(uint64_t *states, size_t n, int mode, void *stream) {
  CUZK_CTX(c);
  (void)c;
  if (n == 0) return CUZK_OK;
  debug_mds_kernel<<<grid_for(n, kBlock), kBlock, 0, S(stream)>>>(reinterpret_cast<uint4 *>(states), n, mode);
  return check_launch("debug_mds_kernel");
}

int cuzk_debug_fast_ops(int op, const uint64_t *a, const uint64_t *b, uint64_t *out, uint32_t *flags, size_t n, void *stream) {
  CUZK_CTX(c);
  (void)c;
  if (op < 0 || op > 3) return fail(CUZK_ERR_INVALID, "unknown fast op");
  if (n == 0) return CUZK_OK;
  if (!a || !out || !flags || (op == 1 && !b)) return fail(CUZK_ERR_INVALID, "null pointer");
  debug_fast_ops_kernel<<<grid_for(n, kBlock), kBlock, 0, S(stream)>>>(op, reinterpret_cast<const uint4 *>(a), reinterpret_cast<const uint4 *>(b),
                                                                      reinterpret_cast<uint4 *>(out), flags, n);
  return check_launch("debug_fast_ops_kernel");
}

int cuzk_poseidon_sponge(const uint64_t *in, size_t width, uint64_t ds, uint64_t *out, size_t n, int mem, void *stream) {
  CUZK_CTX(c);
  if (width > ((size_t)1 << 20)) return fail(CUZK_ERR_INVALID, "sponge width must be <= 2^20");
  if (n == 0) return CUZK_OK;
  if (!out || (width && !in)) return fail(CUZK_ERR_INVALID, "null pointer");
  auto run = [&](cudaStream_t st, const void *din, void *dout, size_t m) {
    if (const CoopKind kind = coop_kind(m)) {
      CUZK_COOP_LAUNCH(kind, coop_sponge_kernel, m, st, static_cast<const uint4 *>(din), (int)width, (u32)ds, (u32)(ds >> 32),
                       static_cast<uint4 *>(dout), m);
      return check_launch("coop_sponge_kernel");
    }
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
  return host_pipeline(c, n, hash_chunk(n), 1, ins, in_bytes, out, 32, false,
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
  CUZK_CTX(c);
  int rc;
  if ((rc = check_arity(arity))) return rc;
  if (height >= (unsigned)kMaxPadLevels) return fail(CUZK_ERR_INVALID, "padding level too high");
  if ((rc = ensure_padding(c, arity, (int)height + 1))) return rc;
  std::lock_guard<std::mutex> lk(g_pad_mu);
  memcpy(out, g_h_pad[arity][height], 32);
  return CUZK_OK;
}

int cuzk_merkle_build(const uint64_t *leaves, size_t n, unsigned arity, uint64_t *levels_out, int mem, void *stream) {
  CUZK_CTX(c);
  int rc;
  if ((rc = check_arity(arity))) return rc;
  if (n == 0) return fail(CUZK_ERR_INVALID, "cuzk_merkle_build: n must be >= 1");
  if (!leaves || !levels_out) return fail(CUZK_ERR_INVALID, "null pointer");
  cudaStream_t st = S(stream);
  if (mem == CUZK_MEM_DEVICE) return merkle_build_dev(c, leaves, n, arity, levels_out, st);
  std::lock_guard<std::mutex> lk(c.hp_mu);
  size_t tot = cuzk_merkle_total_nodes(n, arity);
  void *dl, *dv;
  if ((rc = ws_get(c, 0, n * 32, &dl)) || (rc = ws_get(c, 1, tot * 32, &dv))) return rc;
  if ((rc = bulk_upload(c, dl, leaves, n * 32, st))) return rc;
  rc = merkle_build_dev(c, static_cast<uint64_t *>(dl), n, arity, static_cast<uint64_t *>(dv), st);
  if (rc) return rc;
  return bulk_download(c, levels_out, dv, tot * 32, st);
}

int cuzk_merkle_build_batch(const uint64_t *leaves, size_t n, size_t num_trees, unsigned arity, uint64_t *levels_out, int mem,
                            void *stream) {
  CUZK_CTX(c);
  int rc;
  if ((rc = check_arity(arity))) return rc;
  if (num_trees == 0) return CUZK_OK;
  if (n == 0) return fail(CUZK_ERR_INVALID, "cuzk_merkle_build_batch: n must be >= 1");
  if (!leaves || !levels_out) return fail(CUZK_ERR_INVALID, "null pointer");
  cudaStream_t st = S(stream);
  if (mem == CUZK_MEM_DEVICE) return merkle_build_dev(c, leaves, n, arity, levels_out, st, num_trees);
  std::lock_guard<std::mutex> lk(c.hp_mu);
  const size_t tot = cuzk_merkle_total_nodes(n, arity) * num_trees;
  void *dl, *dv;
  if ((rc = ws_get(c, 0, n * num_trees * 32, &dl)) || (rc = ws_get(c, 1, tot * 32, &dv))) return rc;
  if ((rc = bulk_upload(c, dl, leaves, n * num_trees * 32, st))) return rc;
  rc = merkle_build_dev(c, static_cast<uint64_t *>(dl), n, arity, static_cast<uint64_t *>(dv), st, num_trees);
  if (rc) return rc;
  return bulk_download(c, levels_out, dv, tot * 32, st);
}

int cuzk_merkle_subtree_roots(const uint64_t *leaves, size_t n, unsigned arity, unsigned height, size_t count,
                              uint64_t *roots_out, int mem, void *stream) {
  CUZK_CTX(c);
  int rc;
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
  return subtree_roots_dev(c, leaves, n, arity, height, count, roots_out, S(stream));
}

int cuzk_merkle_top_root(const uint64_t *nodes, size_t count, unsigned arity, uint64_t *root_out, int mem, void *stream) {
  CUZK_CTX(c);
  int rc;
  if ((rc = check_arity(arity))) return rc;
  if (count == 0 || !nodes || !root_out) return fail(CUZK_ERR_INVALID, "bad arguments");
  size_t p = 1;
  unsigned h = 0;
  while (p < count) { p *= arity; ++h; }
  if (p != count) return fail(CUZK_ERR_INVALID, "count must be a power of arity");
  cudaStream_t st = S(stream);
  if (mem == CUZK_MEM_DEVICE) return subtree_roots_dev(c, nodes, count, arity, h, 1, root_out, st);
  std::lock_guard<std::mutex> lk(c.hp_mu);
  void *dn, *dr;
  if ((rc = ws_get(c, 0, count * 32, &dn)) || (rc = ws_get(c, 1, 32, &dr))) return rc;
  CK(cudaMemcpyAsync(dn, nodes, count * 32, cudaMemcpyHostToDevice, st));
  rc = subtree_roots_dev(c, static_cast<uint64_t *>(dn), count, arity, h, 1, static_cast<uint64_t *>(dr), st);
  if (rc) return rc;
  CK(cudaMemcpyAsync(root_out, dr, 32, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return CUZK_OK;
}

int cuzk_merkle_prove_batch(const uint64_t *levels, size_t n, unsigned arity, const uint64_t *indices, size_t num_proofs,
                            uint64_t *siblings_out, uint32_t *positions_out, int mem, void *stream) {
  CUZK_CTX(c);
  int rc;
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
  std::lock_guard<std::mutex> lk(c.hp_mu);
  size_t tot = cuzk_merkle_total_nodes(n, arity);
  void *dl, *di, *ds, *dp;
  if ((rc = ws_get(c, 0, tot * 32, &dl)) || (rc = ws_get(c, 1, num_proofs * 8, &di)) || (rc = ws_get(c, 2, threads * (arity - 1) * 32, &ds)) ||
      (rc = ws_get(c, 3, threads * 4, &dp)))
    return rc;
  CK(cudaMemcpyAsync(dl, levels, tot * 32, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(di, indices, num_proofs * 8, cudaMemcpyHostToDevice, st));
  merkle_prove_kernel<<<grid_for(threads, 256), 256, 0, st>>>(static_cast<uint4 *>(dl), n, padded, (int)arity, (int)nlv,
                                                              static_cast<u64 *>(di), num_proofs, static_cast<uint4 *>(ds),
                                                              static_cast<u32 *>(dp));
  if ((rc = check_launch("merkle_prove_kernel"))) return rc;
  if ((rc = bulk_download(c, siblings_out, ds, threads * (arity - 1) * 32, st))) return rc;
  return bulk_download(c, positions_out, dp, threads * 4, st);
}

// verification of m device-resident proofs on stream st (root from device memory, or by value when root == nullptr)
static int verify_dev(const uint4 *leaves, const uint4 *sib, const u32 *pos, size_t levels, unsigned arity, const uint4 *root, uint4 root_lo,
                      uint4 root_hi, uint8_t *results, size_t m, cudaStream_t st) {
  if (const CoopKind kind = coop_kind(m)) {
    CUZK_COOP_LAUNCH(kind, coop_merkle_verify_kernel, m, st, leaves, sib, pos, (int)levels, (int)arity, root, root_lo, root_hi, results, m);
    return check_launch("coop_merkle_verify_kernel");
  }
  merkle_verify_kernel<<<grid_for(m, kBlock), kBlock, 0, st>>>(leaves, sib, pos, (int)levels, (int)arity, root, root_lo, root_hi, results, m);
  return check_launch("merkle_verify_kernel");
}

int cuzk_merkle_verify_batch(const uint64_t *leaf_values, const uint64_t *siblings, const uint32_t *positions, size_t levels,
                             unsigned arity, const uint64_t *root, uint8_t *results_out, size_t num_proofs, int mem, void *stream) {
  CUZK_CTX(c);
  int rc;
  if ((rc = check_arity(arity))) return rc;
  if (num_proofs == 0) return CUZK_OK;
  if (!leaf_values || !root || !results_out || (levels && (!siblings || !positions))) return fail(CUZK_ERR_INVALID, "null pointer");
  if (mem == CUZK_MEM_DEVICE)
    return verify_dev(reinterpret_cast<const uint4 *>(leaf_values), reinterpret_cast<const uint4 *>(siblings), positions, levels, arity,
                      reinterpret_cast<const uint4 *>(root), uint4{}, uint4{}, results_out, num_proofs, S(stream));
  uint4 root_lo, root_hi;   // host-buffer call: the 32-byte root travels as a kernel argument
  memcpy(&root_lo, root, 16);
  memcpy(&root_hi, root + 2, 16);
  const void *ins[3] = {leaf_values, siblings, positions};
  const size_t in_bytes[3] = {32, levels * (arity - 1) * 32, levels * 4};
  // proofs are independent: chunk them like hashes (each costs levels x ceil(arity/2) permutations)
  return host_pipeline(c, num_proofs, hash_chunk(num_proofs), levels ? 3 : 1, ins, in_bytes, results_out, 1, false,
                       [&](cudaStream_t s2, void **d_in, void *d_out, size_t m) {
                         return verify_dev(static_cast<const uint4 *>(d_in[0]), static_cast<const uint4 *>(d_in[1]),
                                           static_cast<const u32 *>(d_in[2]), levels, arity, nullptr, root_lo, root_hi,
                                           static_cast<uint8_t *>(d_out), m, s2);
                       });
}

// ---- device-resident tree handle ----
// A tree lives on the device it was built on; every call on the handle runs there whatever the caller's current device is.
// Calls on one tree must be issued in order on ONE stream at a time (the handle remembers the last stream used and returns
// its memory to the pool in that stream's order).
struct cuzk_tree {
  uint64_t *levels = nullptr;   // device, level-major, cuzk_merkle_total_nodes elements
  size_t n = 0, padded = 0, total = 0, nlevels = 0;
  unsigned arity = 0;
  int device = 0;
  cudaStream_t stream = nullptr;        // the stream of the last operation on the tree
  unsigned long long *d_oob = nullptr;  // device counter: update indices that were out of range (skipped)
};

int cuzk_tree_build(const uint64_t *leaves, size_t n, unsigned arity, int mem, void *stream, cuzk_tree_t **out) {
  CUZK_CTX(c);
  int rc;
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
  t->device = c.device;
  cudaStream_t st = S(stream);
  t->stream = st;
  // stream-ordered pool allocation: repeated builds reuse the pool's memory without a device-wide synchronisation
  cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&t->levels), t->total * 32 + 8, st);
  if (e != cudaSuccess) { delete t; return cuda_fail(e, "cudaMallocAsync(tree levels)"); }
  t->d_oob = reinterpret_cast<unsigned long long *>(t->levels + t->total * 4);
  e = cudaMemsetAsync(t->d_oob, 0, 8, st);
  if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemsetAsync");
  if (!rc && mem == CUZK_MEM_DEVICE) {
    rc = merkle_build_dev(c, leaves, n, arity, t->levels, st);
  } else if (!rc) {
    std::lock_guard<std::mutex> lk(c.hp_mu);
    void *dl;
    if (!(rc = ws_get(c, 0, n * 32, &dl)) && !(rc = bulk_upload(c, dl, leaves, n * 32, st))) {
      rc = merkle_build_dev(c, static_cast<uint64_t *>(dl), n, arity, t->levels, st);
      if (!rc && (e = cudaStreamSynchronize(st)) != cudaSuccess) rc = cuda_fail(e, "cudaStreamSynchronize");
    }
  }
  if (rc) {
    cudaFreeAsync(t->levels, st);
    delete t;
    return rc;
  }
  *out = t;
  return CUZK_OK;
}

int cuzk_tree_free(cuzk_tree_t *t) {
  if (!t) return CUZK_OK;
  CtxGuard guard(t->device);
  cudaError_t e = cudaFreeAsync(t->levels, t->stream);   // ordered after the work already enqueued on the tree's stream
  delete t;
  if (e != cudaSuccess) return cuda_fail(e, "cudaFreeAsync(tree levels)");
  return CUZK_OK;
}

size_t cuzk_tree_leaf_count(const cuzk_tree_t *t) { return t ? t->n : 0; }
size_t cuzk_tree_num_levels(const cuzk_tree_t *t) { return t ? t->nlevels : 0; }
size_t cuzk_tree_total_nodes(const cuzk_tree_t *t) { return t ? t->total : 0; }
unsigned cuzk_tree_arity(const cuzk_tree_t *t) { return t ? t->arity : 0; }
int cuzk_tree_device(const cuzk_tree_t *t) { return t ? t->device : -1; }
const uint64_t *cuzk_tree_device_levels(const cuzk_tree_t *t) { return t ? t->levels : nullptr; }

int cuzk_tree_root(const cuzk_tree_t *t, uint64_t *root_out, int mem, void *stream) {
  if (!t || !root_out) return fail(CUZK_ERR_INVALID, "null pointer");
  CUZK_CTX_ON(c, t->device);
  (void)c;
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
  CUZK_CTX_ON(c, t->device);
  cudaStream_t st = S(stream);
  if (mem == CUZK_MEM_DEVICE) {
    CK(cudaMemcpyAsync(levels_out, t->levels, t->total * 32, cudaMemcpyDeviceToDevice, st));
    return CUZK_OK;
  }
  std::lock_guard<std::mutex> lk(c.hp_mu);
  return bulk_download(c, levels_out, t->levels, t->total * 32, st);
}

int cuzk_tree_level(const cuzk_tree_t *t, size_t level, uint64_t *level_out, int mem, void *stream) {
  if (!t || !level_out) return fail(CUZK_ERR_INVALID, "null pointer");
  if (level >= t->nlevels) return fail(CUZK_ERR_INVALID, "cuzk_tree_level: no such level");
  CUZK_CTX_ON(c, t->device);
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
  std::lock_guard<std::mutex> lk(c.hp_mu);
  return bulk_download(c, level_out, t->levels + 4 * off, width * 32, st);
}

int cuzk_tree_prove_batch(const cuzk_tree_t *t, const uint64_t *indices, size_t num_proofs, uint64_t *siblings_out,
                          uint32_t *positions_out, int mem, void *stream) {
  if (!t) return fail(CUZK_ERR_INVALID, "null tree");
  CUZK_CTX_ON(c, t->device);
  int rc;
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
  std::lock_guard<std::mutex> lk(c.hp_mu);
  void *di, *ds, *dp;
  if ((rc = ws_get(c, 1, num_proofs * 8, &di)) || (rc = ws_get(c, 2, sib_bytes, &ds)) || (rc = ws_get(c, 3, threads * 4, &dp))) return rc;
  CK(cudaMemcpyAsync(di, indices, num_proofs * 8, cudaMemcpyHostToDevice, st));
  if ((rc = launch(static_cast<u64 *>(di), static_cast<uint4 *>(ds), static_cast<u32 *>(dp)))) return rc;
  if ((rc = bulk_download(c, siblings_out, ds, sib_bytes, st))) return rc;
  return bulk_download(c, positions_out, dp, threads * 4, st);
}

int cuzk_tree_verify_batch(const cuzk_tree_t *t, const uint64_t *leaf_values, const uint64_t *siblings, const uint32_t *positions,
                           uint8_t *results_out, size_t num_proofs, int mem, void *stream) {
  if (!t) return fail(CUZK_ERR_INVALID, "null tree");
  CtxGuard guard(t->device);   // the nested calls below then find the tree's device current
  if (!guard.get()) return fail(CUZK_ERR_CUDA, CUZK_NOT_INIT_MSG);
  const size_t nlv = t->nlevels - 1;
  if (mem == CUZK_MEM_DEVICE)
    return cuzk_merkle_verify_batch(leaf_values, siblings, positions, nlv, t->arity, t->levels + (t->total - 1) * 4, results_out, num_proofs,
                                    mem, stream);
  uint64_t root[4];
  int rc = cuzk_tree_root(t, root, CUZK_MEM_HOST, stream);
  if (rc) return rc;
  return cuzk_merkle_verify_batch(leaf_values, siblings, positions, nlv, t->arity, root, results_out, num_proofs, mem, stream);
}

// NaryMerkleTree::update_leaf (merkle_tree.cpp:294-301; a full rebuild per leaf in the reference) for a batch, by path re-hash.
// The result equals a serial loop of update_leaf calls: for an index that occurs more than once the LAST value wins.
// Indices >= leaf count: host-buffer calls are refused as a whole (nothing written; the reference throws std::out_of_range,
// :296-298); device-pointer calls are asynchronous, skip such entries and count them (cuzk_tree_oob_count).
int cuzk_tree_update_leaves(cuzk_tree_t *t, const uint64_t *indices, const uint64_t *values, size_t count, int mem, void *stream) {
  if (!t) return fail(CUZK_ERR_INVALID, "null tree");
  CUZK_CTX_ON(c, t->device);
  int rc;
  if (count == 0) return CUZK_OK;
  if (!indices || !values) return fail(CUZK_ERR_INVALID, "null pointer");
  if (count > 0xffffffffull) return fail(CUZK_ERR_INVALID, "cuzk_tree_update_leaves: at most 2^32 - 1 updates per call");
  cudaStream_t st = S(stream);
  t->stream = st;
  const u64 *di = indices;
  const uint4 *dv = reinterpret_cast<const uint4 *>(values);
  std::unique_lock<std::mutex> lk(c.hp_mu, std::defer_lock);
  if (mem != CUZK_MEM_DEVICE) {
    for (size_t q = 0; q < count; ++q)
      if (indices[q] >= t->n) return fail(CUZK_ERR_INVALID, "cuzk_tree_update_leaves: leaf index out of range");
    lk.lock();
    void *wi, *wv;
    if ((rc = ws_get(c, 1, count * 8, &wi)) || (rc = ws_get(c, 2, count * 32, &wv))) return rc;
    CK(cudaMemcpyAsync(wi, indices, count * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(wv, values, count * 32, cudaMemcpyHostToDevice, st));
    di = static_cast<const u64 *>(wi);
    dv = static_cast<const uint4 *>(wv);
  }
  uint4 *cur = reinterpret_cast<uint4 *>(t->levels);
  if (count == 1) {
    merkle_write_leaves_kernel<<<1, 32, 0, st>>>(cur, di, nullptr, dv, 1, t->n, t->d_oob);
    if ((rc = check_launch("merkle_write_leaves_kernel"))) return rc;
  } else {
    // stable sort of (index, batch position) so that the last writer of every leaf is known
    u64 *keys = nullptr;
    u32 *pos_in = nullptr, *pos_out = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, di, keys, pos_in, pos_out, (int)count, 0, 64, st);
    void *block = nullptr;
    const size_t keys_off = 0, pin_off = count * 8, pout_off = pin_off + ((count * 4 + 7) & ~(size_t)7),
                 tmp_off = pout_off + ((count * 4 + 255) & ~(size_t)255);
    cudaError_t e = cudaMallocAsync(&block, tmp_off + tmp_bytes + 256, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync(update scratch)");
    keys = reinterpret_cast<u64 *>(static_cast<char *>(block) + keys_off);
    pos_in = reinterpret_cast<u32 *>(static_cast<char *>(block) + pin_off);
    pos_out = reinterpret_cast<u32 *>(static_cast<char *>(block) + pout_off);
    tmp = static_cast<char *>(block) + tmp_off;
    iota_u32_kernel<<<grid_for(count, 256), 256, 0, st>>>(pos_in, count);
    rc = check_launch("iota_u32_kernel");
    if (!rc) {
      e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, di, keys, pos_in, pos_out, (int)count, 0, 64, st);
      if (e != cudaSuccess) rc = cuda_fail(e, "cub::DeviceRadixSort::SortPairs");
    }
    if (!rc) {
      merkle_write_leaves_kernel<<<grid_for(count, 256), 256, 0, st>>>(cur, keys, pos_out, dv, count, t->n, t->d_oob);
      rc = check_launch("merkle_write_leaves_kernel");
    }
    cudaFreeAsync(block, st);
    if (rc) return rc;
  }
  size_t p = t->padded;
  u64 divisor = 1;
  while (p > 1) {
    divisor *= t->arity;
    if (const CoopKind kind = coop_kind(count))
      CUZK_COOP_LAUNCH(kind, coop_merkle_update_level_kernel, count, st, cur, cur + 2 * p, di, count, t->n, divisor, (int)t->arity);
    else
      merkle_update_level_kernel<<<grid_for(count, kBlock), kBlock, 0, st>>>(cur, cur + 2 * p, di, count, t->n, divisor, (int)t->arity);
    if ((rc = check_launch("merkle_update_level_kernel"))) return rc;
    cur += 2 * p;
    p /= t->arity;
  }
  if (mem != CUZK_MEM_DEVICE) CK(cudaStreamSynchronize(st));
  return CUZK_OK;
}

int cuzk_tree_oob_count(const cuzk_tree_t *t, uint64_t *count_out) {
  if (!t || !count_out) return fail(CUZK_ERR_INVALID, "null pointer");
  CUZK_CTX_ON(c, t->device);
  (void)c;
  unsigned long long v = 0;
  CK(cudaMemcpyAsync(&v, t->d_oob, 8, cudaMemcpyDeviceToHost, t->stream));
  CK(cudaStreamSynchronize(t->stream));
  *count_out = v;
  return CUZK_OK;
}

// leaf indices n .. n+count-1 for the append path
__global__ void iota_u64_kernel(u64 *out, u64 first, size_t count) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = first + i;
}

int cuzk_tree_append_leaves(cuzk_tree_t *t, const uint64_t *values, size_t count, int mem, void *stream) {
  if (!t) return fail(CUZK_ERR_INVALID, "null tree");
  CUZK_CTX_ON(c, t->device);
  int rc = CUZK_OK;
  if (count == 0) return CUZK_OK;
  if (!values) return fail(CUZK_ERR_INVALID, "null pointer");
  cudaStream_t st = S(stream);
  t->stream = st;
  // bring the new values to the device
  void *dv = nullptr;
  cudaError_t e = cudaMallocAsync(&dv, count * 32, st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync(append values)");
  e = cudaMemcpyAsync(dv, values, count * 32, mem == CUZK_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) { cudaFreeAsync(dv, st); return cuda_fail(e, "cudaMemcpyAsync(append values)"); }
  if (t->n + count <= t->padded) {
    // the padded shape does not change: the new leaves replace padding slots and only their ancestors are re-hashed
    void *di = nullptr;
    if ((e = cudaMallocAsync(&di, count * 8, st)) != cudaSuccess) { cudaFreeAsync(dv, st); return cuda_fail(e, "cudaMallocAsync(indices)"); }
    iota_u64_kernel<<<grid_for(count, 256), 256, 0, st>>>(static_cast<u64 *>(di), t->n, count);
    rc = check_launch("iota_u64_kernel");
    const size_t old_n = t->n;
    t->n += count;   // update_leaves checks indices against the new count
    if (!rc) rc = cuzk_tree_update_leaves(t, static_cast<const uint64_t *>(di), static_cast<const uint64_t *>(dv), count, CUZK_MEM_DEVICE, stream);
    if (rc) t->n = old_n;
    cudaFreeAsync(di, st);
    cudaFreeAsync(dv, st);
    if (!rc && mem != CUZK_MEM_DEVICE) CK(cudaStreamSynchronize(st));
    return rc;
  }
  // the tree outgrows its padded leaf level: build the larger tree from the old leaves (still in level 0) plus the new ones
  const size_t n2 = t->n + count;
  void *all = nullptr;
  if ((e = cudaMallocAsync(&all, n2 * 32, st)) != cudaSuccess) { cudaFreeAsync(dv, st); return cuda_fail(e, "cudaMallocAsync(leaves)"); }
  if ((e = cudaMemcpyAsync(all, t->levels, t->n * 32, cudaMemcpyDeviceToDevice, st)) != cudaSuccess ||
      (e = cudaMemcpyAsync(static_cast<char *>(all) + t->n * 32, dv, count * 32, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) {
    cudaFreeAsync(all, st);
    cudaFreeAsync(dv, st);
    return cuda_fail(e, "cudaMemcpyAsync(leaves)");
  }
  const size_t total2 = cuzk_merkle_total_nodes(n2, t->arity);
  uint64_t *levels2 = nullptr;
  if ((e = cudaMallocAsync(reinterpret_cast<void **>(&levels2), total2 * 32 + 8, st)) != cudaSuccess) {
    cudaFreeAsync(all, st);
    cudaFreeAsync(dv, st);
    return cuda_fail(e, "cudaMallocAsync(tree levels)");
  }
  rc = merkle_build_dev(c, static_cast<const uint64_t *>(all), n2, t->arity, levels2, st);
  if (!rc && (e = cudaMemcpyAsync(levels2 + total2 * 4, t->d_oob, 8, cudaMemcpyDeviceToDevice, st)) != cudaSuccess)
    rc = cuda_fail(e, "cudaMemcpyAsync(counter)");
  cudaFreeAsync(all, st);
  cudaFreeAsync(dv, st);
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess && !rc) rc = cuda_fail(e, "cudaStreamSynchronize");
  if (rc) {
    cudaFreeAsync(levels2, st);
    return rc;
  }
  cudaFreeAsync(t->levels, st);
  t->levels = levels2;
  t->d_oob = reinterpret_cast<unsigned long long *>(levels2 + total2 * 4);
  t->n = n2;
  t->padded = cuzk_merkle_padded_leaves(n2, t->arity);
  t->total = total2;
  t->nlevels = cuzk_merkle_num_levels(n2, t->arity);
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

#include "multi_gpu.cuh"
