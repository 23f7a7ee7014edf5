// host_path.cuh -- host-side machinery behind the C ABI (included once, by cuzk_kernels.cu): staging buffers and the chunked
// double-buffered pipeline for host-buffer calls, the copy pool for pageable memory, padding constants, Merkle level
// scheduling (per-level launches, grouped subtree passes on internal streams).
#pragma once

// ------------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------------
namespace {

// Host-buffer calls (mem == CUZK_MEM_HOST) stage through library-owned device buffers that are kept between calls
// (the reference mallocs and frees on every call, poseidon_cuda.cu:374-408) and are cut into chunks that alternate
// over kPipeStreams internal streams, so the H2D copies of the next chunks and the D2H copies of the previous ones overlap
// the kernel of chunk c.
#ifndef CUZK_PIPE_STREAMS
#define CUZK_PIPE_STREAMS 4   // measured: 2 streams x 1 wave 171 M/s, 3 x 1 178, 4 x 1/2 181, 8 x 1/4 182 (pinned, 1 M pairs)
#endif
constexpr int kPipeStreams = CUZK_PIPE_STREAMS;
constexpr int kPipeSlots = 4;                       // up to 3 inputs + 1 output per stream
constexpr size_t kHashChunk = 148 * CUZK_MIN_BLOCKS * CUZK_BLOCK / 2;   // half a resident wave of one-thread-per-hash CTAs per chunk
constexpr size_t kCheapChunk = 1 << 20;             // element-wise field ops
constexpr int kWsSlots = 6;

struct HostPath {
  cudaStream_t stream[kPipeStreams] = {};
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

// ---- cooperative (sixteen lanes per unit) dispatch ----------------------------------------------------------------------
// A launch of the one-thread-per-unit kernels takes one permutation latency (~186 us) however few units it has; the
// cooperative kernels (coop_kernels.cuh) take a fraction of that while they fit the chip at about one warp per SM
// sub-partition (148 x 4 warps x 2 units) and lose to the one-thread kernels once their larger instruction stream per unit
// fills the issue slots.  Launchers switch at g_coop_max units (cuzk_debug_set_coop_max; 0 = never).
#ifndef CUZK_COOP_MAX_DEFAULT
#define CUZK_COOP_MAX_DEFAULT 2368
#endif
size_t g_coop_max = CUZK_COOP_MAX_DEFAULT;
inline bool use_coop(size_t units) { return units != 0 && units <= g_coop_max; }
inline unsigned coop_grid(size_t units) { return grid_for(units * 16, kCoopBlock); }

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
  if (const char *env = getenv("CUZK_CHUNK_PERCENT")) {   // tuning knob: chunk size in percent of the default
    const long pct = atol(env);
    if (pct > 0) chunk = std::max<size_t>(1, chunk * (size_t)pct / 100);
  }
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
  if (g_coop_max != 0) coop_padding_chain_kernel<<<1, 32>>>(reinterpret_cast<uint4 *>(d), (int)arity, start, end);
  else padding_chain_kernel<<<1, 32>>>(reinterpret_cast<uint4 *>(d), (int)arity, start, end);
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
  if (use_coop(out_count * ntrees)) {
    coop_merkle_level_kernel<<<coop_grid(out_count * ntrees), kCoopBlock, 0, st>>>(in, out, in_real, out_count, (int)arity, pad_in, pad_in + 2,
                                                                               ntrees, tree_stride);
    return check_launch("coop_merkle_level_kernel");
  }
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
