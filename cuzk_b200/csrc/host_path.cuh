// host_path.cuh -- host-side machinery behind the C ABI (included once, by cuzk_kernels.cu): the per-device context, staging
// buffers and the chunked multi-stream pipeline for host-buffer calls, the copy pool for pageable memory, padding constants,
// Merkle level scheduling (per-level launches, groups of subtrees on internal streams, cooperative kernels for narrow levels).
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
constexpr int kSubtreeStreams = 16;   // internal streams per device; g_build_streams of them are used

struct PinnedStage {   // per stream: pinned twins of the device staging buffers + "results are in the bounce buffer" event
  void *buf[kPipeSlots] = {};
  size_t cap[kPipeSlots] = {};
  cudaEvent_t done = nullptr;
};

// Everything the library keeps on ONE device.  A process may initialise several devices (cuzk_init(d) for each); every
// entry point runs on the calling thread's current CUDA device and uses that device's context (see CtxGuard).
struct Ctx {
  int device = -1;
  int refcount = 0;
  int sm_count = 148;
  // padding constants E_l per arity on this device (the host copy g_h_pad is shared by all devices)
  uint64_t *d_pad[9] = {nullptr};
  int pad_levels[9] = {0};
  std::mutex pad_mu;
  // staging for host-buffer calls
  cudaStream_t stream[kPipeStreams] = {};
  void *buf[kPipeStreams][kPipeSlots] = {};
  size_t cap[kPipeStreams][kPipeSlots] = {};
  void *ws[kWsSlots] = {};
  size_t ws_cap[kWsSlots] = {};
  bool hp_ready = false;
  PinnedStage pin[kPipeStreams];
  std::mutex hp_mu;   // host-buffer calls on one device serialise on its staging buffers
  // internal streams for groups of subtrees
  cudaStream_t sub_stream[kSubtreeStreams] = {};
  cudaEvent_t sub_fork = nullptr, sub_join[kSubtreeStreams] = {};
  std::mutex sub_mu;
};
constexpr int kMaxDevices = 32;
Ctx g_ctx[kMaxDevices];

// The context of the calling thread's current device.  When that device was never initialised but exactly one other device
// was (the common single-GPU program whose worker threads never call cudaSetDevice), the call switches to it and restores
// the caller's device on return: kernels, constants and staging buffers always belong to one and the same device.
class CtxGuard {
 public:
  CtxGuard() {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); return; }
    if (cur >= 0 && cur < kMaxDevices && g_ctx[cur].refcount > 0) { ctx_ = &g_ctx[cur]; return; }
    int only = -1, count = 0;
    for (int d = 0; d < kMaxDevices; ++d)
      if (g_ctx[d].refcount > 0) { only = d; ++count; }
    if (count == 1 && cudaSetDevice(only) == cudaSuccess) {
      restore_ = cur;
      ctx_ = &g_ctx[only];
    }
  }
  explicit CtxGuard(int device) {   // run on `device` whatever the caller's current device is (tree handles, multi-GPU layer)
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); return; }
    if (device < 0 || device >= kMaxDevices || g_ctx[device].refcount <= 0) return;
    if (cur != device) {
      if (cudaSetDevice(device) != cudaSuccess) return;
      restore_ = cur;
    }
    ctx_ = &g_ctx[device];
  }
  ~CtxGuard() {
    if (restore_ >= 0) cudaSetDevice(restore_);
  }
  CtxGuard(const CtxGuard &) = delete;
  CtxGuard &operator=(const CtxGuard &) = delete;
  Ctx *get() const { return ctx_; }

 private:
  Ctx *ctx_ = nullptr;
  int restore_ = -1;
};
#define CUZK_NOT_INIT_MSG "cuzk_b200: library not initialised on the calling thread's CUDA device (call cuzk_init)"
#define CUZK_CTX(c)                                          \
  CtxGuard guard__;                                          \
  if (!guard__.get()) return fail(CUZK_ERR_CUDA, CUZK_NOT_INIT_MSG); \
  Ctx &c = *guard__.get()
#define CUZK_CTX_ON(c, device)                               \
  CtxGuard guard__(device);                                  \
  if (!guard__.get()) return fail(CUZK_ERR_CUDA, CUZK_NOT_INIT_MSG); \
  Ctx &c = *guard__.get()

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
int ws_get(Ctx &c, int slot, size_t bytes, void **out) {
  int rc = hp_reserve(c.ws[slot], c.ws_cap[slot], bytes ? bytes : 1);
  *out = c.ws[slot];
  return rc;
}
int hp_start(Ctx &c) {
  if (c.hp_ready) return CUZK_OK;
  for (int i = 0; i < kPipeStreams; ++i) CK(cudaStreamCreateWithFlags(&c.stream[i], cudaStreamNonBlocking));
  c.hp_ready = true;
  return CUZK_OK;
}
void hp_stop(Ctx &c) {
  for (int i = 0; i < kPipeStreams; ++i) {
    for (int j = 0; j < kPipeSlots; ++j) {
      if (c.buf[i][j]) cudaFree(c.buf[i][j]);
      c.buf[i][j] = nullptr;
      c.cap[i][j] = 0;
      if (c.pin[i].buf[j]) cudaFreeHost(c.pin[i].buf[j]);
      c.pin[i].buf[j] = nullptr;
      c.pin[i].cap[j] = 0;
    }
    if (c.pin[i].done) cudaEventDestroy(c.pin[i].done);
    c.pin[i].done = nullptr;
    if (c.stream[i]) cudaStreamDestroy(c.stream[i]);
    c.stream[i] = nullptr;
  }
  for (int j = 0; j < kWsSlots; ++j) {
    if (c.ws[j]) cudaFree(c.ws[j]);
    c.ws[j] = nullptr;
    c.ws_cap[j] = 0;
  }
  c.hp_ready = false;
}

int check_launch(const char *what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, what);
  return CUZK_OK;
}

// ---- cooperative (a group of lanes per unit) dispatch ------------------------------------------------------------------
// A launch of the one-thread-per-unit kernels takes one permutation latency (~186 us) however few units it has.  The
// cooperative kernels (coop_kernels.cuh) take a fraction of that while the chip is not full:
//   Wide16  (16 lanes per unit, 2 units per warp): the lowest latency; measured 93 us up to 1184 units (one warp per SM
//           sub-partition), 115 us up to 2368, 148 us at 3552, 189 us at 4736
//   Narrow8 (8 lanes per unit, 4 units per warp): fewer instructions per unit; measured 138-141 us up to 2368 units, 166 us
//           up to 4736, 215 us at 5920 -- the better choice between 2368 and 4736 units, where the one-thread kernels'
//           186 us win again
// Thresholds: g_coop_wide_max and g_coop_max units (cuzk_debug_set_coop_wide_max / cuzk_debug_set_coop_max; coop_max 0 =
// never cooperative).
#ifndef CUZK_COOP_WIDE_MAX_DEFAULT
#define CUZK_COOP_WIDE_MAX_DEFAULT 2368
#endif
#ifndef CUZK_COOP_MAX_DEFAULT
#define CUZK_COOP_MAX_DEFAULT 4736
#endif
// Full-tree builds: a large tree is cut into g_build_groups groups of subtrees dealt over g_build_streams internal streams;
// inside the groups only levels of at most g_group_coop_max nodes take the cooperative kernels (cuzk_debug_set_build_plan).
// Measured (profiles/r02_build_plans.jsonl; 2^20 leaves arity 4 / 2^23 leaves arity 8 / 2^20 leaves arity 2, ms): one stream
// 6.16 / 30.1 / 8.30; 4 groups 6.44 / 30.6 / 8.05; 8 groups, cap 1184: 5.82 / 29.7 / 7.85; 16 groups 7.3 / 31.9 / 9.5.
#ifndef CUZK_BUILD_GROUPS_DEFAULT
#define CUZK_BUILD_GROUPS_DEFAULT 8
#endif
#ifndef CUZK_BUILD_STREAMS_DEFAULT
#define CUZK_BUILD_STREAMS_DEFAULT 8
#endif
#ifndef CUZK_GROUP_COOP_MAX_DEFAULT
#define CUZK_GROUP_COOP_MAX_DEFAULT 1184
#endif
std::atomic<int> g_build_groups{CUZK_BUILD_GROUPS_DEFAULT};
std::atomic<int> g_build_streams{CUZK_BUILD_STREAMS_DEFAULT};
std::atomic<size_t> g_group_coop_max{CUZK_GROUP_COOP_MAX_DEFAULT};
std::atomic<size_t> g_coop_wide_max{CUZK_COOP_WIDE_MAX_DEFAULT};
std::atomic<size_t> g_coop_max{CUZK_COOP_MAX_DEFAULT};
enum CoopKind { kOneThread = 0, kWide16 = 1, kNarrow8 = 2 };
inline CoopKind coop_kind(size_t units) {
  if (units == 0 || units > g_coop_max.load(std::memory_order_relaxed)) return kOneThread;
  return units <= g_coop_wide_max.load(std::memory_order_relaxed) ? kWide16 : kNarrow8;
}
inline bool use_coop(size_t units) { return coop_kind(units) != kOneThread; }
inline unsigned coop_grid(size_t units, int lanes) { return grid_for(units * (size_t)lanes, kCoopBlock); }
// launches kernel<Wide16> or kernel<Narrow8> over `units` units
#define CUZK_COOP_LAUNCH(kind, kernel, units, st, ...)                                              \
  do {                                                                                              \
    if ((kind) == kWide16) kernel<Wide16><<<coop_grid(units, 16), kCoopBlock, 0, st>>>(__VA_ARGS__); \
    else kernel<Narrow8><<<coop_grid(units, 8), kCoopBlock, 0, st>>>(__VA_ARGS__);                   \
  } while (0)

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
    std::lock_guard<std::mutex> one_at_a_time(call_mu_);   // the contexts of several devices share the pool
    const int parts = (int)workers_.size() + 1;
    const size_t slice = (((bytes + parts - 1) / parts) + 4095) & ~(size_t)4095;   // parts * slice >= bytes
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
    std::lock_guard<std::mutex> lk(call_mu_);
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
  std::mutex mu_, call_mu_;
  std::condition_variable cv_, done_cv_;
  char *dst_ = nullptr;
  const char *src_ = nullptr;
  size_t bytes_ = 0, slice_ = 0;
  int pending_ = 0;
  uint64_t generation_ = 0;
  bool stop_ = false, started_ = false;
};
CopyPool g_copy_pool;

int pin_reserve(void *&p, size_t &cap, size_t bytes) {
  if (bytes <= cap) return CUZK_OK;
  if (p) CK(cudaFreeHost(p));
  p = nullptr;
  cap = 0;
  CK(cudaHostAlloc(&p, bytes, cudaHostAllocPortable));
  cap = bytes;
  return CUZK_OK;
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
// that `host` may be reused (upload) / read (download) by the caller.  Call with c.hp_mu held.
constexpr size_t kBulkPiece = (size_t)8 << 20;
int bulk_upload(Ctx &c, void *dev, const void *host, size_t bytes, cudaStream_t st) {
  if (bytes < ((size_t)4 << 20) || !is_pageable(host)) {
    CK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, st));
    return CUZK_OK;
  }
  int rc;
  for (int s = 0; s < kPipeStreams; ++s) {
    if ((rc = pin_reserve(c.pin[s].buf[0], c.pin[s].cap[0], kBulkPiece))) return rc;
    if (!c.pin[s].done) CK(cudaEventCreateWithFlags(&c.pin[s].done, cudaEventDisableTiming));
  }
  bool used[kPipeStreams] = {};
  size_t at = 0;
  for (int k = 0; at < bytes; ++k) {
    const int s = k % kPipeStreams;
    const size_t m = std::min(kBulkPiece, bytes - at);
    if (used[s]) CK(cudaEventSynchronize(c.pin[s].done));   // the piece that used this bounce buffer has left it
    g_copy_pool.copy(c.pin[s].buf[0], static_cast<const char *>(host) + at, m);
    CK(cudaMemcpyAsync(static_cast<char *>(dev) + at, c.pin[s].buf[0], m, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(c.pin[s].done, st));
    used[s] = true;
    at += m;
  }
  for (int s = 0; s < kPipeStreams; ++s)
    if (used[s]) CK(cudaEventSynchronize(c.pin[s].done));
  return CUZK_OK;
}
int bulk_download(Ctx &c, void *host, const void *dev, size_t bytes, cudaStream_t st) {
  if (bytes < ((size_t)4 << 20) || !is_pageable(host)) {
    CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return CUZK_OK;
  }
  int rc;
  for (int s = 0; s < kPipeStreams; ++s) {
    if ((rc = pin_reserve(c.pin[s].buf[0], c.pin[s].cap[0], kBulkPiece))) return rc;
    if (!c.pin[s].done) CK(cudaEventCreateWithFlags(&c.pin[s].done, cudaEventDisableTiming));
  }
  size_t pending_at[kPipeStreams] = {}, pending_m[kPipeStreams] = {};
  auto drain = [&](int s) -> int {
    if (!pending_m[s]) return CUZK_OK;
    CK(cudaEventSynchronize(c.pin[s].done));
    g_copy_pool.copy(static_cast<char *>(host) + pending_at[s], c.pin[s].buf[0], pending_m[s]);
    pending_m[s] = 0;
    return CUZK_OK;
  };
  size_t at = 0;
  for (int k = 0; at < bytes; ++k) {
    const int s = k % kPipeStreams;
    const size_t m = std::min(kBulkPiece, bytes - at);
    if ((rc = drain(s))) return rc;
    CK(cudaMemcpyAsync(c.pin[s].buf[0], static_cast<const char *>(dev) + at, m, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(c.pin[s].done, st));
    pending_at[s] = at;
    pending_m[s] = m;
    at += m;
  }
  for (int s = 0; s < kPipeStreams; ++s)
    if ((rc = drain(s))) return rc;
  return CUZK_OK;
}

// Small host-buffer calls skip the copy engines: the kernel reads its inputs from, and writes its results to, pinned host
// memory through PCIe (the caller's own buffers when they are pinned, pinned bounce copies of pageable ones).  A 4096-pair
// call moves 384 KiB -- a few microseconds of PCIe -- while two uploads, a download and the engine hand-overs between them
// cost ~60 us on top of a ~170 us kernel (profiles/r02_tuning_notes.md).  cuzk_debug_set_direct_max(0) turns it off.
#ifndef CUZK_DIRECT_MAX_DEFAULT
#define CUZK_DIRECT_MAX_DEFAULT ((size_t)1 << 20)
#endif
constexpr size_t kDirectBounceMin = (size_t)1 << 20;   // bounce buffers are not re-allocated call by call as batches grow
std::atomic<size_t> g_direct_max{CUZK_DIRECT_MAX_DEFAULT};   // bytes (inputs + outputs) up to which a call goes direct
// the device's address of pinned / registered host memory, nullptr for anything else
inline void *device_alias(const void *host) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, host) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return attr.type == cudaMemoryTypeHost ? attr.devicePointer : nullptr;
}

// Chunked, multi-stream host->device->host pass.  `nin` input arrays of `in_bytes[k]` bytes per unit, one output
// array of `out_bytes` per unit (out may alias in[0] for in-place ops).  launch(stream, d_in[], d_out, m) enqueues the
// kernel(s) for m units.  Returns after every result byte is in `out`.
template <class Launch>
int host_pipeline(Ctx &c, size_t n, size_t chunk, int nin, const void *const *in, const size_t *in_bytes, void *out, size_t out_bytes,
                  bool out_aliases_in0, Launch launch) {
  std::lock_guard<std::mutex> lk(c.hp_mu);
  int rc = hp_start(c);
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
  if (n * unit_bytes <= g_direct_max.load(std::memory_order_relaxed)) {
    void *d_in[kPipeSlots] = {};
    bool bounced0 = false;
    for (int k = 0; k < nin; ++k) {
      d_in[k] = device_alias(in[k]);
      if (d_in[k] == nullptr) {
        if ((rc = pin_reserve(c.pin[0].buf[k], c.pin[0].cap[k], std::max(n * in_bytes[k], kDirectBounceMin)))) return rc;
        memcpy(c.pin[0].buf[k], in[k], n * in_bytes[k]);
        d_in[k] = device_alias(c.pin[0].buf[k]);
        if (k == 0) bounced0 = true;
      }
    }
    void *d_out = out_aliases_in0 ? d_in[0] : device_alias(out);
    const bool bounced_out = out_aliases_in0 ? bounced0 : d_out == nullptr;
    if (!out_aliases_in0 && bounced_out) {
      if ((rc = pin_reserve(c.pin[0].buf[kPipeSlots - 1], c.pin[0].cap[kPipeSlots - 1], std::max(n * out_bytes, kDirectBounceMin)))) return rc;
      d_out = device_alias(c.pin[0].buf[kPipeSlots - 1]);
    }
    bool mapped = d_out != nullptr;
    for (int k = 0; k < nin; ++k) mapped = mapped && d_in[k] != nullptr;
    if (mapped) {
      if ((rc = launch(c.stream[0], d_in, d_out, n))) return rc;
      CK(cudaStreamSynchronize(c.stream[0]));
      if (bounced_out) memcpy(out, out_aliases_in0 ? c.pin[0].buf[0] : c.pin[0].buf[kPipeSlots - 1], n * out_bytes);
      return CUZK_OK;
    }
    // pinned memory this device cannot address: the copy path below
  }
  for (int s = 0; s < kPipeStreams; ++s) {
    for (int k = 0; k < nin; ++k)
      if ((rc = hp_reserve(c.buf[s][k], c.cap[s][k], chunk * in_bytes[k]))) return rc;
    if (!out_aliases_in0 && (rc = hp_reserve(c.buf[s][kPipeSlots - 1], c.cap[s][kPipeSlots - 1], chunk * out_bytes))) return rc;
    if (staged) {
      for (int k = 0; k < nin; ++k)
        if ((rc = pin_reserve(c.pin[s].buf[k], c.pin[s].cap[k], chunk * in_bytes[k]))) return rc;
      if ((rc = pin_reserve(c.pin[s].buf[kPipeSlots - 1], c.pin[s].cap[kPipeSlots - 1], chunk * out_bytes))) return rc;
      if (!c.pin[s].done) CK(cudaEventCreateWithFlags(&c.pin[s].done, cudaEventDisableTiming));
    }
  }
  size_t done = 0;
  size_t drain_at[kPipeStreams] = {}, drain_m[kPipeStreams] = {};   // staged mode: the chunk whose results sit in each bounce buffer
  auto drain = [&](int s) -> int {
    if (drain_m[s] == 0) return CUZK_OK;
    CK(cudaEventSynchronize(c.pin[s].done));
    g_copy_pool.copy(static_cast<char *>(out) + drain_at[s] * out_bytes, c.pin[s].buf[kPipeSlots - 1], drain_m[s] * out_bytes);
    drain_m[s] = 0;
    return CUZK_OK;
  };
  for (int k2 = 0; done < n; ++k2) {
    const int s = k2 % kPipeStreams;
    const size_t m = (n - done < chunk) ? n - done : chunk;
    cudaStream_t st = c.stream[s];
    void *d_in[kPipeSlots] = {};
    if (staged && (rc = drain(s))) return rc;   // the slot's previous results leave before its buffers are reused
    for (int k = 0; k < nin; ++k) {
      d_in[k] = c.buf[s][k];
      const char *src = static_cast<const char *>(in[k]) + done * in_bytes[k];
      if (staged) {
        g_copy_pool.copy(c.pin[s].buf[k], src, m * in_bytes[k]);
        src = static_cast<const char *>(c.pin[s].buf[k]);
      }
      CK(cudaMemcpyAsync(d_in[k], src, m * in_bytes[k], cudaMemcpyHostToDevice, st));
    }
    void *d_out = out_aliases_in0 ? d_in[0] : c.buf[s][kPipeSlots - 1];
    if ((rc = launch(st, d_in, d_out, m))) return rc;
    if (staged) {
      CK(cudaMemcpyAsync(c.pin[s].buf[kPipeSlots - 1], d_out, m * out_bytes, cudaMemcpyDeviceToHost, st));
      CK(cudaEventRecord(c.pin[s].done, st));
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
  for (int s = 0; s < kPipeStreams; ++s) CK(cudaStreamSynchronize(c.stream[s]));
  return CUZK_OK;
}

// makes the padding constants E_0 .. E_{need-1} of `arity` available on the device.  The chain is sequential
// (ceil(arity/2) permutations per level), so it is computed only as far as trees need it, extended on demand, and the
// values -- constants of the hash function -- are cached on the host across cuzk_shutdown / cuzk_init cycles and devices.
// Upload and chain run on the legacy stream and are complete before this returns (the blocking copies and an explicit
// stream synchronisation), so consumers on any stream, blocking or not, see the constants.
int ensure_padding(Ctx &c, unsigned arity, int need = 2) {
  if (need > kMaxPadLevels) return fail(CUZK_ERR_INVALID, "tree too tall");
  std::lock_guard<std::mutex> lk(c.pad_mu);
  if (c.d_pad[arity] && c.pad_levels[arity] >= need) return CUZK_OK;
  if (!c.d_pad[arity]) {
    uint64_t *d = nullptr;
    CK(cudaMalloc(&d, (size_t)kMaxPadLevels * 32));
    c.d_pad[arity] = d;
    c.pad_levels[arity] = 0;
  }
  uint64_t *d = c.d_pad[arity];
  std::lock_guard<std::mutex> host_lk(g_pad_mu);   // the host copy is shared by all devices
  if (c.pad_levels[arity] < g_h_pad_levels[arity]) {   // bring the device copy up to what the host already knows
    CK(cudaMemcpy(d, g_h_pad[arity], (size_t)g_h_pad_levels[arity] * 32, cudaMemcpyHostToDevice));
    CK(cudaStreamSynchronize(0));   // a pageable upload may return before the DMA has landed
    c.pad_levels[arity] = g_h_pad_levels[arity];
  }
  if (c.pad_levels[arity] >= need) return CUZK_OK;
  const int start = c.pad_levels[arity], end = std::min(kMaxPadLevels, std::max(need, start + 4));
  if (g_coop_max.load() != 0) coop_padding_chain_kernel<<<1, 32>>>(reinterpret_cast<uint4 *>(d), (int)arity, start, end);
  else padding_chain_kernel<<<1, 32>>>(reinterpret_cast<uint4 *>(d), (int)arity, start, end);
  int rc = check_launch("padding_chain_kernel");
  if (rc) return rc;
  CK(cudaMemcpy(g_h_pad[arity][start], d + 4 * start, (size_t)(end - start) * 32, cudaMemcpyDeviceToHost));   // waits for the chain
  c.pad_levels[arity] = g_h_pad_levels[arity] = end;
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

// one level for `ntrees` trees of identical shape: narrow launches go to the cooperative kernel
// coop_cap: the widest level that may take the cooperative kernels (they cost 5-8x the instructions per node, which only
// pays while the chip is otherwise idle; builds that run several groups side by side lower it)
int launch_level(const uint4 *in, uint4 *out, size_t in_real, size_t out_count, unsigned arity, const uint4 *pad_in, cudaStream_t st,
                 size_t ntrees = 1, size_t tree_stride = 0, size_t coop_cap = ~(size_t)0) {
  if (out_count * ntrees == 0) return CUZK_OK;
  if (const CoopKind kind = out_count * ntrees <= coop_cap ? coop_kind(out_count * ntrees) : kOneThread) {
    CUZK_COOP_LAUNCH(kind, coop_merkle_level_kernel, out_count * ntrees, st, in, out, in_real, out_count, (int)arity, pad_in, pad_in + 2,
                     ntrees, tree_stride);
    return check_launch("coop_merkle_level_kernel");
  }
  merkle_level_kernel<<<grid_for(out_count * ntrees, kBlock), kBlock, 0, st>>>(in, out, in_real, out_count, (int)arity, pad_in, pad_in + 2,
                                                                               ntrees, tree_stride);
  return check_launch("merkle_level_kernel");
}

int subtree_streams_start(Ctx &c) {
  if (c.sub_fork) return CUZK_OK;
  for (int i = 0; i < kSubtreeStreams; ++i) {
    CK(cudaStreamCreateWithFlags(&c.sub_stream[i], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c.sub_join[i], cudaEventDisableTiming));
  }
  CK(cudaEventCreateWithFlags(&c.sub_fork, cudaEventDisableTiming));
  return CUZK_OK;
}
void subtree_streams_stop(Ctx &c) {
  for (int i = 0; i < kSubtreeStreams; ++i) {
    if (c.sub_stream[i]) cudaStreamDestroy(c.sub_stream[i]);
    if (c.sub_join[i]) cudaEventDestroy(c.sub_join[i]);
    c.sub_stream[i] = nullptr;
    c.sub_join[i] = nullptr;
  }
  if (c.sub_fork) cudaEventDestroy(c.sub_fork);
  c.sub_fork = nullptr;
}

// builds `ntrees` trees of n leaves each: one launch per level for the whole forest (build_tree_bottom_up,
// merkle_tree.cpp:66-97 / the per-level loop of merkle_tree_cuda.cu:188-192).  A single large tree is cut into groups of
// subtrees that run their lower levels on separate internal streams, so the narrow upper levels of one group (each costs a
// permutation latency however few nodes it has) hide behind the wide lower levels of the next; the levels above the cut run
// on the caller's stream after the groups have joined.  Every level is stored.
// leaves: ntrees x n elements; levels_out: ntrees flat level-major trees of total_nodes(n) elements each.
// padded_override (0 = derive from n) forces the padded leaf count (a power of arity >= n): shards of a larger tree.
// allow_groups = false keeps everything on `st` (callers that run several builds on the internal streams themselves; they
// pass the cooperative cap for builds that run side by side as level_coop_cap).
// pad_shift: the "leaves" are nodes of level pad_shift of a larger tree, so missing ones are the constant E_pad_shift.
int merkle_build_dev(Ctx &c, const uint64_t *leaves, size_t n, unsigned arity, uint64_t *levels_out, cudaStream_t st, size_t ntrees = 1,
                     size_t padded_override = 0, bool allow_groups = true, int pad_shift = 0, size_t level_coop_cap = ~(size_t)0) {
  const size_t padded = padded_override ? padded_override : cuzk_merkle_padded_leaves(n, arity);
  int nlv = 0;   // levels including the leaves
  size_t stride = 0;
  for (size_t p = padded;; p /= arity) {
    stride += p;
    ++nlv;
    if (p == 1) break;
  }
  int rc = ensure_padding(c, arity, nlv + 1 + pad_shift);
  if (rc) return rc;
  const uint4 *pad = reinterpret_cast<const uint4 *>(c.d_pad[arity]) + 2 * pad_shift;
  uint4 *base = reinterpret_cast<uint4 *>(levels_out);
  merkle_pad_leaves_kernel<<<grid_for(padded * ntrees, 256), 256, 0, st>>>(reinterpret_cast<const uint4 *>(leaves), n, padded, pad, base,
                                                                          ntrees, stride);
  if ((rc = check_launch("merkle_pad_leaves_kernel"))) return rc;

  // the cut: the tallest subtrees of which the real leaves still fill the wanted number of groups; only for trees large
  // enough that their lower levels are throughput-bound (several waves) while their upper ones are latency-bound
  const size_t want_groups = (size_t)std::max(1, g_build_groups.load(std::memory_order_relaxed));
  const int nstreams = std::min(kSubtreeStreams, std::max(1, g_build_streams.load(std::memory_order_relaxed)));
  const size_t group_coop = g_group_coop_max.load(std::memory_order_relaxed);
  size_t span = 1;
  int cut = 0;   // levels [0, cut) are hashed group by group, the rest on the caller's stream
  if (allow_groups && ntrees == 1 && c.sub_fork != nullptr && want_groups >= 2 && n >= ((size_t)1 << 17)) {
    while (cut < nlv - 1 && ceil_div(n, span * arity) >= want_groups) {
      span *= arity;
      ++cut;
    }
  }
  const size_t real_subtrees = ceil_div(n, span);
  const size_t groups = cut >= 2 ? std::min<size_t>(real_subtrees, want_groups) : 1;
  if (groups >= 2) {
    std::lock_guard<std::mutex> lk(c.sub_mu);
    CK(cudaEventRecord(c.sub_fork, st));
    const size_t total_subtrees = padded / span;
    for (int g = 0; g < nstreams && (size_t)g < groups; ++g) CK(cudaStreamWaitEvent(c.sub_stream[g], c.sub_fork, 0));
    for (size_t g = 0; g < groups; ++g) {
      const size_t lo = real_subtrees * g / groups;
      const size_t hi = (g + 1 == groups) ? total_subtrees : real_subtrees * (g + 1) / groups;   // the last group also writes the padding nodes
      cudaStream_t sg = c.sub_stream[g % nstreams];
      uint4 *cur = base;
      size_t p = padded, first = lo * span, width = (hi - lo) * span;   // this group's nodes on the current level
      size_t real = n > first ? std::min(n - first, width) : 0;
      for (int level = 0; level < cut; ++level) {
        if ((rc = launch_level(cur + 2 * first, cur + 2 * p + 2 * (first / arity), real, width / arity, arity, pad + 2 * level, sg, 1, 0,
                               group_coop)))
          return rc;
        cur += 2 * p;
        p /= arity;
        first /= arity;
        width /= arity;
        real = ceil_div(real, arity);
      }
    }
    for (int g = 0; g < nstreams && (size_t)g < groups; ++g) {
      CK(cudaEventRecord(c.sub_join[g], c.sub_stream[g]));
      CK(cudaStreamWaitEvent(st, c.sub_join[g], 0));
    }
  } else {
    cut = 0;
  }
  // the remaining levels, whole width, on the caller's stream
  uint4 *cur = base;
  size_t p = padded, real = n;
  for (int level = 0; level < cut; ++level) {
    cur += 2 * p;
    p /= arity;
    real = ceil_div(real, arity);
  }
  for (int level = cut; p > 1; ++level) {
    const size_t q = p / arity;
    if ((rc = launch_level(cur, cur + 2 * p, real, q, arity, pad + 2 * level, st, ntrees, stride, level_coop_cap))) return rc;
    cur += 2 * p;
    real = ceil_div(real, arity);
    p = q;
  }
  return CUZK_OK;
}

// roots of `count` consecutive subtrees of arity^height (virtual) leaves whose first n leaves are in memory
int subtree_roots_one_stream(Ctx &c, const uint64_t *leaves, size_t n, unsigned arity, unsigned height, size_t count, uint64_t *roots_out,
                             cudaStream_t st) {
  int rc = ensure_padding(c, arity, (int)height + 2);
  if (rc) return rc;
  const uint4 *pad = reinterpret_cast<const uint4 *>(c.d_pad[arity]);
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
  int flip = 0;
  for (unsigned level = 0; level < height; ++level) {
    const bool last = level + 1 == height;
    const size_t out_real = ceil_div(real, arity);
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
    rc = launch_level(cur, dst, real, out_count, arity, pad + 2 * level, st);
    if (rc) { release(); return rc; }
    cur = dst;
    real = out_real;
  }
  release();
  return CUZK_OK;
}

// The upper levels of a subtree are narrow: a level with fewer nodes than the chip has thread slots takes one node-hash
// latency however few nodes it has.  With several subtrees per call, groups of subtrees run on separate internal streams,
// so the narrow levels of one group hide behind the wide levels of the next.
int subtree_roots_dev(Ctx &c, const uint64_t *leaves, size_t n, unsigned arity, unsigned height, size_t count, uint64_t *roots_out,
                      cudaStream_t st) {
  size_t span = 1;
  for (unsigned i = 0; i < height; ++i) span *= arity;
  const size_t real_subtrees = ceil_div(n, span);   // subtrees with at least one real leaf; the rest are padding constants
  const size_t groups = std::min<size_t>(real_subtrees, (size_t)std::min(kSubtreeStreams, std::max(1, g_build_streams.load(std::memory_order_relaxed))));
  if (height < 3 || groups < 2 || c.sub_fork == nullptr) return subtree_roots_one_stream(c, leaves, n, arity, height, count, roots_out, st);
  std::lock_guard<std::mutex> lk(c.sub_mu);
  CK(cudaEventRecord(c.sub_fork, st));
  int rc = CUZK_OK;
  for (size_t g = 0; g < groups; ++g) {
    const size_t lo = real_subtrees * g / groups;
    const size_t hi = (g + 1 == groups) ? count : real_subtrees * (g + 1) / groups;   // the last group also writes the padding roots
    const size_t first_leaf = lo * span;
    const size_t n_g = std::min(n - first_leaf, (hi - lo) * span);
    cudaStream_t sg = c.sub_stream[g];
    CK(cudaStreamWaitEvent(sg, c.sub_fork, 0));
    const int r = subtree_roots_one_stream(c, leaves + 4 * first_leaf, n_g, arity, height, hi - lo, roots_out + 4 * lo, sg);
    if (r && !rc) rc = r;
    CK(cudaEventRecord(c.sub_join[g], sg));
    CK(cudaStreamWaitEvent(st, c.sub_join[g], 0));
  }
  return rc;
}

}  // namespace
