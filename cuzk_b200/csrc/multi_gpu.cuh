// multi_gpu.cuh -- cuzk_mg_*: the hot path over several B200s of one box, behind the C ABI (included once, at the end of
// cuzk_kernels.cu).  The reference is single-GPU (cudaSetDevice(0), field_arithmetic_cuda.cu:20); this layer is new and
// follows BASELINE.json's north star: leaves shard into contiguous subtrees, one block of subtrees per GPU; every GPU keeps
// ALL levels of its subtrees in HBM (they serve proofs); only the 32-byte subtree roots cross NVLink, in ONE NCCL all-gather,
// and every GPU hashes the few top levels itself.  Batch hashing and batch verification are independent units: contiguous
// slices per GPU, no collective.
//
// Two ways to span GPUs, one handle type:
//   cuzk_mg_init_local(ngpus, devices)          one process drives ngpus devices (ncclCommInitAll); what a C++ caller of the
//                                               drop-in classes uses (CudaNaryMerkleTree's multi-GPU constructor)
//   cuzk_mg_init_rank(rank, nranks, device, id) one process per GPU (torchrun); the 128-byte NCCL id comes from
//                                               cuzk_mg_unique_id on rank 0 and travels by whatever the host program has
// NCCL is loaded with dlopen at the first cuzk_mg_init_* (libnccl.so.2: the copy already in the process, e.g. torch's, or the
// system one), so single-GPU users of the library do not need it.
#pragma once
#include <dlfcn.h>

namespace {

// ---- the few NCCL entry points used, by hand (no header dependency) -----------------------------------------------
typedef void *nccl_comm_t;
struct nccl_unique_id {
  char internal[128];
};
constexpr int kNcclUint8 = 1;   // ncclUint8
struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(nccl_unique_id *) = nullptr;
  int (*CommInitRank)(nccl_comm_t *, int, nccl_unique_id, int) = nullptr;
  int (*CommInitAll)(nccl_comm_t *, int, const int *) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int *) = nullptr;
} g_nccl;
std::mutex g_nccl_mu;

int nccl_load() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.handle) return CUZK_OK;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(CUZK_ERR_CUDA, std::string("cuzk_mg: cannot load NCCL (libnccl.so.2): ") + dlerror());
  NcclApi a;
  a.handle = h;
#define CUZK_NCCL_SYM(field, name)                                                        \
  *reinterpret_cast<void **>(&a.field) = dlsym(h, name);                                  \
  if (!a.field) return fail(CUZK_ERR_CUDA, std::string("cuzk_mg: NCCL symbol missing: ") + name)
  CUZK_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
  CUZK_NCCL_SYM(CommInitRank, "ncclCommInitRank");
  CUZK_NCCL_SYM(CommInitAll, "ncclCommInitAll");
  CUZK_NCCL_SYM(CommDestroy, "ncclCommDestroy");
  CUZK_NCCL_SYM(AllGather, "ncclAllGather");
  CUZK_NCCL_SYM(GroupStart, "ncclGroupStart");
  CUZK_NCCL_SYM(GroupEnd, "ncclGroupEnd");
  CUZK_NCCL_SYM(GetErrorString, "ncclGetErrorString");
  CUZK_NCCL_SYM(GetVersion, "ncclGetVersion");
#undef CUZK_NCCL_SYM
  g_nccl = a;
  return CUZK_OK;
}
int nccl_fail(int r, const char *what) {
  return fail(CUZK_ERR_CUDA, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error"));
}
#define NK(call)                               \
  do {                                         \
    int r__ = (call);                          \
    if (r__ != 0) return nccl_fail(r__, #call); \
  } while (0)

constexpr int kMaxLocal = 16;

// the shard plan: the padded tree is cut at the tallest level that still gives every rank a subtree with real leaves;
// real subtrees are dealt to ranks in contiguous blocks of per_rank (cuzk_b200/distributed.py: plan_merkle_shards)
struct MgPlan {
  size_t n = 0, padded = 0, span = 1, total_subtrees = 1, real_subtrees = 1, per_rank = 1;
  unsigned arity = 2, levels = 0, height = 0;   // levels: hash levels of the whole tree; height: of one subtree
  void subtrees_of(int rank, size_t &lo, size_t &hi) const {
    lo = std::min((size_t)rank * per_rank, real_subtrees);
    hi = std::min(lo + per_rank, real_subtrees);
  }
  void leaves_of(int rank, size_t &l0, size_t &l1) const {
    size_t lo, hi;
    subtrees_of(rank, lo, hi);
    l0 = std::min(lo * span, n);
    l1 = std::min(hi * span, n);
  }
};
MgPlan mg_plan(size_t n, unsigned arity, int world) {
  MgPlan p;
  p.n = n;
  p.arity = arity;
  p.padded = 1;
  while (p.padded < n) {
    p.padded *= arity;
    ++p.levels;
  }
  while (p.height < p.levels) {
    const size_t nxt = p.span * arity;
    if ((n + nxt - 1) / nxt < (size_t)world) break;
    p.span = nxt;
    ++p.height;
  }
  p.real_subtrees = (n + p.span - 1) / p.span;
  p.total_subtrees = p.padded / p.span;
  p.per_rank = (p.real_subtrees + world - 1) / world;
  return p;
}

// proofs from a sharded tree: one thread per (proof, level).  Levels below the cut come from the shard's forest (subtree t at
// shard + t * sub_total, level-major), levels above from the replicated top tree.  Leaves this shard does not own: sentinel.
__global__ void mg_prove_kernel(const uint4 *__restrict__ shard, size_t sub_total, size_t span, int height, size_t first_subtree,
                                size_t num_subtrees, const uint4 *__restrict__ top, size_t total_subtrees, int arity, int nlv, size_t n,
                                const u64 *__restrict__ indices, size_t num_proofs, uint4 *__restrict__ sib, u32 *__restrict__ pos) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= num_proofs * (size_t)nlv) return;
  const size_t q = t / nlv;
  const int l = (int)(t % nlv);
  u64 idx = indices[q];
  const size_t subtree = idx / span;
  if (idx >= n || subtree < first_subtree || subtree >= first_subtree + num_subtrees) {
    pos[t] = 0xFFFFFFFFu;
    return;
  }
  const uint4 *levels;
  size_t p;
  int k;
  if (l < height) {
    levels = shard + 2 * (subtree - first_subtree) * sub_total;
    idx -= subtree * span;
    p = span;
    k = l;
  } else {
    levels = top;
    idx = subtree;
    p = total_subtrees;
    k = l - height;
  }
  size_t off = 0;
  for (int i = 0; i < k; ++i) {
    off += p;
    p /= arity;
    idx /= arity;
  }
  const u32 my = (u32)(idx % arity);
  const size_t base = off + (idx - my);
  pos[t] = my;
  uint4 *dst = sib + 2 * t * (size_t)(arity - 1);
  int w = 0;
  for (int ch = 0; ch < arity; ++ch) {
    if (ch == (int)my) continue;
    dst[2 * w] = levels[2 * (base + ch)];
    dst[2 * w + 1] = levels[2 * (base + ch) + 1];
    ++w;
  }
}

}  // namespace

struct cuzk_mg {
  int nranks = 1, nlocal = 1, first_rank = 0;
  int device[kMaxLocal] = {};
  nccl_comm_t comm[kMaxLocal] = {};
  cudaStream_t stream[kMaxLocal] = {};
};

struct cuzk_mg_tree {
  cuzk_mg *mg = nullptr;
  MgPlan plan;
  size_t sub_total = 0;                 // nodes of one subtree (all its levels)
  size_t top_total = 0;                 // nodes of the top tree
  uint64_t *shard[kMaxLocal] = {};      // per local device: its subtrees' levels
  size_t shard_first[kMaxLocal] = {}, shard_count[kMaxLocal] = {};
  uint64_t *top[kMaxLocal] = {};        // per local device: the levels above the cut (replicated)
  uint64_t *gather_buf[kMaxLocal] = {}; // per local device: send slot + gathered roots + top-level leaves
  uint64_t root[4] = {};
};

extern "C" const char *cuzk_last_error(void);
namespace {
// runs fn(local device index, first unit, unit count) for contiguous slices of n units, one host thread per local device
template <class Fn>
int mg_parallel_slices(const cuzk_mg *mg, size_t n, Fn fn) {
  std::vector<int> rcs(mg->nlocal, CUZK_OK);
  std::vector<std::string> errs(mg->nlocal);
  std::vector<std::thread> th;
  for (int r = 0; r < mg->nlocal; ++r) {
    const size_t a = n * (size_t)r / mg->nlocal, b = n * (size_t)(r + 1) / mg->nlocal;
    if (b == a) continue;
    th.emplace_back([&, r, a, b] {
      cudaSetDevice(mg->device[r]);
      rcs[r] = fn(r, a, b - a);
      if (rcs[r]) errs[r] = cuzk_last_error();
    });
  }
  for (auto &x : th) x.join();
  for (int r = 0; r < mg->nlocal; ++r)
    if (rcs[r]) return fail(rcs[r], errs[r]);
  return CUZK_OK;
}

}  // namespace

extern "C" {

int cuzk_mg_unique_id(uint8_t out[128]) {
  int rc = nccl_load();
  if (rc) return rc;
  if (!out) return fail(CUZK_ERR_INVALID, "null pointer");
  nccl_unique_id id;
  NK(g_nccl.GetUniqueId(&id));
  memcpy(out, id.internal, 128);
  return CUZK_OK;
}

static int mg_finish_init(cuzk_mg *mg) {
  for (int r = 0; r < mg->nlocal; ++r) {
    int rc = cuzk_init(mg->device[r]);
    if (rc) return rc;
    CK(cudaStreamCreateWithFlags(&mg->stream[r], cudaStreamNonBlocking));
  }
  return CUZK_OK;
}

int cuzk_mg_init_local(int ngpus, const int *devices, cuzk_mg_t **out) {
  if (!out) return fail(CUZK_ERR_INVALID, "null pointer");
  *out = nullptr;
  if (ngpus < 1 || ngpus > kMaxLocal) return fail(CUZK_ERR_INVALID, "cuzk_mg_init_local: 1..16 devices");
  if (ngpus > cuzk_device_count()) return fail(CUZK_ERR_INVALID, "cuzk_mg_init_local: more devices asked for than the box has");
  int prev = 0;
  cudaGetDevice(&prev);
  cuzk_mg *mg = new (std::nothrow) cuzk_mg;
  if (!mg) return fail(CUZK_ERR_INVALID, "out of host memory");
  mg->nranks = mg->nlocal = ngpus;
  for (int r = 0; r < ngpus; ++r) mg->device[r] = devices ? devices[r] : r;
  int rc = mg_finish_init(mg);
  if (!rc && ngpus > 1) {
    if (!(rc = nccl_load())) {
      const int r2 = g_nccl.CommInitAll(mg->comm, ngpus, mg->device);
      if (r2 != 0) rc = nccl_fail(r2, "ncclCommInitAll");
    }
  }
  cudaSetDevice(prev);
  if (rc) { delete mg; return rc; }
  *out = mg;
  return CUZK_OK;
}

int cuzk_mg_init_rank(int rank, int nranks, int device, const uint8_t id[128], cuzk_mg_t **out) {
  if (!out) return fail(CUZK_ERR_INVALID, "null pointer");
  *out = nullptr;
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(CUZK_ERR_INVALID, "cuzk_mg_init_rank: bad rank");
  if (nranks > 1 && !id) return fail(CUZK_ERR_INVALID, "cuzk_mg_init_rank: NCCL id needed");
  cuzk_mg *mg = new (std::nothrow) cuzk_mg;
  if (!mg) return fail(CUZK_ERR_INVALID, "out of host memory");
  mg->nranks = nranks;
  mg->nlocal = 1;
  mg->first_rank = rank;
  mg->device[0] = device;
  int rc = mg_finish_init(mg);
  if (!rc && nranks > 1) {
    if (!(rc = nccl_load())) {
      nccl_unique_id uid;
      memcpy(uid.internal, id, 128);
      const int r2 = g_nccl.CommInitRank(&mg->comm[0], nranks, uid, rank);
      if (r2 != 0) rc = nccl_fail(r2, "ncclCommInitRank");
    }
  }
  if (rc) { delete mg; return rc; }
  *out = mg;
  return CUZK_OK;
}

int cuzk_mg_free(cuzk_mg_t *mg) {
  if (!mg) return CUZK_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  for (int r = 0; r < mg->nlocal; ++r) {
    cudaSetDevice(mg->device[r]);
    if (mg->stream[r]) {
      cudaStreamSynchronize(mg->stream[r]);
      cudaStreamDestroy(mg->stream[r]);
    }
    if (mg->comm[r] && g_nccl.CommDestroy) g_nccl.CommDestroy(mg->comm[r]);
    cuzk_shutdown();
  }
  cudaSetDevice(prev);
  delete mg;
  return CUZK_OK;
}

int cuzk_mg_world(const cuzk_mg_t *mg, int *nranks, int *nlocal, int *first_rank) {
  if (!mg) return fail(CUZK_ERR_INVALID, "null handle");
  if (nranks) *nranks = mg->nranks;
  if (nlocal) *nlocal = mg->nlocal;
  if (first_rank) *first_rank = mg->first_rank;
  return CUZK_OK;
}
int cuzk_mg_device(const cuzk_mg_t *mg, int local) { return (mg && local >= 0 && local < mg->nlocal) ? mg->device[local] : -1; }
void *cuzk_mg_stream(const cuzk_mg_t *mg, int local) { return (mg && local >= 0 && local < mg->nlocal) ? mg->stream[local] : nullptr; }
int cuzk_mg_nccl_version(void) {
  int v = 0;
  if (nccl_load() || g_nccl.GetVersion(&v) != 0) return 0;
  return v;
}

// which leaves rank `rank` of `nranks` holds in a sharded tree over n leaves: [first, first + count)
int cuzk_mg_shard_leaves(size_t n, unsigned arity, int nranks, int rank, size_t *first_out, size_t *count_out) {
  int rc = check_arity(arity);
  if (rc) return rc;
  if (n == 0 || nranks < 1 || rank < 0 || rank >= nranks) return fail(CUZK_ERR_INVALID, "bad shard arguments");
  const MgPlan p = mg_plan(n, arity, nranks);
  size_t l0, l1;
  p.leaves_of(rank, l0, l1);
  if (first_out) *first_out = l0;
  if (count_out) *count_out = l1 - l0;
  return CUZK_OK;
}

int cuzk_mg_tree_free(cuzk_mg_tree_t *t) {
  if (!t) return CUZK_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  for (int r = 0; r < t->mg->nlocal; ++r) {
    cudaSetDevice(t->mg->device[r]);
    cudaStream_t st = t->mg->stream[r];
    if (t->shard[r]) cudaFreeAsync(t->shard[r], st);
    if (t->top[r]) cudaFreeAsync(t->top[r], st);
    if (t->gather_buf[r]) cudaFreeAsync(t->gather_buf[r], st);
  }
  cudaSetDevice(prev);
  delete t;
  return CUZK_OK;
}

// Builds the sharded tree.  local_leaves[r]: the leaves of local device r's rank (cuzk_mg_shard_leaves), in that device's
// memory (mem == CUZK_MEM_DEVICE), or -- mem == CUZK_MEM_HOST -- local_leaves[0] is the WHOLE leaf array in host memory and
// every local device uploads its own slice.  Collective in one-process-per-GPU mode: every rank must call it.
int cuzk_mg_tree_build(cuzk_mg_t *mg, const uint64_t *const *local_leaves, size_t n, unsigned arity, int mem, cuzk_mg_tree_t **out) {
  if (!mg || !out || !local_leaves) return fail(CUZK_ERR_INVALID, "null pointer");
  *out = nullptr;
  int rc = check_arity(arity);
  if (rc) return rc;
  if (n == 0) return fail(CUZK_ERR_INVALID, "cuzk_mg_tree_build: n must be >= 1");
  cuzk_mg_tree *t = new (std::nothrow) cuzk_mg_tree;
  if (!t) return fail(CUZK_ERR_INVALID, "out of host memory");
  t->mg = mg;
  t->plan = mg_plan(n, arity, mg->nranks);
  const MgPlan &P = t->plan;
  t->sub_total = cuzk_merkle_total_nodes(P.span, arity);
  t->top_total = cuzk_merkle_total_nodes(P.total_subtrees, arity);
  int prev = 0;
  cudaGetDevice(&prev);
  auto bail = [&](int code) {
    cudaSetDevice(prev);
    cuzk_mg_tree_free(t);
    return code;
  };
  // phase 1, per local device: the subtrees of its rank, every level kept; roots into the gather send slot
  for (int r = 0; r < mg->nlocal; ++r) {
    const int rank = mg->first_rank + r;
    CtxGuard guard(mg->device[r]);
    if (!guard.get()) return bail(fail(CUZK_ERR_CUDA, CUZK_NOT_INIT_MSG));
    Ctx &c = *guard.get();
    cudaStream_t st = mg->stream[r];
    size_t lo, hi, l0, l1;
    P.subtrees_of(rank, lo, hi);
    P.leaves_of(rank, l0, l1);
    t->shard_first[r] = lo;
    t->shard_count[r] = hi - lo;
    cudaError_t e;
    // gather buffer: [send slot: per_rank][gathered: per_rank * nranks][top-tree leaves: total_subtrees]
    const size_t gb = (P.per_rank * (1 + (size_t)mg->nranks) + P.total_subtrees) * 32;
    if ((e = cudaMallocAsync(reinterpret_cast<void **>(&t->gather_buf[r]), gb, st)) != cudaSuccess) return bail(cuda_fail(e, "cudaMallocAsync"));
    if ((e = cudaMemsetAsync(t->gather_buf[r], 0, P.per_rank * 32, st)) != cudaSuccess) return bail(cuda_fail(e, "cudaMemsetAsync"));
    if ((e = cudaMallocAsync(reinterpret_cast<void **>(&t->top[r]), t->top_total * 32, st)) != cudaSuccess) return bail(cuda_fail(e, "cudaMallocAsync"));
    if (hi == lo) continue;
    if ((e = cudaMallocAsync(reinterpret_cast<void **>(&t->shard[r]), (hi - lo) * t->sub_total * 32, st)) != cudaSuccess)
      return bail(cuda_fail(e, "cudaMallocAsync(shard levels)"));
    const uint64_t *leaves = nullptr;
    void *staged = nullptr;
    if (mem == CUZK_MEM_DEVICE) {
      leaves = local_leaves[r];
    } else {
      if ((e = cudaMallocAsync(&staged, (l1 - l0) * 32, st)) != cudaSuccess) return bail(cuda_fail(e, "cudaMallocAsync(leaves)"));
      std::lock_guard<std::mutex> lk(c.hp_mu);
      if ((rc = bulk_upload(c, staged, local_leaves[0] + 4 * l0, (l1 - l0) * 32, st))) return bail(rc);
      leaves = static_cast<const uint64_t *>(staged);
    }
    if (!leaves) return bail(fail(CUZK_ERR_INVALID, "cuzk_mg_tree_build: missing leaves of a local shard"));
    const size_t nsub_here = hi - lo;
    if (nsub_here == 1) {
      // one large subtree: on the handle's stream, cut into groups over the internal streams by merkle_build_dev itself
      // (two subtrees are faster side by side on two streams: 108.6 ms against 112.1 ms for 2 x 2^24 leaves, arity 8)
      for (size_t s = 0; s < nsub_here; ++s) {
        const size_t first = s * P.span;
        const size_t n_s = std::min(P.span, (l1 - l0) - first);
        uint64_t *dst = t->shard[r] + 4 * s * t->sub_total;
        if ((rc = merkle_build_dev(c, leaves + 4 * first, n_s, arity, dst, st, 1, P.span, /*allow_groups=*/true))) return bail(rc);
        if ((e = cudaMemcpyAsync(t->gather_buf[r] + 4 * s, dst + 4 * (t->sub_total - 1), 32, cudaMemcpyDeviceToDevice, st)) != cudaSuccess)
          return bail(cuda_fail(e, "cudaMemcpyAsync(root)"));
      }
    } else {
      // several subtrees: dealt over the internal streams, so that the narrow upper levels of one hide behind the wide levels
      // of the next
      std::lock_guard<std::mutex> lk(c.sub_mu);
      if ((e = cudaEventRecord(c.sub_fork, st)) != cudaSuccess) return bail(cuda_fail(e, "cudaEventRecord"));
      const size_t nsub = hi - lo;
      bool used[kSubtreeStreams] = {};
      for (size_t s = 0; s < nsub; ++s) {
        const int g = (int)(s % (size_t)std::min(kSubtreeStreams, std::max(1, g_build_streams.load(std::memory_order_relaxed))));
        cudaStream_t sg = c.sub_stream[g];
        if (!used[g]) {
          if ((e = cudaStreamWaitEvent(sg, c.sub_fork, 0)) != cudaSuccess) return bail(cuda_fail(e, "cudaStreamWaitEvent"));
          used[g] = true;
        }
        const size_t first = s * P.span;
        const size_t n_s = std::min(P.span, (l1 - l0) - first);
        uint64_t *dst = t->shard[r] + 4 * s * t->sub_total;
        if ((rc = merkle_build_dev(c, leaves + 4 * first, n_s, arity, dst, sg, 1, P.span, /*allow_groups=*/false, 0,
                                   nsub > 1 ? g_group_coop_max.load(std::memory_order_relaxed) : ~(size_t)0)))
          return bail(rc);
        // the subtree's root goes straight into the all-gather send slot
        if ((e = cudaMemcpyAsync(t->gather_buf[r] + 4 * s, dst + 4 * (t->sub_total - 1), 32, cudaMemcpyDeviceToDevice, sg)) != cudaSuccess)
          return bail(cuda_fail(e, "cudaMemcpyAsync(root)"));
      }
      for (int g = 0; g < kSubtreeStreams; ++g) {
        if (!used[g]) continue;
        if ((e = cudaEventRecord(c.sub_join[g], c.sub_stream[g])) != cudaSuccess) return bail(cuda_fail(e, "cudaEventRecord"));
        if ((e = cudaStreamWaitEvent(st, c.sub_join[g], 0)) != cudaSuccess) return bail(cuda_fail(e, "cudaStreamWaitEvent"));
      }
    }
    if (staged) cudaFreeAsync(staged, st);
  }
  // phase 2: ONE all-gather of per_rank x 32 bytes per rank
  if (mg->nranks > 1) {
    if (mg->nlocal > 1) {
      const int r0 = g_nccl.GroupStart();
      if (r0 != 0) return bail(nccl_fail(r0, "ncclGroupStart"));
    }
    for (int r = 0; r < mg->nlocal; ++r) {
      cudaSetDevice(mg->device[r]);
      uint64_t *send = t->gather_buf[r], *recv = send + 4 * P.per_rank;
      const int r1 = g_nccl.AllGather(send, recv, P.per_rank * 32, kNcclUint8, mg->comm[r], mg->stream[r]);
      if (r1 != 0) return bail(nccl_fail(r1, "ncclAllGather"));
    }
    if (mg->nlocal > 1) {
      const int r2 = g_nccl.GroupEnd();
      if (r2 != 0) return bail(nccl_fail(r2, "ncclGroupEnd"));
    }
  }
  // phase 3, per local device: top-tree leaves = gathered roots + padding constants, then the top levels on the same stream
  for (int r = 0; r < mg->nlocal; ++r) {
    CtxGuard guard(mg->device[r]);
    if (!guard.get()) return bail(fail(CUZK_ERR_CUDA, CUZK_NOT_INIT_MSG));
    Ctx &c = *guard.get();
    cudaStream_t st = mg->stream[r];
    uint64_t *send = t->gather_buf[r];
    uint64_t *gathered = mg->nranks > 1 ? send + 4 * P.per_rank : send;
    uint64_t *nodes = send + 4 * P.per_rank * (1 + (size_t)mg->nranks);
    cudaError_t e;
    if ((e = cudaMemcpyAsync(nodes, gathered, P.real_subtrees * 32, cudaMemcpyDeviceToDevice, st)) != cudaSuccess)
      return bail(cuda_fail(e, "cudaMemcpyAsync(roots)"));
    // all-padding subtrees are the constant E_height: merkle_build_dev pads with the level-0 constant of ITS tree, so the top
    // tree is built with its real leaf count and the padding constants shifted by `height` levels
    if ((rc = merkle_build_dev(c, nodes, P.real_subtrees, arity, t->top[r], st, 1, P.total_subtrees, false, (int)P.height))) return bail(rc);
  }
  // the root (host copy) from local device 0
  {
    cudaSetDevice(mg->device[0]);
    cudaError_t e = cudaMemcpyAsync(t->root, t->top[0] + 4 * (t->top_total - 1), 32, cudaMemcpyDeviceToHost, mg->stream[0]);
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMemcpyAsync(root)"));
    for (int r = 0; r < mg->nlocal; ++r) {
      cudaSetDevice(mg->device[r]);
      if ((e = cudaStreamSynchronize(mg->stream[r])) != cudaSuccess) return bail(cuda_fail(e, "cudaStreamSynchronize"));
    }
  }
  cudaSetDevice(prev);
  *out = t;
  return CUZK_OK;
}

int cuzk_mg_tree_root(const cuzk_mg_tree_t *t, uint64_t root_out[4]) {
  if (!t || !root_out) return fail(CUZK_ERR_INVALID, "null pointer");
  memcpy(root_out, t->root, 32);
  return CUZK_OK;
}
size_t cuzk_mg_tree_num_levels(const cuzk_mg_tree_t *t) { return t ? t->plan.levels + 1 : 0; }
size_t cuzk_mg_tree_leaf_count(const cuzk_mg_tree_t *t) { return t ? t->plan.n : 0; }
size_t cuzk_mg_tree_subtree_height(const cuzk_mg_tree_t *t) { return t ? t->plan.height : 0; }
const uint64_t *cuzk_mg_tree_shard_levels(const cuzk_mg_tree_t *t, int local, size_t *first_subtree, size_t *num_subtrees, size_t *nodes_per_subtree) {
  if (!t || local < 0 || local >= t->mg->nlocal) return nullptr;
  if (first_subtree) *first_subtree = t->shard_first[local];
  if (num_subtrees) *num_subtrees = t->shard_count[local];
  if (nodes_per_subtree) *nodes_per_subtree = t->sub_total;
  return t->shard[local];
}

// Proofs for `num_proofs` leaf indices (host memory), served from the levels the shards keep: each query is routed to the
// local device whose rank owns the leaf.  Output rows in the level-uniform wire format of cuzk_merkle_prove_batch, levels of
// the whole tree.  Queries owned by a rank of ANOTHER process (one-process-per-GPU mode) and indices >= n get the sentinel
// position 0xFFFFFFFF on every level, as out-of-range indices do in cuzk_merkle_prove_batch.
int cuzk_mg_tree_prove_batch(const cuzk_mg_tree_t *t, const uint64_t *indices, size_t num_proofs, uint64_t *siblings_out,
                             uint32_t *positions_out) {
  if (!t) return fail(CUZK_ERR_INVALID, "null tree");
  const MgPlan &P = t->plan;
  const size_t nlv = P.levels;
  if (num_proofs == 0 || nlv == 0) return CUZK_OK;
  if (!indices || !siblings_out || !positions_out) return fail(CUZK_ERR_INVALID, "null pointer");
  const cuzk_mg *mg = t->mg;
  const size_t sib_row = nlv * (P.arity - 1) * 4;   // u64 words per proof
  // route
  std::vector<std::vector<uint64_t>> idx(mg->nlocal);
  std::vector<std::vector<size_t>> row(mg->nlocal);
  for (size_t q = 0; q < num_proofs; ++q) {
    int owner = -1;
    if (indices[q] < P.n) {
      const size_t rank = (indices[q] / P.span) / P.per_rank;
      if ((int)rank >= mg->first_rank && (int)rank < mg->first_rank + mg->nlocal) owner = (int)rank - mg->first_rank;
    }
    if (owner < 0) {
      for (size_t l = 0; l < nlv; ++l) positions_out[q * nlv + l] = 0xFFFFFFFFu;
      continue;
    }
    idx[owner].push_back(indices[q]);
    row[owner].push_back(q);
  }
  int prev = 0;
  cudaGetDevice(&prev);
  int rc = CUZK_OK;
  std::vector<std::vector<uint64_t>> sib(mg->nlocal);
  std::vector<std::vector<uint32_t>> pos(mg->nlocal);
  std::vector<void *> scratch(mg->nlocal, nullptr);
  for (int r = 0; r < mg->nlocal && !rc; ++r) {   // enqueue on every device, then collect
    const size_t m = idx[r].size();
    if (m == 0) continue;
    cudaSetDevice(mg->device[r]);
    cudaStream_t st = mg->stream[r];
    const size_t threads = m * nlv, sib_bytes = m * sib_row * 8, pos_bytes = threads * 4;
    const size_t pos_off = sib_bytes, idx_off = (pos_off + pos_bytes + 255) & ~(size_t)255;
    cudaError_t e = cudaMallocAsync(&scratch[r], idx_off + m * 8, st);
    if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMallocAsync(proofs)"); break; }
    char *base = static_cast<char *>(scratch[r]);
    sib[r].resize(m * sib_row);
    pos[r].resize(threads);
    if ((e = cudaMemcpyAsync(base + idx_off, idx[r].data(), m * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) { rc = cuda_fail(e, "cudaMemcpyAsync"); break; }
    mg_prove_kernel<<<grid_for(threads, 256), 256, 0, st>>>(reinterpret_cast<const uint4 *>(t->shard[r]), t->sub_total, P.span, (int)P.height,
                                                           t->shard_first[r], t->shard_count[r], reinterpret_cast<const uint4 *>(t->top[r]),
                                                           P.total_subtrees, (int)P.arity, (int)nlv, P.n,
                                                           reinterpret_cast<const u64 *>(base + idx_off), m, reinterpret_cast<uint4 *>(base),
                                                           reinterpret_cast<u32 *>(base + pos_off));
    if ((rc = check_launch("mg_prove_kernel"))) break;
    if ((e = cudaMemcpyAsync(sib[r].data(), base, sib_bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(pos[r].data(), base + pos_off, pos_bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess) {
      rc = cuda_fail(e, "cudaMemcpyAsync(proofs)");
      break;
    }
  }
  for (int r = 0; r < mg->nlocal; ++r) {
    if (!scratch[r]) continue;
    cudaSetDevice(mg->device[r]);
    cudaError_t e = cudaStreamSynchronize(mg->stream[r]);
    if (e != cudaSuccess && !rc) rc = cuda_fail(e, "cudaStreamSynchronize");
    cudaFreeAsync(scratch[r], mg->stream[r]);
    if (rc) continue;
    for (size_t i = 0; i < row[r].size(); ++i) {
      memcpy(siblings_out + row[r][i] * sib_row, sib[r].data() + i * sib_row, sib_row * 8);
      memcpy(positions_out + row[r][i] * nlv, pos[r].data() + i * nlv, nlv * 4);
    }
  }
  cudaSetDevice(prev);
  return rc;
}

// batch verification against the sharded tree's root (host memory): proofs are independent, contiguous slices per local device
int cuzk_mg_tree_verify_batch(const cuzk_mg_tree_t *t, const uint64_t *leaf_values, const uint64_t *siblings, const uint32_t *positions,
                              uint8_t *results_out, size_t num_proofs) {
  if (!t) return fail(CUZK_ERR_INVALID, "null tree");
  if (num_proofs == 0) return CUZK_OK;
  const size_t nlv = t->plan.levels, a1 = t->plan.arity - 1;
  if (!leaf_values || !results_out || (nlv && (!siblings || !positions))) return fail(CUZK_ERR_INVALID, "null pointer");
  return mg_parallel_slices(t->mg, num_proofs, [&](int, size_t first, size_t count) {
    return cuzk_merkle_verify_batch(leaf_values + 4 * first, siblings ? siblings + 4 * first * nlv * a1 : nullptr,
                                    positions ? positions + first * nlv : nullptr, nlv, t->plan.arity, t->root, results_out + first, count,
                                    CUZK_MEM_HOST, nullptr);
  });
}

// batch pair hashing over the local devices (host memory): contiguous slices, no collective
int cuzk_mg_poseidon_hash_pairs(const cuzk_mg_t *mg, const uint64_t *left, const uint64_t *right, uint64_t *out, size_t n) {
  if (!mg) return fail(CUZK_ERR_INVALID, "null handle");
  if (n == 0) return CUZK_OK;
  if (!left || !right || !out) return fail(CUZK_ERR_INVALID, "null pointer");
  return mg_parallel_slices(mg, n, [&](int, size_t first, size_t count) {
    return cuzk_poseidon_hash_pairs(left + 4 * first, right + 4 * first, out + 4 * first, count, CUZK_MEM_HOST, nullptr);
  });
}

}  // extern "C"
