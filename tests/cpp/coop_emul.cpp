// coop_emul.cpp -- TEST INFRASTRUCTURE: runs the cooperative lane-group source (cuzk_b200/csrc/coop.cuh; layout Wide16, or
// Narrow8 with -DCUZK_EMUL_LANES=8) on the host, one thread per lane in lockstep (shuffles and votes go through a
// shared exchange buffer and a spinning barrier), and compares every result with the plain-C oracle.  It checks the
// arithmetic of the cooperative path -- in particular that a unit whose `unc` vote is clear is bit-exact -- without a GPU.
//
// usage: coop_emul [units] [seed] [plain]      exit code 0 = no unflagged mismatch
#define CUZK_COOP_HOST_EMUL 1
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../cuzk_b200/csrc/coop.cuh"
#include "../../cuzk_b200/csrc/coop16.cuh"

extern "C" {
void cuzk_oracle_round_constants(uint64_t *out);
void cuzk_oracle_fr_mul(const uint64_t a[4], const uint64_t b[4], uint64_t r[4]);
void cuzk_oracle_fr_add(const uint64_t a[4], const uint64_t b[4], uint64_t r[4]);
void cuzk_oracle_permutation(uint64_t *state);
void cuzk_oracle_batch_sponge(const uint64_t *in, size_t width, uint64_t ds, uint64_t *out, size_t n);
void cuzk_oracle_batch_mds_layer(uint64_t *states, size_t n);
}

namespace {
#ifndef CUZK_EMUL_LANES
#define CUZK_EMUL_LANES 16
#endif
constexpr int kLanes = CUZK_EMUL_LANES;
thread_local uint32_t tl_lane;
std::atomic<uint32_t> g_count{0};
std::atomic<uint32_t> g_gen{0};
volatile uint32_t g_slot[2][kLanes];
thread_local uint32_t tl_parity = 0;

void barrier() {
  const uint32_t gen = g_gen.load(std::memory_order_acquire);
  if (g_count.fetch_add(1, std::memory_order_acq_rel) + 1 == kLanes) {
    g_count.store(0, std::memory_order_relaxed);
    g_gen.store(gen + 1, std::memory_order_release);
  } else {
    while (g_gen.load(std::memory_order_acquire) == gen) {
      std::this_thread::yield();   // more lanes than cores: give the others the CPU
    }
  }
}
}  // namespace

namespace cuzk {
namespace coop {
u32 emul_lane() { return tl_lane; }
u32 emul_shfl(u32 x, u32 src) {
  const uint32_t par = tl_parity;
  tl_parity ^= 1u;
  g_slot[par][tl_lane] = x;
  std::atomic_thread_fence(std::memory_order_seq_cst);
  barrier();
  const u32 v = g_slot[par][src & (u32)(kLanes - 1)];
  return v;
}
u32 emul_ballot(bool p) {
  const uint32_t par = tl_parity;
  tl_parity ^= 1u;
  g_slot[par][tl_lane] = p ? 1u : 0u;
  std::atomic_thread_fence(std::memory_order_seq_cst);
  barrier();
  u32 b = 0;
  for (int i = 0; i < kLanes; ++i) b |= (g_slot[par][i] & 1u) << i;
  return b;
}
}  // namespace coop
}  // namespace cuzk

using namespace cuzk;
#if CUZK_EMUL_LANES == 8
typedef coop::Narrow8 Y;
#else
typedef coop::Wide16 Y;
#endif
// one interface over the templated algorithms (coop.cuh) and the non-templated sixteen-lane file the kernels use (coop16.cuh,
// -DCUZK_EMUL_LEGACY16)
#ifdef CUZK_EMUL_LEGACY16
namespace impl {
typedef coop16::Lane Lane;
typedef coop16::Flags Flags;
inline Lane make_lane() { return coop16::make_lane(); }
inline void gather(u32 (&r)[8], u32 x) { coop16::gather(r, x); }
inline void mulred1(u32 (&r)[1], const u32 (&a)[1][8], const u32 (&b)[1], const Lane &L, Flags &F) { coop16::mulred<1>(r, a, b, L, F); }
inline void add_reduce1(u32 (&r)[1], const u32 (&a)[1], const u32 (&b)[1], const Lane &L, Flags &F) { coop16::add_reduce<1>(r, a, b, L, F); }
template <class Rc> inline void permute(u32 (&s)[3], const Rc &rct, const Lane &L, Flags &F) { coop16::permute(s, rct, L, F); }
template <class Rc, class Ld> inline u32 sponge(u32 &o, u32 a, u32 b, int w, const Rc &rct, const Lane &L, Ld ld) { return coop16::sponge(o, a, b, w, rct, L, ld); }
inline void mds_arc(u32 (&s)[3], const u32 (&rc)[3], bool h, const Lane &L, Flags &F) { coop16::mds_arc(s, rc, h, L, F); }
inline bool flagged(const Flags &F) { return coop16::flagged(F); }
inline u32 carry_exact(u32 lo, u32 c, const Lane &L) { return coop16::carry_exact8(lo, c, L); }
}  // namespace impl
#else
namespace impl {
typedef coop::Lane Lane;
typedef coop::Flags Flags;
inline Lane make_lane() { return Y::make_lane(); }
inline void gather(u32 (&r)[8], u32 x) { coop::gather<Y>(r, x); }
inline void mulred1(u32 (&r)[1], const u32 (&a)[1][8], const u32 (&b)[1], const Lane &L, Flags &F) { coop::mulred<Y, 1>(r, a, b, L, F); }
inline void add_reduce1(u32 (&r)[1], const u32 (&a)[1], const u32 (&b)[1], const Lane &L, Flags &F) { coop::add_reduce<Y, 1>(r, a, b, L, F); }
template <class Rc> inline void permute(u32 (&s)[3], const Rc &rct, const Lane &L, Flags &F) { coop::permute<Y>(s, rct, L, F); }
template <class Rc, class Ld> inline u32 sponge(u32 &o, u32 a, u32 b, int w, const Rc &rct, const Lane &L, Ld ld) { return coop::sponge<Y>(o, a, b, w, rct, L, ld); }
inline void mds_arc(u32 (&s)[3], const u32 (&rc)[3], bool h, const Lane &L, Flags &F) { coop::mds_arc<Y>(s, rc, h, L, F); }
inline bool flagged(const Flags &F) { return coop::flagged(F); }
inline u32 carry_exact(u32 lo, u32 c, const Lane &L) { return coop::carry_exact<Y>(lo, c, L); }
}  // namespace impl
#endif

namespace {
uint32_t g_rc[192][2];
struct RcTable {
  u32 operator()(int idx, int w) const { return g_rc[idx][w]; }
};

uint64_t splitmix(uint64_t &s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

const uint64_t P64[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};

// element generator: mixes uniform 256-bit values, canonical values, small values and values around multiples of p
int g_plain = 0;   // 1: only uniform / canonical / 64-bit inputs (to read the natural flag rate)
void gen_element(uint64_t &seed, uint64_t out[4], int kind) {
  for (int i = 0; i < 4; ++i) out[i] = splitmix(seed);
  if (g_plain) kind %= 3;
  switch (kind % 8) {
    case 0: break;                                            // any 256-bit value
    case 1: out[3] &= 0x0FFFFFFFFFFFFFFFULL; break;           // canonical-ish
    case 2: out[1] = out[2] = out[3] = 0; break;              // 64-bit (u64 leaves)
    case 3: out[2] = out[3] = 0; out[1] &= 3; break;          // ~66-bit: x^4 straddles 2^256
    case 4: {                                                 // M*p + small delta (M = 0..5)
      const uint64_t M = splitmix(seed) % 6;
      unsigned __int128 c = 0;
      const int64_t delta = (int64_t)(splitmix(seed) % 7) - 3;
      for (int i = 0; i < 4; ++i) {
        c += (unsigned __int128)P64[i] * M;
        out[i] = (uint64_t)c;
        c >>= 64;
      }
      if (M || delta >= 0) {
        unsigned __int128 t = (unsigned __int128)out[0] + (uint64_t)delta;
        if (delta < 0) {   // subtract |delta|
          uint64_t b = (uint64_t)(-delta);
          for (int i = 0; i < 4 && b; ++i) { uint64_t o = out[i]; out[i] = o - b; b = o < b ? 1 : 0; }
        } else {
          out[0] = (uint64_t)t;
          uint64_t cy = (uint64_t)(t >> 64);
          for (int i = 1; i < 4 && cy; ++i) { out[i] += cy; cy = out[i] == 0; }
        }
      }
      break;
    }
    case 5: out[0] = splitmix(seed) % 3; out[1] = out[2] = out[3] = 0; break;   // 0, 1, 2
    case 6: for (int i = 0; i < 4; ++i) out[i] = ~0ULL; out[0] -= splitmix(seed) % 4; break;   // 2^256 - small
    case 7: {                                                 // words of all ones / zeros mixed in (carry ripples)
      for (int i = 0; i < 4; ++i) {
        const uint64_t r = splitmix(seed) % 4;
        if (r == 0) out[i] = ~0ULL;
        else if (r == 1) out[i] = 0;
        else if (r == 2) out[i] |= 0xFFFFFFFF00000000ULL;
      }
      out[3] &= 0x3FFFFFFFFFFFFFFFULL;
      break;
    }
  }
}

struct Job {
  int kind;                       // 0 mul, 1 add, 2 permutation, 3 sponge, 4 mds layer, 5 exact carry resolution
  int width;
  std::vector<uint64_t> in;       // inputs (elements x 4)
  uint64_t out[12];
  uint32_t unc;
};
std::vector<Job> g_jobs;

inline u32 word_of(const uint64_t *el, u32 m) { return m < 8 ? (u32)(el[m >> 1] >> (32 * (m & 1))) : 0u; }
inline void put_word(uint64_t *el, u32 m, u32 v) {
  // called by all lanes into distinct 32-bit halves: go through a u32 view to avoid read-modify-write races
  if (m < 8) reinterpret_cast<volatile u32 *>(el)[m] = v;
}

void lane_main(uint32_t lane) {
  tl_lane = lane;
  const impl::Lane L = impl::make_lane();
  RcTable rct;
  for (auto &job : g_jobs) {
    impl::Flags F;
    u32 unc = 0;
    if (job.kind == 0) {
      u32 a[1][8], b[1], r[1];
      impl::gather(a[0], word_of(&job.in[0], lane));
      b[0] = word_of(&job.in[4], lane);
      impl::mulred1(r, a, b, L, F);
      put_word(job.out, lane, r[0]);
    } else if (job.kind == 1) {
      u32 a[1] = {word_of(&job.in[0], lane)}, b[1] = {word_of(&job.in[4], lane)}, r[1];
      impl::add_reduce1(r, a, b, L, F);
      put_word(job.out, lane, r[0]);
    } else if (job.kind == 2) {
      u32 s[3] = {word_of(&job.in[0], lane), word_of(&job.in[4], lane), word_of(&job.in[8], lane)};
      impl::permute(s, rct, L, F);
      for (int i = 0; i < 3; ++i) put_word(job.out + 4 * i, lane, s[i]);
    } else if (job.kind == 3) {
      u32 out;
      const uint64_t *in = job.in.data();
      const u32 vote = impl::sponge(out, 3u, 0u, job.width, rct, L, [&](int i) { return word_of(in + 4 * i, lane); });
      put_word(job.out, lane, out);
      unc = vote;
    } else if (job.kind == 5) {
      // lane values lo + 2^32 c: in[0..3] the low words, in[4..7] the carries
      put_word(job.out, lane, impl::carry_exact(word_of(&job.in[0], lane), word_of(&job.in[4], lane), L));
    } else {
      u32 s[3] = {word_of(&job.in[0], lane), word_of(&job.in[4], lane), word_of(&job.in[8], lane)};
      u32 rc[3] = {0, 0, 0};
      impl::mds_arc(s, rc, false, L, F);
      for (int i = 0; i < 3; ++i) put_word(job.out + 4 * i, lane, s[i]);
    }
    const u32 vote = coop::emul_ballot(unc != 0 || impl::flagged(F));
    if (lane == 0) job.unc = vote;
  }
}
}  // namespace

int main(int argc, char **argv) {
  const int units = argc > 1 ? atoi(argv[1]) : 300;
  uint64_t seed = argc > 2 ? strtoull(argv[2], nullptr, 0) : 1;
  g_plain = argc > 3 ? atoi(argv[3]) : 0;
  {
    std::vector<uint64_t> rc(192 * 4);
    cuzk_oracle_round_constants(rc.data());
    for (int i = 0; i < 192; ++i) {
      if (rc[4 * i + 1] | rc[4 * i + 2] | rc[4 * i + 3]) { printf("round constant above 2^64\n"); return 2; }
      g_rc[i][0] = (uint32_t)rc[4 * i];
      g_rc[i][1] = (uint32_t)(rc[4 * i] >> 32);
    }
  }
  // jobs: many cheap field operations, fewer permutations / sponges
  for (int u = 0; u < units * 40; ++u) {
    Job j;
    j.kind = u & 1;
    j.width = 0;
    j.in.resize(8);
    gen_element(seed, &j.in[0], (int)(splitmix(seed) % 8));
    gen_element(seed, &j.in[4], (int)(splitmix(seed) % 8));
    if ((u % 16) == 2) memcpy(&j.in[4], &j.in[0], 32);   // squares
    g_jobs.push_back(j);
  }
  for (int u = 0; u < units * 4; ++u) {
    Job j;
    j.kind = 4;
    j.width = 0;
    j.in.resize(12);
    for (int e = 0; e < 3; ++e) {
      do gen_element(seed, &j.in[4 * e], (int)(splitmix(seed) % 8));
      while (!(j.in[4 * e + 3] < P64[3]));   // canonical states only (strictly below p's top limb keeps it simple)
    }
    g_jobs.push_back(j);
  }
  for (int u = 0; u < units; ++u) {
    Job j;
    j.kind = 2 + (u & 1);
    j.width = (j.kind == 3) ? 1 + (int)(splitmix(seed) % 8) : 0;
    const int nel = (j.kind == 2) ? 3 : j.width;
    j.in.resize(4 * nel);
    for (int e = 0; e < nel; ++e) gen_element(seed, &j.in[4 * e], (int)(splitmix(seed) % 8));
    g_jobs.push_back(j);
  }
  // carry resolution: words of all ones (propagate), near all ones (generate with a carry), zero and random; carries up to 2^9
  for (int u = 0; u < units * 40; ++u) {
    Job j;
    j.kind = 5;
    j.width = 0;
    j.in.resize(8);
    u32 *w = reinterpret_cast<u32 *>(j.in.data());
    for (int g = 0; g < 8; ++g) {
      const uint64_t r = splitmix(seed);
      const u32 pick = (u32)(r % 6);
      w[g] = pick <= 1 ? 0xFFFFFFFFu : (pick == 2 ? 0xFFFFFFFFu - (u32)((r >> 8) % 600) : (pick == 3 ? 0u : (u32)(r >> 16)));
      w[8 + g] = (u32)((r >> 48) % ((r & 0x80) ? 3u : 513u));
    }
    g_jobs.push_back(j);
  }
  std::vector<std::thread> th;
  for (uint32_t l = 0; l < (uint32_t)kLanes; ++l) th.emplace_back(lane_main, l);
  for (auto &t : th) t.join();

  long checked[6] = {0}, flagged[6] = {0}, bad[6] = {0}, bad_flagged[6] = {0};
  for (auto &job : g_jobs) {
    uint64_t want[12];
    int nout = 4;
    if (job.kind == 0) cuzk_oracle_fr_mul(&job.in[0], &job.in[4], want);
    else if (job.kind == 1) cuzk_oracle_fr_add(&job.in[0], &job.in[4], want);
    else if (job.kind == 2) { memcpy(want, job.in.data(), 96); cuzk_oracle_permutation(want); nout = 12; }
    else if (job.kind == 3) cuzk_oracle_batch_sponge(job.in.data(), (size_t)job.width, 3, want, 1);
    else if (job.kind == 5) {   // sum_g (lo_g + 2^32 c_g) 2^(32 g) mod 2^256
      const u32 *w = reinterpret_cast<const u32 *>(job.in.data());
      u32 *o = reinterpret_cast<u32 *>(want);
      uint64_t carry = 0;
      for (int g = 0; g < 8; ++g) {
        const uint64_t t = (uint64_t)w[g] + (g ? (uint64_t)w[8 + g - 1] : 0) + carry;
        o[g] = (u32)t;
        carry = t >> 32;
      }
    }
    else { memcpy(want, job.in.data(), 96); cuzk_oracle_batch_mds_layer(want, 1); nout = 12; }
    const bool same = memcmp(want, job.out, 8 * nout) == 0;
    ++checked[job.kind];
    if (job.unc) {
      ++flagged[job.kind];
      if (!same) ++bad_flagged[job.kind];
    } else if (!same) {
      if (bad[job.kind]++ < 3) {
        printf("MISMATCH kind %d width %d\n  in :", job.kind, job.width);
        for (size_t i = 0; i < job.in.size() && i < 12; ++i) printf(" %016llx", (unsigned long long)job.in[i]);
        printf("\n  got:");
        for (int i = 0; i < nout; ++i) printf(" %016llx", (unsigned long long)job.out[i]);
        printf("\n  exp:");
        for (int i = 0; i < nout; ++i) printf(" %016llx", (unsigned long long)want[i]);
        printf("\n");
      }
    }
  }
  const char *names[6] = {"multiply", "add", "permutation", "sponge", "mds_layer", "carry_exact"};
  long total_bad = 0;
  for (int k = 0; k < 6; ++k) {
    printf("{\"op\": \"%s\", \"checked\": %ld, \"flagged\": %ld, \"flagged_and_different\": %ld, \"unflagged_mismatches\": %ld}\n", names[k],
           checked[k], flagged[k], bad_flagged[k], bad[k]);
    total_bad += bad[k];
  }
  return total_bad ? 1 : 0;
}
