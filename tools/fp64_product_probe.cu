// Measured answer to "put the S-box products on the FP64 pipe": the 256 x 256 -> 512-bit product, in isolation, on both pipes.
//
//   int   : mul_wide_8x8 of fr.cuh (the product the kernels use: 64 IMAD.WIDE.U32 in even/odd carry lanes, 8 x u32 words)
//   fp64  : 5 x 52-bit limbs held in doubles; per partial product  hi = fma_rz(a, b, 2^104),  lo = fma_rz(a, b, (2^104 + 2^52) - hi)
//           (both exact: hi = 2^104 + floor(ab / 2^52) 2^52, lo = 2^52 + ab mod 2^52), the raw bit patterns added as 64-bit integers
//           column by column (the exponent fields cancel against a constant), then one carry pass to 52-bit limbs
//   fp64 + conv : the same, plus turning the ten result limbs back into doubles (what the next product needs)
//
// Every thread runs a chain of dependent products (each product's operands come from the previous result), `inner` independent
// chains per thread for ILP.  The first product of every thread is also compared between the two forms word for word.
//
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I cuzk_b200/csrc tools/fp64_product_probe.cu -o tools/fp64_product_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fr.cuh"

using cuzk::u32;
using cuzk::u64;

namespace {

constexpr double kBiasHi = 20282409603651670423947251286016.0;      // 2^104
constexpr double kBiasSub = 20282409603651674927546878656512.0;     // 2^104 + 2^52
constexpr u64 kHiBits = 0x4670000000000000ull;                      // bit pattern of 2^104
constexpr u64 kLoBits = 0x4330000000000000ull;                      // bit pattern of 2^52
constexpr u64 kMask52 = (1ull << 52) - 1;

__device__ __forceinline__ void to_limbs(u64 (&l)[5], const u32 (&w)[8]) {
  const u64 x0 = (u64)w[0] | ((u64)w[1] << 32), x1 = (u64)w[2] | ((u64)w[3] << 32), x2 = (u64)w[4] | ((u64)w[5] << 32),
            x3 = (u64)w[6] | ((u64)w[7] << 32);
  l[0] = x0 & kMask52;
  l[1] = ((x0 >> 52) | (x1 << 12)) & kMask52;
  l[2] = ((x1 >> 40) | (x2 << 24)) & kMask52;
  l[3] = ((x2 >> 28) | (x3 << 36)) & kMask52;
  l[4] = x3 >> 16;
}
// integer limb (< 2^52) -> double, through the 2^52 bias (one logic op on the high word and one DADD)
__device__ __forceinline__ double limb_to_double(u64 l) { return __longlong_as_double((long long)(l | kLoBits)) - 4503599627370496.0; }

// 5 x 5 limbs -> ten normalised 52-bit limbs
__device__ __forceinline__ void mul_fp64(u64 (&r)[10], const double (&a)[5], const double (&b)[5]) {
  u64 col[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    // minus the exponent fields this column will collect: lo parts of the products with i + j == k, hi parts of i + j == k - 1
    const int nlo = k < 5 ? k + 1 : (k < 9 ? 9 - k : 0);
    const int nhi = k >= 1 ? (k - 1 < 5 ? k : 10 - k) : 0;
    col[k] = 0ull - (u64)nlo * kLoBits - (u64)nhi * kHiBits;
  }
#pragma unroll
  for (int i = 0; i < 5; ++i) {
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const double hi = __fma_rz(a[i], b[j], kBiasHi);
      const double lo = __fma_rz(a[i], b[j], kBiasSub - hi);
      col[i + j] += (u64)__double_as_longlong(lo);
      col[i + j + 1] += (u64)__double_as_longlong(hi);
    }
  }
  u64 carry = 0;
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    const u64 v = col[k] + carry;
    r[k] = v & kMask52;
    carry = v >> 52;
  }
}

__device__ __forceinline__ void limbs_to_words(u32 (&w)[16], const u64 (&l)[10]) {
  // 520 bits of limbs -> 512 bits of words (the top 8 bits are zero for a 512-bit product)
  u64 x[8];
  x[0] = l[0] | (l[1] << 52);
  x[1] = (l[1] >> 12) | (l[2] << 40);
  x[2] = (l[2] >> 24) | (l[3] << 28);
  x[3] = (l[3] >> 36) | (l[4] << 16);
  x[4] = (l[4] >> 48) | (l[5] << 4) | (l[6] << 56);
  x[5] = (l[6] >> 8) | (l[7] << 44);
  x[6] = (l[7] >> 20) | (l[8] << 32);
  x[7] = (l[8] >> 32) | (l[9] << 20);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    w[2 * i] = (u32)x[i];
    w[2 * i + 1] = (u32)(x[i] >> 32);
  }
}

template <int INNER>
__global__ void __launch_bounds__(256) int_chain_kernel(const uint4 *in, uint4 *out, int steps) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  u32 a[INNER][8], b[8];
  {
    const uint4 lo = in[2 * t], hi = in[2 * t + 1];
    b[0] = lo.x; b[1] = lo.y; b[2] = lo.z; b[3] = lo.w; b[4] = hi.x; b[5] = hi.y; b[6] = hi.z; b[7] = hi.w;
  }
#pragma unroll
  for (int c = 0; c < INNER; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[c][i] = b[i] + (u32)(c * 0x9e3779b9u) + i;
  for (int s = 0; s < steps; ++s) {
#pragma unroll
    for (int c = 0; c < INNER; ++c) {
      u32 r[16];
      cuzk::mul_wide_8x8(r, a[c], b);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[c][i] = r[i] ^ r[i + 8];
    }
  }
  u32 acc[8] = {};
#pragma unroll
  for (int c = 0; c < INNER; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] ^= a[c][i];
  out[2 * t] = make_uint4(acc[0], acc[1], acc[2], acc[3]);
  out[2 * t + 1] = make_uint4(acc[4], acc[5], acc[6], acc[7]);
}

// CONV: 0 = the next operand is formed on the limbs in integer registers and re-biased (the cheapest possible feed), 1 = as 0 plus
// nothing else; the conversion is always there because a product's result is the next product's operand
template <int INNER>
__global__ void __launch_bounds__(256) fp_chain_kernel(const uint4 *in, uint4 *out, int steps) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  u32 w[8];
  {
    const uint4 lo = in[2 * t], hi = in[2 * t + 1];
    w[0] = lo.x; w[1] = lo.y; w[2] = lo.z; w[3] = lo.w; w[4] = hi.x; w[5] = hi.y; w[6] = hi.z; w[7] = hi.w;
  }
  u64 bl[5];
  to_limbs(bl, w);
  double b[5], a[INNER][5];
#pragma unroll
  for (int i = 0; i < 5; ++i) b[i] = limb_to_double(bl[i]);
#pragma unroll
  for (int c = 0; c < INNER; ++c)
#pragma unroll
    for (int i = 0; i < 5; ++i) a[c][i] = limb_to_double((bl[i] + (u64)(c * 0x9e3779b9u) + i) & kMask52);
  u64 keep[INNER][5];
  for (int s = 0; s < steps; ++s) {
#pragma unroll
    for (int c = 0; c < INNER; ++c) {
      u64 r[10];
      mul_fp64(r, a[c], b);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        keep[c][i] = r[i] ^ r[i + 5];
        a[c][i] = limb_to_double(keep[c][i]);
      }
    }
  }
  u64 acc[4] = {};
#pragma unroll
  for (int c = 0; c < INNER; ++c)
#pragma unroll
    for (int i = 0; i < 5; ++i) acc[i & 3] ^= keep[c][i];
  out[2 * t] = make_uint4((u32)acc[0], (u32)(acc[0] >> 32), (u32)acc[1], (u32)(acc[1] >> 32));
  out[2 * t + 1] = make_uint4((u32)acc[2], (u32)(acc[2] >> 32), (u32)acc[3], (u32)(acc[3] >> 32));
}

// one product per thread in both forms, all sixteen words compared
__global__ void compare_kernel(const uint4 *in, size_t n, unsigned long long *mismatches) {
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t + 1 >= n) return;
  u32 a[8], b[8];
  {
    const uint4 lo = in[2 * t], hi = in[2 * t + 1], lo2 = in[2 * t + 2], hi2 = in[2 * t + 3];
    a[0] = lo.x; a[1] = lo.y; a[2] = lo.z; a[3] = lo.w; a[4] = hi.x; a[5] = hi.y; a[6] = hi.z; a[7] = hi.w;
    b[0] = lo2.x; b[1] = lo2.y; b[2] = lo2.z; b[3] = lo2.w; b[4] = hi2.x; b[5] = hi2.y; b[6] = hi2.z; b[7] = hi2.w;
  }
  u32 ri[16], rf[16];
  cuzk::mul_wide_8x8(ri, a, b);
  u64 al[5], bl[5], r[10];
  to_limbs(al, a);
  to_limbs(bl, b);
  double ad[5], bd[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    ad[i] = limb_to_double(al[i]);
    bd[i] = limb_to_double(bl[i]);
  }
  mul_fp64(r, ad, bd);
  limbs_to_words(rf, r);
  bool bad = false;
#pragma unroll
  for (int i = 0; i < 16; ++i) bad |= ri[i] != rf[i];
  if (bad) atomicAdd(mismatches, 1ull);
}

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      fprintf(stderr, "%s: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__);     \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

template <typename K>
int time_kernel(K kernel, const char *name, int inner, const uint4 *in, uint4 *out, int blocks, int steps, double clock_ghz) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0));
    kernel<<<blocks, 256>>>(in, out, steps);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  const double products = (double)blocks * 256 * inner * steps;
  const double per_s = products / (best * 1e-3);
  // cycles one SM sub-partition spends per warp-wide product
  const double cyc = clock_ghz * 1e9 * 148 * 4 / (per_s / 32);
  printf("{\"kernel\": \"%s\", \"inner\": %d, \"blocks\": %d, \"steps\": %d, \"ms\": %.3f, \"products_per_s\": %.4g, \"cycles_per_warp_product_per_subpartition\": %.1f}\n",
         name, inner, blocks, steps, best, per_s, cyc);
  return 0;
}

}  // namespace

int main(int argc, char **argv) {
  const int steps = argc > 1 ? atoi(argv[1]) : 2000;
  const int blocks = 148 * 6;   // 6 x 256 threads per SM: full residency for both kernels if registers allow
  const size_t n = (size_t)blocks * 256;
  std::vector<uint32_t> h(n * 8);
  uint64_t s = 0x243f6a8885a308d3ull;
  for (auto &v : h) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    v = (uint32_t)(s >> 32);
  }
  // a few extreme operands for the comparison: all ones, zero, single bits
  for (int i = 0; i < 8; ++i) { h[i] = 0xffffffffu; h[8 + i] = 0xffffffffu; h[16 + i] = 0; h[24 + i] = i == 7 ? 0x80000000u : 0; h[32 + i] = 0xffffffffu; }
  uint4 *d_in, *d_out;
  unsigned long long *d_bad, bad = 0;
  CK(cudaMalloc(&d_in, n * 32));
  CK(cudaMalloc(&d_out, n * 32));
  CK(cudaMalloc(&d_bad, 8));
  CK(cudaMemcpy(d_in, h.data(), n * 32, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_bad, 0, 8));
  compare_kernel<<<blocks, 256>>>(d_in, n, d_bad);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost));
  printf("{\"compare\": \"int vs fp64 product, %zu operand pairs\", \"mismatches\": %llu}\n", n - 1, bad);
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  const double ghz = clk_khz * 1e-6;
  if (time_kernel(int_chain_kernel<1>, "int_imad_wide", 1, d_in, d_out, blocks, steps, ghz)) return 1;
  if (time_kernel(int_chain_kernel<2>, "int_imad_wide", 2, d_in, d_out, blocks, steps, ghz)) return 1;
  if (time_kernel(fp_chain_kernel<1>, "fp64_dfma", 1, d_in, d_out, blocks, steps, ghz)) return 1;
  if (time_kernel(fp_chain_kernel<2>, "fp64_dfma", 2, d_in, d_out, blocks, steps, ghz)) return 1;
  return bad ? 2 : 0;
}
