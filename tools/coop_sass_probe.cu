// coop_sass_probe.cu -- compile-only helper: isolates the building blocks of the cooperative path in tiny kernels so their
// SASS can be counted (tools/coop_sass_count.sh).  Not part of the library.
#include <cuda_runtime.h>
#include "../cuzk_b200/csrc/poseidon.cuh"
using namespace cuzk;
#include "../cuzk_b200/csrc/coop.cuh"
#ifndef PROBE_LAYOUT
#define PROBE_LAYOUT Wide16
#endif
typedef coop::PROBE_LAYOUT Y;

__global__ void probe_mulred1(u32 *io) {
  const coop::Lane L = Y::make_lane();
  coop::Flags F;
  u32 a[1][8], b[1] = {io[threadIdx.x]}, r[1];
  coop::gather<Y>(a[0], io[32 + threadIdx.x]);
  coop::mulred<Y, 1>(r, a, b, L, F);
  io[threadIdx.x] = r[0] + F.ovf + F.near;
}
__global__ void probe_gather(u32 *io) {
  u32 a[8];
  coop::gather<Y>(a, io[threadIdx.x]);
  u32 s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= a[i] * (i + 3);
  io[threadIdx.x] = s;
}
__global__ void probe_sbox1(u32 *io) {
  const coop::Lane L = Y::make_lane();
  coop::Flags F;
  u32 x[1] = {io[threadIdx.x]};
  coop::sbox<Y, 1>(x, L, F);
  io[threadIdx.x] = x[0] + F.ovf + F.near;
}
__global__ void probe_sbox3(u32 *io) {
  const coop::Lane L = Y::make_lane();
  coop::Flags F;
  u32 x[3] = {io[threadIdx.x], io[32 + threadIdx.x], io[64 + threadIdx.x]};
  coop::sbox<Y, 3>(x, L, F);
  io[threadIdx.x] = x[0] + x[1] + x[2] + F.ovf + F.near;
}
__global__ void probe_mds(u32 *io) {
  const coop::Lane L = Y::make_lane();
  coop::Flags F;
  u32 s[3] = {io[threadIdx.x], io[32 + threadIdx.x], io[64 + threadIdx.x]};
  u32 rc[3] = {io[96 + threadIdx.x], io[128 + threadIdx.x], io[160 + threadIdx.x]};
  coop::mds_arc<Y>(s, rc, true, L, F);
  io[threadIdx.x] = s[0] + s[1] + s[2] + F.ovf + F.near;
}
