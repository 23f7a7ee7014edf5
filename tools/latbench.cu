// latbench.cu -- single-warp latency / issue-rate probes for the primitives of the cooperative (8 lanes per permutation)
// Poseidon path: SHFL, IMAD.WIDE chains, carry chains, VOTE, shared-memory round trips.  One CTA of `warps` warps per SM
// so the numbers can be read with one warp per SM (no contention) and with four (one per sub-partition).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/latbench tools/latbench.cu ; run: tools/latbench
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

typedef uint32_t u32;
typedef uint64_t u64;

constexpr int ITERS = 512;

template <int TEST>
__global__ void probe(u64 *out, u32 seed) {
  __shared__ u32 smem[1024];
  const u32 lane = threadIdx.x & 31;
  u32 x = seed + threadIdx.x * 2654435761u, y = seed ^ (threadIdx.x * 40503u), z = x ^ y;
  u32 r[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) r[i] = x + i * y;
  smem[threadIdx.x] = x;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    if (TEST == 0) {          // dependent SHFL.IDX chain, width 8, lane-varying source
#pragma unroll
      for (int k = 0; k < 8; ++k) x = __shfl_sync(0xffffffffu, x, (lane + k + x) & 7, 8);
    } else if (TEST == 1) {   // 8 independent SHFL + dependent add (one "round")
      u32 s = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += __shfl_sync(0xffffffffu, r[k], (lane - k) & 7, 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) r[k] += s;
    } else if (TEST == 2) {   // dependent IMAD.WIDE chain: t = a*b + t (64-bit accumulator)
      u64 t = ((u64)x << 32) | y;
#pragma unroll
      for (int k = 0; k < 8; ++k) t = (u64)(u32)t * (u64)z + t;
      x = (u32)t; y = (u32)(t >> 32);
    } else if (TEST == 3) {   // row product chain: t = a_k*b + (t >> 32)
      u64 t = y;
#pragma unroll
      for (int k = 0; k < 8; ++k) { t = (u64)r[k] * (u64)z + (t >> 32); r[k] = (u32)t; }
      z ^= (u32)(t >> 32);
    } else if (TEST == 4) {   // carry chain of 8 addc
      asm volatile(
          "add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %0;\n\taddc.cc.u32 %2, %2, %1;\n\taddc.cc.u32 %3, %3, %2;\n\t"
          "addc.cc.u32 %4, %4, %3;\n\taddc.cc.u32 %5, %5, %4;\n\taddc.cc.u32 %6, %6, %5;\n\taddc.u32 %7, %7, %6;"
          : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
          : "r"(r[8]));
      r[8] = r[7];
    } else if (TEST == 5) {   // dependent ballot chain
#pragma unroll
      for (int k = 0; k < 8; ++k) x = __ballot_sync(0xffffffffu, (x >> (lane & 7)) & 1u) + k + lane;
    } else if (TEST == 6) {   // shared memory round trip: STS then LDS from another lane's slot
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        smem[threadIdx.x] = x;
        __syncwarp();
        x = smem[(threadIdx.x & ~7u) | ((lane + 1 + x) & 7)] + k;
        __syncwarp();
      }
    } else if (TEST == 7) {   // dependent 32-bit IADD chain (ALU latency)
#pragma unroll
      for (int k = 0; k < 8; ++k) { x = x + (y ^ x); }
    } else if (TEST == 8) {   // dependent IMAD (32-bit) chain
#pragma unroll
      for (int k = 0; k < 8; ++k) { x = x * y + z; }
    } else if (TEST == 9) {   // dependent 64-bit add chain (IADD3 + IADD3.X)
      u64 t = ((u64)x << 32) | y;
#pragma unroll
      for (int k = 0; k < 8; ++k) t = t + (t >> 7) + z;
      x = (u32)t; y = (u32)(t >> 32);
    } else if (TEST == 10) {  // 9 independent SHFL then 64-bit accumulate of 9 words (transposed-sum shape)
      u64 t = 0;
#pragma unroll
      for (int k = 0; k < 9; ++k) t += __shfl_sync(0xffffffffu, r[k], (lane - k) & 7, 8);
#pragma unroll
      for (int k = 0; k < 9; ++k) r[k] ^= (u32)t + (u32)(t >> 32);
    } else if (TEST == 11) {  // independent IMAD.WIDE issue rate: 8 independent accumulators
      u64 t[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) t[k] = ((u64)r[k] << 32) | x;
#pragma unroll
      for (int rep = 0; rep < 4; ++rep)
#pragma unroll
        for (int k = 0; k < 8; ++k) t[k] = (u64)(u32)t[k] * (u64)z + t[k];
#pragma unroll
      for (int k = 0; k < 8; ++k) r[k] = (u32)t[k] ^ (u32)(t[k] >> 32);
    } else if (TEST == 12) {  // independent SHFL issue rate: 32 shuffles, no dependence between them
#pragma unroll
      for (int rep = 0; rep < 4; ++rep)
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = __shfl_sync(0xffffffffu, r[k], (lane + k + 1) & 7, 8);
    } else if (TEST == 13) {  // SHFL -> IMAD.WIDE -> SHFL dependent (cross-pipe)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        x = __shfl_sync(0xffffffffu, x, (lane + 1) & 7, 8);
        u64 t = (u64)x * (u64)z + y;
        x = (u32)t ^ (u32)(t >> 32);
      }
    } else if (TEST == 14) {  // dependent match-free vote.any (predicate result)
#pragma unroll
      for (int k = 0; k < 8; ++k) x += __any_sync(0xffffffffu, (x & 3u) == 1u) ? 3u : 5u;
    }
  }
  long long t1 = clock64();
  u32 acc = x ^ y ^ z;
#pragma unroll
  for (int i = 0; i < 9; ++i) acc ^= r[i];
  if (threadIdx.x == 0) out[blockIdx.x] = (u64)(t1 - t0);
  if (acc == 0x1234567u) out[gridDim.x] = acc;
}

template <int TEST>
void run(const char *name, int ops_per_iter, int warps, u64 *d_out, int blocks) {
  u64 h[2048];
  probe<TEST><<<blocks, 32 * warps>>>(d_out, 12345u);
  probe<TEST><<<blocks, 32 * warps>>>(d_out, 54321u);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  cudaMemcpy(h, d_out, sizeof(u64) * blocks, cudaMemcpyDeviceToHost);
  double best = 1e30;
  for (int i = 0; i < blocks; ++i) if ((double)h[i] < best) best = (double)h[i];
  printf("{\"test\": \"%s\", \"warps_per_sm\": %d, \"cycles_per_iter\": %.1f, \"cycles_per_op\": %.2f}\n", name, warps,
         best / ITERS, best / ITERS / ops_per_iter);
}

int main() {
  u64 *d_out;
  cudaMalloc(&d_out, sizeof(u64) * 2049);
  const int blocks = 148;
  for (int warps : {1, 4, 8}) {
    run<0>("shfl_dependent_x8", 8, warps, d_out, blocks);
    run<1>("shfl_round_8indep_plus_add", 1, warps, d_out, blocks);
    run<2>("imad_wide_dependent_acc_x8", 8, warps, d_out, blocks);
    run<3>("imad_wide_row_chain_x8", 8, warps, d_out, blocks);
    run<4>("addc_chain_x8", 8, warps, d_out, blocks);
    run<5>("ballot_dependent_x8", 8, warps, d_out, blocks);
    run<6>("sts_lds_roundtrip_x8", 8, warps, d_out, blocks);
    run<7>("iadd_dependent_x8(2 ops each)", 8, warps, d_out, blocks);
    run<8>("imad32_dependent_x8", 8, warps, d_out, blocks);
    run<9>("add64_dependent_x8", 8, warps, d_out, blocks);
    run<10>("transposed_sum_9shfl_acc64", 1, warps, d_out, blocks);
    run<11>("imad_wide_independent_x32", 32, warps, d_out, blocks);
    run<12>("shfl_independent_x32", 32, warps, d_out, blocks);
    run<13>("shfl_imadwide_pingpong_x4", 4, warps, d_out, blocks);
    run<14>("vote_any_dependent_x8", 8, warps, d_out, blocks);
  }
  cudaFree(d_out);
  return 0;
}
