// On-chip rooflines the DESIGN.md analysis of the fused similarity kernel leans on, measured on the box:
//   1. tcgen05.ld (TMEM -> registers) throughput per SM, with 4 and 8 reading warps and with the
//      32x32b.x32 / .x16 shapes  -> the floor of an epilogue that has to look at every accumulator
//      (the attention row f-1: 128 rows x 1203 columns per anchor tile for K = 32 of MMA work).
//   2. tcgen05.st (registers -> TMEM) throughput (the converters' publish step).
//   3. ld.shared throughput for the converters' access pattern (32 lanes x 4 bytes, one k row per
//      instruction) with 4 warps                            -> 256 cycles per 32 KiB fp32 block.
// Build + run (one GPU):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/onchip_bench tools/onchip_bench.cu && /tmp/onchip_bench
// Prints one JSON line.  A measurement tool; nothing of the product includes it.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),
        "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// mode 0: ld x32, 1: ld x16, 2: st x32, 3: ld.shared (converter pattern), 4: ld x32, two in flight
template <int MODE>
__global__ void __launch_bounds__(256, 1) bench_kernel(int iters, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_ptr;
  extern __shared__ float dyn[];
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (MODE == 3)
    for (int i = threadIdx.x; i < 2 * 64 * 128; i += blockDim.x) dyn[i] = (float)i;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {               // 8 x 32 columns: this warp's half of the 512 columns
        tmem_ld32(base + (uint32_t)((warp >> 2) * 256 + c * 32), r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += __uint_as_float(r[0] ^ r[31]);
      }
    } else if (MODE == 1) {
      uint32_t q[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        tmem_ld16(base + (uint32_t)((warp >> 2) * 256 + c * 16), q);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += __uint_as_float(q[0] ^ q[15]);
      }
    } else if (MODE == 4) {                        // two loads in flight per warp
      uint32_t r2[32];
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        tmem_ld32(base + (uint32_t)((warp >> 2) * 256 + c * 32), r);
        tmem_ld32(base + (uint32_t)((warp >> 2) * 256 + c * 32 + 32), r2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += __uint_as_float(r[0] ^ r2[31]);
      }
    } else if (MODE == 2) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        tmem_st32(base + (uint32_t)((warp >> 2) * 256 + c * 32), r);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
    } else {
      // the converters' pattern: thread = anchor (128 anchors = warps 0-3; warps 4-7 repeat it), one k row
      // of 128 floats per instruction, 64 rows = one 32 KiB block
      const volatile float* col = dyn + (threadIdx.x & 127) + ((it & 1) << 13);
#pragma unroll
      for (int k = 0; k < 64; ++k) acc += col[k * 128];
    }
  }
  const long long t1 = clock64();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_ptr), "r"(512u) : "memory");
}

template <int MODE>
int run(int warps, int iters, double* bytes_per_clk) {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* cyc;
  float* sink;
  CK(cudaMalloc(&cyc, sms * sizeof(long long)));
  CK(cudaMalloc(&sink, 4));
  CK(cudaFuncSetAttribute(bench_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 64 * 128 * 4));
  for (int rep = 0; rep < 2; ++rep) {
    bench_kernel<MODE><<<sms, warps * 32, 2 * 64 * 128 * 4>>>(iters, cyc, sink);
    CK(cudaDeviceSynchronize());
  }
  long long h[256];
  CK(cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
  double mean = 0;
  for (int i = 0; i < sms; ++i) mean += (double)h[i];
  mean /= sms;
  // bytes moved per CTA per iteration: every warp touches 32 lanes x 256 columns x 4 bytes (TMEM modes)
  // or 64 rows x 128 bytes (shared mode)
  const double per_warp = MODE == 3 ? 64.0 * 128.0 : 32.0 * 256.0 * 4.0;
  *bytes_per_clk = per_warp * warps * iters / mean;
  cudaFree(cyc);
  cudaFree(sink);
  return 0;
}

int main() {
  double ld32x2_4, ld32x2_8, ld32_4, ld32_8, ld16_4, ld16_8, st32_4, st32_8, lds_4, lds_8;
  const int iters = 2000;
  if (run<0>(4, iters, &ld32_4) || run<0>(8, iters, &ld32_8) || run<1>(4, iters, &ld16_4) || run<1>(8, iters, &ld16_8) ||
      run<2>(4, iters, &st32_4) || run<2>(8, iters, &st32_8) || run<3>(4, iters, &lds_4) || run<3>(8, iters, &lds_8) || run<4>(4, iters, &ld32x2_4) || run<4>(8, iters, &ld32x2_8))
    return 1;
  printf("{\"what\": \"on-chip throughput per SM, bytes per SM clock, all SMs busy; each tcgen05.ld/st is followed by its wait "
         "(one instruction in flight per warp, as in the epilogue's per-chunk drain)\", "
         "\"tcgen05_ld_32x32b_x32\": {\"warps4\": %.1f, \"warps8\": %.1f}, "
         "\"tcgen05_ld_32x32b_x32_two_in_flight\": {\"warps4\": %.1f, \"warps8\": %.1f}, "
         "\"tcgen05_ld_32x32b_x16\": {\"warps4\": %.1f, \"warps8\": %.1f}, "
         "\"tcgen05_st_32x32b_x32\": {\"warps4\": %.1f, \"warps8\": %.1f}, "
         "\"ld_shared_converter_pattern\": {\"warps4\": %.1f, \"warps8\": %.1f}}\n",
         ld32_4, ld32_8, ld32x2_4, ld32x2_8, ld16_4, ld16_8, st32_4, st32_8, lds_4, lds_8);
  return 0;
}
