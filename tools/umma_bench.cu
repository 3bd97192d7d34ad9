// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16) as a function of N, A source (TMEM / SMEM) and the
// number of distinct accumulator regions the issue stream round-robins over (dependent-accumulate distance).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench tools/umma_bench.cu && ./umma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../nerfmeshes_b200/csrc/nm_ptx.cuh"

using namespace nm::ptx;

// warp-converged issue: every lane runs the loop, one elected lane issues (elect.sync inside the asm statement)
__device__ __forceinline__ void mma_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred pe, p;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_ss_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred pe, p;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) bench_warp(int N, int a_tmem, int regions, int count, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase + 200 * 1024, tptr = bar + 16;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 200 * 1024 + 16);
  if (threadIdx.x < 32) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint64_t adesc = make_kmajor_sw128_desc(sbase);
    const uint64_t bdesc = make_kmajor_sw128_desc(sbase + 32768);
    long long t0 = clock64();
    int reg = 0;
    for (int i = 0; i < count; i += 4) {
      const uint32_t d = tmem + (uint32_t)(reg * N);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        if (a_tmem) mma_ts_elect(d, tmem + 384 + 8 * ks, bdesc + 2 * ks, idesc, 1);
        else mma_ss_elect(d, adesc + 2 * ks, bdesc + 2 * ks, idesc, 1);
      }
      if (++reg == regions) reg = 0;
    }
    long long t1 = clock64();
    commit_elect(bar);
    while (!mbar_try_wait(bar, 0)) {}
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// several issuing warps, each accumulating into its own 64-column region: is the ~100-cycle issue cost per warp
// (parallelisable) or a shared dispatch limit?
__global__ void __launch_bounds__(128, 1) bench_multi(int N, int a_tmem, int nwarps, int count, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase + 200 * 1024, tptr = bar + 64;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { for (int w = 0; w < 4; ++w) mbar_init(bar + 8 * w, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 200 * 1024 + 64);
  const int w = threadIdx.x >> 5;
  if (w < nwarps) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint64_t adesc = make_kmajor_sw128_desc(sbase);
    const uint64_t bdesc = make_kmajor_sw128_desc(sbase + 32768 + w * 8192);
    const uint32_t d = tmem + (uint32_t)(w * N);
    long long t0 = clock64();
    for (int i = 0; i < count; i += 4) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        if (a_tmem) mma_ts_elect(d, tmem + 384 + 8 * ks, bdesc + 2 * ks, idesc, 1);
        else mma_ss_elect(d, adesc + 2 * ks, bdesc + 2 * ks, idesc, 1);
      }
    }
    long long t1 = clock64();
    commit_elect(bar + 8 * w);
    while (!mbar_try_wait(bar + 8 * w, 0)) {}
    long long t2 = clock64();
    if ((threadIdx.x & 31) == 0) { out[2 * w] = t1 - t0; out[2 * w + 1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

__global__ void __launch_bounds__(128, 1) bench(int N, int a_tmem, int regions, int count, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase + 200 * 1024, tptr = bar + 16;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) { tmem_alloc(tptr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 200 * 1024 + 16);
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint64_t adesc = make_kmajor_sw128_desc(sbase);             // 128 x 64 A tile
    const uint64_t bdesc = make_kmajor_sw128_desc(sbase + 32768);      // N x 64 B tile (<= 32 KB)
    long long t0 = clock64();
    for (int i = 0; i < count; ++i) {
      const uint32_t d = tmem + (uint32_t)((i % regions) * N);
      const uint32_t ks = (uint32_t)(i & 3);
      if (a_tmem) mma_ts(d, tmem + 384 + 8 * ks, bdesc + 2 * ks, idesc, 1);
      else mma_ss(d, adesc + 2 * ks, bdesc + 2 * ks, idesc, 1);
    }
    long long t1 = clock64();
    tc_commit(bar);
    while (!mbar_try_wait(bar, 0)) {}
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  const int smem = 200 * 1024 + 64;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int count = 512;
  cudaFuncSetAttribute(bench_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  printf("%5s %6s %6s %8s %12s %12s\n", "style", "N", "A", "regions", "issue/mma", "total/mma");
  for (int style = 0; style < 2; ++style)
  for (int a_tmem = 1; a_tmem >= 0; --a_tmem)
    for (int N : {64, 128, 256})
      for (int regions : {1, 2, 3, 4}) {
        if (regions * N > 384 - (a_tmem ? 0 : 0)) continue;
        for (int rep = 0; rep < 2; ++rep) {
          if (style) bench_warp<<<1, 128, smem>>>(N, a_tmem, regions, count, d);
          else bench<<<1, 128, smem>>>(N, a_tmem, regions, count, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%5s %6d %6s %8d %12.1f %12.1f\n", style ? "warp" : "lane0", N, a_tmem ? "tmem" : "smem", regions, (double)h[0] / count, (double)h[1] / count);
      }
  cudaFuncSetAttribute(bench_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  printf("multi-warp issue, N=64 (per-warp region): nwarps, A, per-warp issue cycles/mma, aggregate cycles/mma\n");
  for (int a_tmem = 1; a_tmem >= 0; --a_tmem)
    for (int nw = 1; nw <= 4; ++nw) {
      for (int rep = 0; rep < 2; ++rep) { bench_multi<<<1, 128, smem>>>(64, a_tmem, nw, count, d); cudaDeviceSynchronize(); }
      long long h[8];
      cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int w = 0; w < nw; ++w) mx = h[2 * w + 1] > mx ? h[2 * w + 1] : mx;
      printf("%d %s %.1f %.1f\n", nw, a_tmem ? "tmem" : "smem", (double)h[0] / count, (double)mx / (count * nw));
    }
  return 0;
}
