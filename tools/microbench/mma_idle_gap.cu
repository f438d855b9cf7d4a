// Microbenchmark: issue time of a warm train of 8 tcgen05.mma (M=64, N=128, K=16) after the tensor pipe has been idle
// for G cycles, and the same with one tiny "keep-warm" MMA (N=8, into spare TMEM columns) issued D cycles before the train.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "smz_tc_ptx.cuh"
using namespace smz_tc;

__device__ __forceinline__ void mma(unsigned d, unsigned long long ad, unsigned long long bd, unsigned idesc, unsigned acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) k(int gap, int warm_before, long long* out) {
  extern __shared__ unsigned char raw[];
  unsigned char* sm = (unsigned char*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  __shared__ unsigned long long bar;
  __shared__ unsigned tmem_base;
  constexpr int M = 64, N = 128;
  unsigned char* A = sm;
  unsigned char* B = sm + 16384;
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) ((unsigned*)sm)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
  const unsigned idesc8 = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(8 >> 3) << 17) | ((unsigned)(M >> 4) << 24);
  if (threadIdx.x < 32) {
    const unsigned long long ad = umma_desc(s32(A), M * 16, 128), bd = umma_desc(s32(B), N * 16, 128);
    const unsigned long long as = (unsigned long long)((2 * M * 16) >> 4), bs = (unsigned long long)((2 * N * 16) >> 4);
    unsigned phase = 0;
    long long st[10];
    for (int rep = 0; rep < 6; ++rep) {
      // idle gap
      const long long g0 = clock64();
      bool warmed = false;
      while (clock64() - g0 < gap) {
        if (warm_before > 0 && !warmed && clock64() - g0 >= gap - warm_before) {
          if (elect_one()) mma(tmem_base + 384, ad, bd, idesc8, 0u);
          __syncwarp();
          warmed = true;
        }
      }
      const long long t0 = clock64();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) { mma(tmem_base, ad + kk * as, bd + kk * bs, idesc, kk > 0); st[kk] = clock64(); }
      }
      __syncwarp();
      if (threadIdx.x == 0) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, phase); phase ^= 1;
      const long long t2 = clock64();
      if (elect_one()) { for (int kk = 0; kk < 8; ++kk) out[kk] = st[kk] - t0; out[8] = t2 - t0; }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 128);
  const int smem = 49152 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int warm : {0, 100, 300})
    for (int gap : {0, 100, 300, 600, 1000, 2000, 5000}) {
      if (warm && gap < warm + 100) continue;
      k<<<1, 128, smem>>>(gap, warm, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[9] = {0};
      cudaMemcpy(h, d, 72, cudaMemcpyDeviceToHost);
      printf("idle gap %5d cycles, keep-warm MMA %3d cycles before: issue stamps %4lld %4lld %4lld %4lld %4lld %4lld %4lld %4lld | complete %5lld %s\n",
             gap, warm, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], e == cudaSuccess ? "" : cudaGetErrorString(e));
      fflush(stdout);
    }
  return 0;
}
