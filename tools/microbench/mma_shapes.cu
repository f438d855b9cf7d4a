// Microbenchmark: issue-to-completion time of a train of tcgen05.mma (kind::f16, bf16 operands from shared memory,
// K-major no-swizzle canonical layout) for M in {64, 128} and N in {32, 64, 128, 256}.  One CTA per SM, one thread issues.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../stochastic-muzero_b200/csrc -o mma_shapes mma_shapes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "smz_tc_ptx.cuh"
using namespace smz_tc;

__device__ __forceinline__ void mma(unsigned d, unsigned long long ad, unsigned long long bd, unsigned idesc, unsigned acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) k(int M, int N, int n_mma, int ksteps, long long* out) {
  extern __shared__ unsigned char raw[];
  unsigned char* sm = (unsigned char*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  __shared__ unsigned long long bar;
  __shared__ unsigned tmem_base;
  unsigned char* A = sm;                 // [K/8][M][8] bf16, K = 128
  unsigned char* B = sm + 128 * 128 * 2; // [K/8][N][8] bf16
  for (int i = threadIdx.x; i < (128 * 128 * 2 + 256 * 128 * 2) / 4; i += blockDim.x) ((unsigned*)sm)[i] = 0x3c003c00u;
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
  if (threadIdx.x < 32) {                 // converged issuer warp, elect.sync keeps the descriptors in uniform registers
    const unsigned long long ad = umma_desc(s32(A), M * 16, 128), bd = umma_desc(s32(B), N * 16, 128);
    const unsigned long long as = (unsigned long long)((2 * M * 16) >> 4), bs = (unsigned long long)((2 * N * 16) >> 4);
    unsigned phase = 0;
    long long best = 1ll << 60;
    for (int rep = 0; rep < 5; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; i += 8) {
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) mma(tmem_base, ad + kk * as, bd + kk * bs, idesc, kk > 0);
        }
        __syncwarp();
      }
      const long long t1 = clock64();
      if (threadIdx.x == 0) umma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, phase); phase ^= 1;
      const long long t2 = clock64();
      if (t2 - t0 < best) { best = t2 - t0; if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; } }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  const int smem = 128 * 128 * 2 + 256 * 128 * 2 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int M : {64, 128})
    for (int N : {32, 64, 128, 256})
      for (int n : {8, 64}) {
        k<<<1, 128, smem>>>(M, N, n, 8, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2] = {0, 0};
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("M=%3d N=%3d  %2d MMAs (K=16 each): issue %5lld cycles, issue->complete %5lld cycles  (%.1f / MMA)  %s\n", M, N, n, h[0], h[1],
               (double)h[1] / n, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
