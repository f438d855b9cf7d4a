// Microbenchmark: how much does a train of tcgen05.mma (M=64, N=128, K=16, bf16, SS mode) slow down when the other
// warps of the CTA do what the chain kernel's epilogue does at the same time?
//   mode bit 0: 16 warps store to the A-operand region with st.shared (32-bit stores, the epilogue's pattern)
//   mode bit 1: 16 warps read accumulator columns with tcgen05.ld.16x256b
//   mode bit 2: one thread streams 32 KB weight tiles with cp.async.bulk into a second buffer
//   mode bit 3: the storing warps also execute fence.proxy.async after every batch of stores
//   mode bit 4: 16 warps spin on mbarrier.try_wait (what epilogue warps do while they wait for the accumulator)
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

__global__ void __launch_bounds__(544, 1) k(int mode, int n_mma, const unsigned char* gsrc, long long* out) {
  extern __shared__ unsigned char raw[];
  unsigned char* sm = (unsigned char*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  __shared__ unsigned long long bar, wbar, never;
  __shared__ unsigned tmem_base;
  __shared__ volatile int stop;
  constexpr int M = 64, N = 128;
  unsigned char* A = sm;                    // 16 KB
  unsigned char* B = sm + 16384;            // 32 KB
  unsigned char* A2 = sm + 49152;           // 16 KB: store target (a second A operand, not read by the MMAs)
  unsigned char* W2 = sm + 65536;           // 32 KB: bulk copy target
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 98304 / 4; i += blockDim.x) ((unsigned*)sm)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&wbar, 1); mbar_init(&never, 1); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
  if (warp == 16) {                         // issuer warp
    const unsigned long long ad = umma_desc(s32(A), M * 16, 128), bd = umma_desc(s32(B), N * 16, 128);
    const unsigned long long as = (unsigned long long)((2 * M * 16) >> 4), bs = (unsigned long long)((2 * N * 16) >> 4);
    for (volatile int w = 0; w < 2000; ++w) {}   // let the other warps get going
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 8) {
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) mma(tmem_base, ad + kk * as, bd + kk * bs, idesc, kk > 0);
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (lane == 0) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; stop = 1; }
    __syncwarp();
    if (mode & 32) asm volatile("bar.sync 5, 544;" ::: "memory");
  } else {
    const int q = warp & 3, cb = warp >> 2;
    const unsigned lane_t = tmem_base + ((unsigned)(q * 32) << 16) + 256 + cb * 32;
    unsigned acc = 0;
    int it = 0;
    if (mode & 4) {
      if (tid == 0) {
        unsigned ph = 0;
        while (!stop) {
          mbar_expect_tx(&wbar, 32768);
          bulk_g2s(W2, gsrc + (it & 7) * 32768, 32768, &wbar);
          mbar_wait(&wbar, ph); ph ^= 1; ++it;
        }
      }
    }
    if (mode & 32) { asm volatile("bar.sync 5, 544;" ::: "memory"); }      // truly idle: parked until the issuer joins the barrier
    float fa = (float)tid, fb = 1.0001f;
    while (!(mode & 32) && !stop && it < 100000) {
      ++it;
      if (mode & 64) {
#pragma unroll
        for (int j = 0; j < 64; ++j) fa = fmaf(fa, fb, 0.5f);
        if (fa == 123.f) acc++;
      }
      if (mode & 1) {
        const int rA = 16 * q + (lane >> 2);
#pragma unroll
        for (int g = 0; g < 4; ++g)
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(s32(A2 + (cb * 4 + g) * 1024 + rA * 16 + (lane & 3) * 4)), "r"(acc + g) : "memory");
        if (mode & 8) fence_async_smem();
      }
      if (mode & 2) {
        unsigned r[16];
        tmem_ld16x256_x4(lane_t, r);
        tmem_wait_ld();
        acc += r[0] + r[5];
      }
      if (mode & 16) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(s32(&never)), "r"(0u) : "memory");
        acc += ok;
      }
      if (!(mode & (19 | 64))) __nanosleep(200);
    }
    if (acc == 0x12345678u) out[3] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  unsigned char* g; cudaMalloc(&g, 8 * 32768); cudaMemset(g, 0, 8 * 32768);
  const int smem = 98304 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"napping (nanosleep 200)", "st.shared", "tcgen05.ld", "st.shared + tcgen05.ld", "bulk copies", "st.shared + fence.proxy.async",
                         "parked at a barrier", "busy FMA loops"};
  const int modes[] = {0, 1, 2, 3, 4, 9, 32, 64};
  for (int m = 0; m < 8; ++m)
    for (int n : {8, 96}) {
      k<<<1, 544, smem>>>(modes[m], n, g, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2] = {0, 0};
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("%-32s %2d MMAs: issue %5lld cycles, issue->complete %5lld cycles (%.1f / MMA) %s\n", names[m], n, h[0], h[1], (double)h[1] / n,
             e == cudaSuccess ? "" : cudaGetErrorString(e));
      fflush(stdout);
    }
  return 0;
}
