// GPU micro-benchmark (debugging aid): cycles per tcgen05.mma kind::tf32 (M=128, K=8) as a function of N and of
// operand major-ness, operands resident in shared memory (contents irrelevant), one CTA, one issuing thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o bench_umma bench_umma.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__global__ void __launch_bounds__(128, 1) k(int N, int mn_major, int iters, int ctas_share, long long *out, int a_shift) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t *bar = (uint64_t *)(smem + 160 * 1024);
  uint32_t *slot = (uint32_t *)(bar + 1);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) ((float *)smem)[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem = *slot;
  tmem = __shfl_sync(0xffffffffu, tmem, 0);
  if (threadIdx.x < 32) {
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    uint64_t hi;
    uint32_t a_lo, b_lo, step;
    if (mn_major) {
      idesc |= (1u << 15) | (1u << 16);
      hi = ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
      a_lo = (((smem_u32(smem) + a_shift) >> 4) & 0x3FFF) | ((128u >> 4) << 16);
      b_lo = ((smem_u32(smem + 64 * 1024) >> 4) & 0x3FFF) | ((16384u >> 4) << 16);
      step = 64;  // 1 KB per K-step
    } else {
      hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      a_lo = (((smem_u32(smem) + a_shift) >> 4) & 0x3FFF) | (1u << 16);
      b_lo = ((smem_u32(smem + 64 * 1024) >> 4) & 0x3FFF) | (1u << 16);
      step = 2;  // 32 B per K-step inside the swizzle row
    }
    uint32_t elected;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(elected));
    long long t0 = clock64();
    if (elected) {
      for (int i = 0; i < iters; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint32_t o = (uint32_t)u * step;
          uint32_t col = tmem + (uint32_t)((u % ctas_share) * N);
          asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                       ::"r"(col), "l"(hi | (uint64_t)(a_lo + o)), "l"(hi | (uint64_t)(b_lo + o)), "r"(idesc), "r"(1u) : "memory");
        }
      }
    }
    __syncwarp();
    long long t1 = clock64();
    if (elected) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    __syncwarp();
    mbar_wait(bar, 0);
    long long t2 = clock64();
    if (elected) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
int main() {
  long long *d, h[2];
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int Ns[] = {16, 32, 48, 64, 128, 256};
  int shifts[] = {0, 128, 640, 1024};
  for (int si = 0; si < 4; ++si)
  for (int mn = 0; mn < 2; ++mn)
    for (int share = 1; share <= 2; ++share)
      for (int ni = 0; ni < 6; ++ni) {
        int N = Ns[ni], iters = 2000;
        k<<<1, 128, 200 * 1024>>>(N, mn, iters, share, d, shifts[si]);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("A shift %4d B  %s N=%3d accumulators=%d : issue %.1f cyc/mma, complete %.1f cyc/mma (math floor %.0f)\n", shifts[si], mn ? "MN-major" : "K-major ", N,
               share, (double)h[0] / iters, (double)h[1] / iters, 128.0 * N / 256.0);
      }
  return 0;
}
