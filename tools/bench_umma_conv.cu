// GPU micro-benchmark (debugging aid): tcgen05.mma issue pattern of k_conv_sl in isolation -- one CTA, operands resident in
// shared memory, no TMA, no epilogue.  Variants isolate what makes the in-kernel MMAs slower than the 32+N/4 cycle model:
//   v0: constant descriptors            v1: conv address pattern (tap shift, M-tile, K-step)     v2: v1 + per-K-block fence/elect
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o bench_umma_conv bench_umma_conv.cu
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }" ::"r"(d), "l"(a),
               "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128, 1) k(int N, int MTB, int BW, int kh, int kw, int chunks, int bands, int variant, long long *out, int fill) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t *bar = (uint64_t *)(smem + 200 * 1024);
  uint32_t *slot = (uint32_t *)(bar + 1);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) ((float *)smem)[i] = fill == 0 ? 1.0f : (fill == 1 ? (float)(i % 7) - 3.f : (fill == 2 ? 0.f : 1.0f + (float)(i % 5) * 0.25f));
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
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t d_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t lbo = 1u << 16;
    const uint32_t a_addr = smem_u32(smem), b_addr0 = smem_u32(smem + 120 * 1024);
    const uint32_t b_stage = (uint32_t)N * 128u;
    uint32_t elected;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(elected));
    long long t0 = clock64();
    long long nmma = 0;
    if (elected) {
      for (int band = 0; band < bands; ++band) {
        const uint32_t tacc = tmem + (uint32_t)((band & 1) * MTB * N);
        for (int c = 0; c < chunks; ++c) {
          uint32_t a_tap = (((a_addr + (uint32_t)(c * 40960)) >> 4) & 0x3FFF) | lbo;
          uint32_t b_lo = ((b_addr0 >> 4) & 0x3FFF) | lbo;
          for (int r = 0; r < kh; ++r) {
            for (int s = 0; s < kw; ++s) {
              const uint32_t acc0 = (c | r | s) ? 1u : 0u;
              if (variant == 0) {         // constant operands, one accumulator
#pragma unroll
                for (int u = 0; u < 16; ++u)
                  if (u < 4 * MTB) umma(tacc, ((uint64_t)d_hi << 32) | a_tap, ((uint64_t)d_hi << 32) | b_lo, idesc, 1u);
              } else if (variant == 1) {  // K-step outer, M-tile inner, A tiles 16 KB apart, constant tap
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    if (j < MTB) umma(tacc + (uint32_t)(j * N), ((uint64_t)d_hi << 32) | (uint64_t)(a_tap + j * 1024 + 2 * k4), ((uint64_t)d_hi << 32) | (uint64_t)(b_lo + 2 * k4), idesc, k4 ? 1u : acc0);
              } else if (variant == 2) {  // M-tile outer, K-step inner
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                  for (int k4 = 0; k4 < 4; ++k4)
                    if (j < MTB) umma(tacc + (uint32_t)(j * N), ((uint64_t)d_hi << 32) | (uint64_t)(a_tap + j * 1024 + 2 * k4), ((uint64_t)d_hi << 32) | (uint64_t)(b_lo + 2 * k4), idesc, k4 ? 1u : acc0);
              } else {                    // variant 1 + moving taps and weight stages (the real pattern)
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    if (j < MTB) umma(tacc + (uint32_t)(j * N), ((uint64_t)d_hi << 32) | (uint64_t)(a_tap + j * 1024 + 2 * k4), ((uint64_t)d_hi << 32) | (uint64_t)(b_lo + 2 * k4), idesc, k4 ? 1u : acc0);
                a_tap += 8u;
                b_lo += b_stage >> 4;
              }
              nmma += 4 * MTB;
            }
            if (variant == 3) a_tap += (uint32_t)(BW - kw) * 8u;
          }
        }
      }
    }
    __syncwarp();
    long long t1 = clock64();
    if (threadIdx.x == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    __syncwarp();
    mbar_wait(bar, 0);
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = nmma; }
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
int main() {
  long long *d, h[3];
  cudaMalloc(&d, 24);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024 + 64);
  struct { const char *name; int N, MTB, BW, kh, kw, chunks; } cfg[] = {{"espcn L2 fprop", 32, 3, 60, 3, 3, 2}, {"vdsr body", 64, 4, 45, 3, 3, 2},
                                                                         {"N=64 MTB=1", 64, 1, 45, 3, 3, 2}, {"N=32 MTB=1", 32, 1, 60, 3, 3, 2}};
  for (int fill = 0; fill < 1; ++fill)
  for (auto &c : cfg)
    for (int v = 0; v < 4; ++v) {
      k<<<1, 128, 201 * 1024 + 64>>>(c.N, c.MTB, c.BW, c.kh, c.kw, c.chunks, 20, v, d, fill);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
      printf("fill %d %-16s variant %d: %lld MMAs, issue %.1f cyc/mma, complete %.1f cyc/mma (model %.0f)\n", fill, c.name, v, h[2], (double)h[0] / h[2],
             (double)h[1] / h[2], 32.0 + c.N / 4.0);
    }
  return 0;
}
