// GPU micro-benchmark (debugging aid): the "row-stacked" tcgen05.mma issue pattern in isolation -- one CTA, operands resident
// in shared memory, no TMA, no epilogue.  An M tile is one input row (128 pixel slots); one MMA multiplies it with the filter
// taps of ALL kh filter rows at once (N = kh * NT) and accumulates into the TMEM column window of the kh output rows it
// feeds, so consecutive input rows write OVERLAPPING windows shifted by NT columns.
//   order 0: rows 0,1,2,...           order 1: rows 0,kh,2kh,..,1,kh+1,.. (consecutive MMAs touch disjoint windows)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o bench_umma_rows bench_umma_rows.cu
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
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }" ::"r"(d), "l"(a),
               "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int BHC>
__global__ void __launch_bounds__(128, 1) k(int NT, int kh, int kw, int chunks, int TH, int bands, int order, int rowslots, long long *out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t *bar = (uint64_t *)(smem + 200 * 1024);
  uint32_t *slot = (uint32_t *)(bar + 1);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) ((float *)smem)[i] = 1.0f;
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
    const uint32_t idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 4) << 24);
    const uint64_t d_hi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
    const uint32_t lbo = 1u << 16;
    const uint32_t a_addr = smem_u32(smem), b_addr0 = smem_u32(smem + 150 * 1024);
    const int BH = TH + kh - 1;
    uint32_t elected;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(elected));
    long long t0 = clock64();
    long long nmma = 0, ncols = 0;
    if (elected) {
      // per-row constants (registers): D column, A row offset, B row offset, instruction descriptor
      uint32_t r_tcol[BHC], r_a[BHC], r_b[BHC], r_id[BHC];
      int ord[BHC];
      {
        int n = 0;
        const int nph = order ? kh : 1, step = order ? kh : 1;
        for (int ph = 0; ph < nph; ++ph)
          for (int i = ph; i < BH; i += step) ord[n++] = i;
      }
#pragma unroll
      for (int q = 0; q < BHC; ++q) {
        const int i = ord[q];
        int lo = i - (kh - 1); if (lo < 0) lo = 0;
        int hi = i < TH - 1 ? i : TH - 1;
        const int nb = hi - lo + 1;
        r_id[q] = idesc0 | ((uint32_t)((nb * NT) >> 3) << 17);
        r_b[q] = (uint32_t)((kh - 1) - (i - lo)) * (uint32_t)NT * 8u;
        r_a[q] = (uint32_t)((i & 7) * rowslots * 8);
        r_tcol[q] = (uint32_t)(lo * NT);
        ncols += nb * NT;
      }
      for (int band = 0; band < bands; ++band) {
        const uint32_t tacc = tmem + (uint32_t)((band & 1) * TH * NT);
        for (int c = 0; c < chunks; ++c) {
          const uint32_t a_chunk = (((a_addr + (uint32_t)(c * 8192)) >> 4) & 0x3FFF) | lbo;
          for (int s = 0; s < kw; ++s) {
            const uint32_t b_lo = (((b_addr0 + (uint32_t)(((c * kw + s) & 1) * kh * NT * 128)) >> 4) & 0x3FFF) | lbo;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
#pragma unroll
              for (int q = 0; q < BHC; ++q)
                umma(tacc + r_tcol[q], d_hi | (uint64_t)(a_chunk + r_a[q] + (uint32_t)(s * 8 + 2 * k4)),
                     d_hi | (uint64_t)(b_lo + r_b[q] + 2u * k4), r_id[q], (c | s | k4) ? 1u : 0u);
              nmma += BHC;
            }
          }
        }
      }
      ncols *= (long long)bands * chunks * kw * 4;
    }
    __syncwarp();
    long long t1 = clock64();
    if (threadIdx.x == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    __syncwarp();
    mbar_wait(bar, 0);
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = nmma; out[3] = ncols; }
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
int main() {
  long long *d, h[4];
  cudaMalloc(&d, 32);
  struct { const char *name; int NT, kh, kw, chunks, TH; } cfg[] = {{"espcn L2 fprop 64->32", 32, 3, 3, 2, 8}, {"espcn L3 fprop 32->48", 48, 3, 3, 1, 5},
                                                                     {"64->64 k3", 64, 3, 3, 2, 4}, {"espcn L2 dgrad 32->64", 64, 3, 3, 1, 4},
                                                                     {"k5 NT16", 16, 5, 5, 2, 12}, {"64->32 TH8 again", 32, 3, 3, 2, 8}};
  for (auto &c : cfg)
    for (int order = 0; order < 2; ++order) {
      const int BH = c.TH + c.kh - 1;
#define RUN(B) case B: cudaFuncSetAttribute(k<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024 + 64); k<B><<<1, 128, 201 * 1024 + 64>>>(c.NT, c.kh, c.kw, c.chunks, c.TH, 20, order, 128, d); break;
      switch (BH) { RUN(6) RUN(7) RUN(10) RUN(16) default: printf("no instantiation for BH %d\n", BH); continue; }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
      const double px = 20.0 * c.TH * 128;  // output pixel slots produced
      const double old_cyc = (32.0 + c.NT / 4.0 > c.NT / 2.0 ? 32.0 + c.NT / 4.0 : c.NT / 2.0) * c.kh * c.kw * c.chunks * 4 / 128.0;
      printf("%-24s order %d: %lld MMAs, mean N %.0f, issue %.1f cyc/mma, complete %.1f cyc/mma -> %.2f cyc/pixel-slot (per-tap model %.2f)\n", c.name, order,
             h[2], (double)h[3] / h[2], (double)h[0] / h[2], (double)h[1] / h[2], (double)h[1] / px, old_cyc);
    }
  return 0;
}
