// GPU probe (debugging aid, not part of the library): does a tcgen05 shared-memory descriptor whose start
// address is offset by a multiple of 128 B (not 1024 B aligned) inside a 128B-swizzled TMA tile read the
// rows it should?  Answers the "shifted view" question for the halo-tile conv/wgrad kernels:
//   mode 0: K-major A, rows m -> smem row (m + shift)            D[m][n] = sum_k X[m+shift][k] * Y[n][k]
//   mode 1: MN-major A and B, K = smem rows (pixels)             D[(b,c)][n] = sum_p X[p+shift+b][c] * Z[p][n]
// for base_offset in {0, (start>>7)&7}.   Build: nvcc -gencode arch=compute_100a,code=sm_100a -o probe_umma probe_umma.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <math.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}

constexpr int XR = 320;  // rows of X
constexpr int ZR = 128;  // rows of Y / Z

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapZ, float *D, int mode, int shift,
      int bo, int ksteps, int lbo_bytes) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t *sX = smem;                 // XR * 128 B
  uint8_t *sZ = smem + XR * 128;      // ZR * 128 B  (XR*128 is a multiple of 1024)
  uint64_t *bar = (uint64_t *)(sZ + ZR * 128);
  uint64_t *bar2 = bar + 1;
  uint32_t *slot = (uint32_t *)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"((uint32_t)((XR + ZR) * 128)) : "memory");
    // X in two boxes (<= 256 rows each)
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sX)), "l"(&mapX), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sX + 160 * 128)), "l"(&mapX), "r"(smem_u32(bar)), "r"(0), "r"(160) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sZ)), "l"(&mapZ), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t aaddr = smem_u32(sX) + shift * 128;
    const uint32_t baddr = smem_u32(sZ);
    uint64_t hi_common = ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    if (mode == 0) {
      // K-major A (M=128 rows from row `shift`), K-major B (N=64 rows of Z), K = 32 (4 steps of 8)
      uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
      for (int k = 0; k < 4; ++k) {
        uint64_t ad = (uint64_t)(((aaddr + k * 32) >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | hi_common | ((uint64_t)(bo & 7) << 49);
        uint64_t bd = (uint64_t)(((baddr + k * 32) >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | hi_common;
        uint32_t acc = k ? 1u : 0u;
        asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                     ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
    } else {
      // MN-major A: M = 4 blocks of 32 channels, block b starts lbo_bytes after block b-1; K = rows.
      // MN-major B: N = 32 channels of Z, K = rows.
      uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
      for (int k = 0; k < ksteps; ++k) {
        const uint64_t hi1 = ((uint64_t)1 << 46) | ((uint64_t)1 << 61);  // SWIZZLE_128B_BASE32B
        uint64_t ad = (uint64_t)(((aaddr + k * 1024) >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)32 << 32) | hi1 | ((uint64_t)(bo & 7) << 49);
        uint64_t bd = (uint64_t)(((baddr + k * 1024) >> 4) & 0x3FFF) | ((uint64_t)64 << 16) | ((uint64_t)32 << 32) | hi1;
        uint32_t acc = k ? 1u : 0u;
        asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                     ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar2)) : "memory");
  }
  mbar_wait(bar2, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int ncols = mode == 0 ? 64 : 32;
  for (int j0 = 0; j0 < ncols; j0 += 16) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + j0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * 64 + j0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main() {
  float *hX = (float *)malloc(XR * 32 * 4), *hZ = (float *)malloc(ZR * 32 * 4);
  srand(1);
  for (int i = 0; i < XR * 32; ++i) hX[i] = (float)((rand() % 17) - 8);
  for (int i = 0; i < ZR * 32; ++i) hZ[i] = (float)((rand() % 13) - 6);
  float *dX, *dZ, *dD;
  CK(cudaMalloc(&dX, XR * 32 * 4)); CK(cudaMalloc(&dZ, ZR * 32 * 4)); CK(cudaMalloc(&dD, 128 * 64 * 4));
  CK(cudaMemcpy(dX, hX, XR * 32 * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dZ, hZ, ZR * 32 * 4, cudaMemcpyHostToDevice));
  CUtensorMap mX, mZ;
  cuuint32_t es[2] = {1, 1};
  for (int mode = 0; mode < 2; ++mode) {
  CUtensorMapSwizzle SWZ = mode ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  {
    cuuint64_t dims[2] = {32, XR}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {32, 160};
    CUresult r = cuTensorMapEncodeTiled(&mX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dX, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        SWZ, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode X failed %d\n", (int)r); return 1; }
  }
  {
    cuuint64_t dims[2] = {32, ZR}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {32, ZR};
    CUresult r = cuTensorMapEncodeTiled(&mZ, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dZ, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        SWZ, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode Z failed %d\n", (int)r); return 1; }
  }
  size_t smem = (XR + ZR) * 128 + 1024 + 64;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  float *hD = (float *)malloc(128 * 64 * 4);
  {
    int lbos[3] = {128, 256, 4096};
    for (int li = 0; li < (mode == 0 ? 1 : 3); ++li) {
      int lbo = lbos[li];
      for (int shift = 0; shift <= 11; shift += (mode == 0 ? 5 : 1)) {
        for (int bmode = 0; bmode < 2; ++bmode) {
          int bo = bmode ? (shift & 7) : 0;
          if (bmode && bo == 0) continue;
          int ksteps = 8;
          CK(cudaMemset(dD, 0, 128 * 64 * 4));
          probe<<<1, 128, smem>>>(mX, mZ, dD, mode, shift, bo, ksteps, lbo);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("mode %d shift %d bo %d: launch failed: %s\n", mode, shift, bo, cudaGetErrorString(e)); return 1; }
          CK(cudaMemcpy(hD, dD, 128 * 64 * 4, cudaMemcpyDeviceToHost));
          double maxerr = 0; int bad = 0;
          if (mode == 0) {
            for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
              double ref = 0; for (int k = 0; k < 32; ++k) ref += (double)hX[(m + shift) * 32 + k] * hZ[n * 32 + k];
              double d = fabs(ref - hD[m * 64 + n]); if (d > maxerr) maxerr = d; if (d > 1e-3) bad++;
            }
          } else {
            for (int b = 0; b < 4; ++b) for (int c = 0; c < 32; ++c) for (int n = 0; n < 32; ++n) {
              double ref = 0;
              for (int p = 0; p < ksteps * 8; ++p) ref += (double)hX[(p + shift + b * (lbo / 128)) * 32 + c] * hZ[p * 32 + n];
              double d = fabs(ref - hD[(b * 32 + c) * 64 + n]); if (d > maxerr) maxerr = d; if (d > 1e-3) bad++;
            }
          }
          printf("mode %d (%s) lbo %4d shift %2d base_offset %d : max err %.3g, bad %d %s\n", mode, mode ? "MN-major" : "K-major", lbo, shift, bo,
                 maxerr, bad, bad ? "WRONG" : "ok");
        }
      }
    }
  }
  }
  return 0;
}
