// GPU probe (debugging aid): does cuTensorMapEncodeTiled accept OVERLAPPING strides (dim-1 stride 16 B while
// dim 0 spans 128 B) and does the TMA then deliver, for box row m, the 32 floats starting at pixel m?
// This is what lets a Cin=3 (padded to 4) conv feed tcgen05 straight from an NHWC4 image: row m of the A tile
// = 8 neighbouring pixels x 4 channels, with no im2col buffer in global memory.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap map, float *out, int x0, int y0) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t *bar = (uint64_t *)(smem + 128 * 128);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(128u * 128u) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(smem)), "l"(&map), "r"(smem_u32(bar)), "r"(0), "r"(x0), "r"(y0), "r"(0) : "memory");
    uint32_t done;
    do {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
    } while (!done);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) out[i] = ((float *)smem)[i];
}
int main() {
  const int W = 40, H = 12, C = 4;  // NHWC4 image, one batch entry
  float *h = (float *)malloc(W * H * C * 4 + 1024);
  for (int i = 0; i < W * H * C + 256; ++i) h[i] = (float)i;
  float *d, *dout;
  cudaMalloc(&d, W * H * C * 4 + 1024);
  cudaMalloc(&dout, 128 * 32 * 4);
  cudaMemcpy(d, h, W * H * C * 4 + 1024, cudaMemcpyHostToDevice);
  CUtensorMap map;
  cuuint64_t dims[4] = {32, (cuuint64_t)(W - 7), (cuuint64_t)H, 1};
  cuuint64_t strides[3] = {16, (cuuint64_t)W * 16, (cuuint64_t)W * H * 16};
  cuuint32_t box[4] = {32, 32, 4, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  for (int swz = 0; swz < 2; ++swz) {
    CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode (swizzle %d) -> CUresult %d\n", swz, (int)r);
    if (r) continue;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    k<<<1, 128, 64 * 1024>>>(map, dout, 1, 2);
    cudaError_t e = cudaDeviceSynchronize();
    if (e) { printf("kernel error %s\n", cudaGetErrorString(e)); return 1; }
    float *o = (float *)malloc(128 * 32 * 4);
    cudaMemcpy(o, dout, 128 * 32 * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m) {
      int ty = m / 32, tx = m % 32;               // box row m -> pixel (y0+ty, x0+tx)
      for (int j = 0; j < 32; ++j) {
        int chunk = j / 4, jj = j % 4;
        int pj = swz ? ((chunk ^ (m & 7)) * 4 + jj) : j;   // 128B swizzle: 16-byte chunk index XOR (row & 7)
        float expect = (float)(((2 + ty) * W + (1 + tx)) * 4 + j);
        if (o[m * 32 + pj] != expect) { if (bad < 5) printf("  m=%d j=%d got %.0f expect %.0f\n", m, j, o[m * 32 + pj], expect); ++bad; }
      }
    }
    printf("swizzle %d: %d mismatches of 4096 %s\n", swz, bad, bad ? "WRONG" : "ok");
  }
  return 0;
}
