// GPU probe (debugging aid): MEASURE the shared-memory address function of a tcgen05 matrix descriptor instead of
// testing a hypothesis.  The operand under test is read against a one-hot K-major SW128 partner, so that
//   D[m][n] = A(m, k = n)   (A under test, B(n,k) = delta(n,k), N = 8)      or
//   D[m][n] = B(n, k = m)   (B under test, A(m,k) = delta(m,k), rows m < 8)
// Shared memory holds the float INDEX of every word (two passes: index % 1024 and index / 1024, both exact in tf32),
// so D directly lists which word of shared memory each matrix element came from.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -o probe_desc2 probe_desc2.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <math.h>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int SM_FLOATS = 40960;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}

struct Args {
  uint64_t a_rest, b_rest;
  uint32_t a_off, b_off;
  uint32_t idesc;
  int ncols;
};

__global__ void __launch_bounds__(128, 1) probe(const float *tab, float *D, Args a) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t *bar = (uint64_t *)(smem + SM_FLOATS * 4);
  uint32_t *slot = (uint32_t *)(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < SM_FLOATS; i += 128) ((float *)smem)[i] = tab[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t base = smem_u32(smem);
    uint64_t ad = a.a_rest | (uint64_t)(((base + a.a_off) >> 4) & 0x3FFF);
    uint64_t bd = a.b_rest | (uint64_t)(((base + a.b_off) >> 4) & 0x3FFF);
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                 ::"r"(tmem), "l"(ad), "l"(bd), "r"(a.idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  mbar_wait(bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int j0 = 0; j0 < a.ncols; j0 += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + j0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * 256 + j0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

static uint64_t desc_rest(uint32_t lbo, uint32_t sbo, int layout, int base_off) {
  return ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)(base_off & 7) << 49) | ((uint64_t)layout << 61);
}
static uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
static uint32_t swz128(uint32_t a) { return a ^ (((a >> 7) & 7) << 4); }

struct Case {
  std::string name;
  int side;       // 0: operand under test is A (M = 128 rows listed), 1: operand under test is B (N rows listed)
  int rows;       // M (side 0) or N (side 1) of the operand under test
  uint64_t rest;  // its descriptor (no address)
  uint32_t off;   // its start offset (bytes)
};

int main() {
  float *dT, *dD;
  CK(cudaMalloc(&dT, SM_FLOATS * 4));
  CK(cudaMalloc(&dD, 128 * 256 * 4));
  size_t smem = SM_FLOATS * 4 + 1024 + 64;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const uint32_t POFF = 128 * 1024;  // one-hot partner region (K-major SW128, 128 rows x 128 B)
  const int TEST_FLOATS = POFF / 4;

  std::vector<Case> cases;
  // MN-major candidates for the A side (M = 128)
  cases.push_back({"A MN-major NOSWZ   LBO=4096 SBO=128", 0, 128, desc_rest(4096, 128, 0, 0), 0});
  cases.push_back({"A MN-major NOSWZ   LBO=128 SBO=4096", 0, 128, desc_rest(128, 4096, 0, 0), 0});
  cases.push_back({"A MN-major NOSWZ   LBO=256 SBO=8192", 0, 128, desc_rest(256, 8192, 0, 0), 0});
  cases.push_back({"A MN-major NOSWZ   LBO=16 SBO=128 (overlap?)", 0, 128, desc_rest(16, 128, 0, 0), 0});
  cases.push_back({"A MN-major NOSWZ   LBO=128 SBO=16 (overlap?)", 0, 128, desc_rest(128, 16, 0, 0), 0});
  cases.push_back({"A MN-major SW128_BASE32B LBO=16384 SBO=512", 0, 128, desc_rest(16384, 512, 1, 0), 0});
  cases.push_back({"A MN-major SW128_BASE32B LBO=16384 SBO=512 start+128", 0, 128, desc_rest(16384, 512, 1, 0), 128});
  cases.push_back({"A MN-major SW128   LBO=16384 SBO=1024", 0, 128, desc_rest(16384, 1024, 2, 0), 0});
  cases.push_back({"A MN-major SW64    LBO=8192 SBO=512", 0, 128, desc_rest(8192, 512, 4, 0), 0});
  cases.push_back({"A MN-major SW32    LBO=4096 SBO=256", 0, 128, desc_rest(4096, 256, 6, 0), 0});
  // B side, N = 32
  cases.push_back({"B MN-major NOSWZ   LBO=4096 SBO=128", 1, 32, desc_rest(4096, 128, 0, 0), 0});
  cases.push_back({"B MN-major NOSWZ   LBO=128 SBO=4096", 1, 32, desc_rest(128, 4096, 0, 0), 0});
  cases.push_back({"B MN-major NOSWZ   LBO=16 SBO=128 (overlap?)", 1, 32, desc_rest(16, 128, 0, 0), 0});
  cases.push_back({"B MN-major NOSWZ   LBO=128 SBO=16 (overlap?)", 1, 32, desc_rest(128, 16, 0, 0), 0});
  cases.push_back({"B MN-major SW32    LBO=4096 SBO=256", 1, 32, desc_rest(4096, 256, 6, 0), 0});
  cases.push_back({"B MN-major SW128_BASE32B LBO=16384 SBO=512", 1, 32, desc_rest(16384, 512, 1, 0), 0});

  std::vector<float> tab(SM_FLOATS), hD(128 * 256);
  for (auto &c : cases) {
    std::vector<int> idx(128 * 8, 0);
    for (int pass = 0; pass < 2; ++pass) {
      for (int i = 0; i < SM_FLOATS; ++i) tab[i] = 0.f;
      for (int i = 0; i < TEST_FLOATS; ++i) tab[i] = pass ? (float)(i / 1024) : (float)(i % 1024);
      // one-hot partner: rows j < 8 have a 1 at k = j
      for (int j = 0; j < 8; ++j) tab[swz128(POFF + j * 128 + j * 4) / 4] = 1.f;
      CK(cudaMemcpy(dT, tab.data(), SM_FLOATS * 4, cudaMemcpyHostToDevice));
      CK(cudaMemset(dD, 0, 128 * 256 * 4));
      Args a;
      if (c.side == 0) {
        a = Args{c.rest, desc_rest(16, 1024, 2, 0), c.off, POFF, make_idesc(128, 8, 1, 0), 8};
      } else {
        a = Args{desc_rest(16, 1024, 2, 0), c.rest, POFF, c.off, make_idesc(128, c.rows, 0, 1), c.rows};
      }
      probe<<<1, 128, smem>>>(dT, dD, a);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s : LAUNCH FAILED %s\n", c.name.c_str(), cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(hD.data(), dD, 128 * 256 * 4, cudaMemcpyDeviceToHost));
      for (int r = 0; r < c.rows; ++r)
        for (int k = 0; k < 8; ++k) {
          float v = c.side == 0 ? hD[r * 256 + k] : hD[k * 256 + r];
          int iv = (int)lrintf(v);
          idx[r * 8 + k] += pass ? iv * 1024 : iv;
        }
    }
    printf("== %s  (byte offsets of element (row, k=0..7))\n", c.name.c_str());
    int show[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127};
    for (int r : show) {
      if (r >= c.rows) continue;
      printf("   row %3d:", r);
      for (int k = 0; k < 8; ++k) printf(" %6d", idx[r * 8 + k] * 4);
      printf("\n");
    }
  }
  return 0;
}
