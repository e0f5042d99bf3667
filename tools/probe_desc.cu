// GPU probe (debugging aid, not part of the library): semantics of tcgen05 shared-memory matrix descriptors for the
// "shifted view" / "overlapping view" operand layouts the halo-tile conv kernels rely on.
//
// Shared memory is filled LINEARLY from a table of small integers (no TMA, no swizzled writes), the kernel issues
// tf32 MMAs with the given descriptors, and the host recomputes D under a hypothesised address function
//     phys_byte(row, k)   (including the XOR swizzle on the absolute shared-memory address)
// A case passes iff every D element matches.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -o probe_desc probe_desc.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <math.h>
#include <functional>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int SM_FLOATS = 40960;  // 160 KB image of shared memory

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}

struct Args {
  uint64_t a_rest, b_rest;   // descriptors without the start-address field
  uint32_t a_off, b_off;     // byte offsets of the operand starts from the 1024-aligned smem base
  uint32_t a_step, b_step;   // start-address advance per K step (bytes)
  uint32_t idesc;
  int ksteps, ncols;
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
  // zero the accumulator lanes first so that unwritten lanes (M = 64) read back as a marker-free 0
  if (threadIdx.x == 0) {
    const uint32_t base = smem_u32(smem);
    for (int k = 0; k < a.ksteps; ++k) {
      uint64_t ad = a.a_rest | (uint64_t)(((base + a.a_off + k * a.a_step) >> 4) & 0x3FFF);
      uint64_t bd = a.b_rest | (uint64_t)(((base + a.b_off + k * a.b_step) >> 4) & 0x3FFF);
      uint32_t acc = k ? 1u : 0u;
      asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                   ::"r"(tmem), "l"(ad), "l"(bd), "r"(a.idesc), "r"(acc) : "memory");
    }
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

// ---------------------------------------------------------------- host side
static uint64_t desc_rest(uint32_t lbo, uint32_t sbo, int layout, int base_off) {
  return ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)(base_off & 7) << 49) | ((uint64_t)layout << 61);
}
static uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
static uint32_t swz128(uint32_t a) { return a ^ (((a >> 7) & 7) << 4); }      // 16B atoms, 8 rows
static uint32_t swz128_32(uint32_t a) { return a ^ (((a >> 7) & 3) << 5); }   // 32B atoms, 4 rows

struct Case {
  std::string name;
  int M, N, ksteps;
  Args args;
  // physical byte address (relative to the aligned smem base) of A(m, k) / B(n, k), k in [0, 8*ksteps)
  std::function<uint32_t(int, int)> pa, pb;
};

int main() {
  std::vector<float> tab(SM_FLOATS);
  srand(7);
  for (auto &v : tab) v = (float)((rand() % 15) - 7);
  float *dT, *dD;
  CK(cudaMalloc(&dT, SM_FLOATS * 4));
  CK(cudaMalloc(&dD, 128 * 256 * 4));
  CK(cudaMemcpy(dT, tab.data(), SM_FLOATS * 4, cudaMemcpyHostToDevice));
  size_t smem = SM_FLOATS * 4 + 1024 + 64;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

  std::vector<Case> cases;
  const uint32_t BOFF = 96 * 1024;  // B region (1024-aligned)

  // ---- standard B operands
  auto b_sw128 = [&](int) { return desc_rest(16, 1024, 2, 0); };
  auto pb_sw128 = [=](int n, int k) { return swz128(BOFF + n * 128 + (k % 32) * 4 + (k / 32) * 0) ; };
  (void)b_sw128;

  // C1/C2: K-major SW128 A, start shifted by multiples of 128 B; K = 32 per tile (4 steps of +32 B)
  for (int shift : {0, 1, 3, 5, 8, 13, 66}) {
    for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
      int bo = bo_mode ? (shift & 7) : 0;
      if (bo_mode && bo == 0) continue;
      Case c;
      c.name = "Kmajor-SW128 shift=" + std::to_string(shift) + " base_offset=" + std::to_string(bo) + " [abs-address swizzle]";
      c.M = 128; c.N = 64; c.ksteps = 4;
      c.args = Args{desc_rest(16, 1024, 2, bo), desc_rest(16, 1024, 2, 0), (uint32_t)shift * 128, BOFF, 32, 32,
                    make_idesc(128, 64, 0, 0), 4, 64};
      c.pa = [=](int m, int k) { return swz128((uint32_t)shift * 128 + m * 128 + k * 4); };
      c.pb = [=](int n, int k) { return swz128(BOFF + n * 128 + k * 4); };
      cases.push_back(c);
    }
  }
  // C3a: K-major no-swizzle, canonical, asymmetric LBO/SBO to learn the roles.  K = 8 per MMA = 2 core matrices.
  //   hypothesis H1: addr = (r%8)*16 + (r/8)*SBO + (kb/16)*LBO + kb%16
  for (int swap = 0; swap < 2; ++swap) {
    Case c;
    uint32_t lbo = 128, sbo = 256;  // H1 layout: [r/8][k/4][r%8][4 floats]
    c.name = std::string("Kmajor-NOSWZ canonical, roles ") + (swap ? "SWAPPED (LBO=row-group, SBO=K)" : "H1 (LBO=K, SBO=row-group)");
    c.M = 128; c.N = 64; c.ksteps = 2;
    uint32_t dl = swap ? sbo : lbo, ds = swap ? lbo : sbo;
    // per K step the operand advances by one full [M/8 or N/8][2][8][4] tile
    c.args = Args{desc_rest(dl, ds, 0, 0), desc_rest(dl, ds, 0, 0), 0, BOFF, 16 * 256, 8 * 256, make_idesc(128, 64, 0, 0), 2, 64};
    c.pa = [=](int m, int k) { int ks = k / 8, kk = k % 8; return (uint32_t)(ks * 16 * 256 + (m % 8) * 16 + (m / 8) * sbo + (kk / 4) * lbo + (kk % 4) * 4); };
    c.pb = [=](int n, int k) { int ks = k / 8, kk = k % 8; return (uint32_t)(BOFF + ks * 8 * 256 + (n % 8) * 16 + (n / 8) * sbo + (kk / 4) * lbo + (kk % 4) * 4); };
    cases.push_back(c);
  }
  // C3b: K-major no-swizzle OVERLAPPING view: A(m, k) = F[start + 16 m + 4 k]  (LBO = 16, SBO = 128), +32 B per K step
  for (int q : {0, 1, 5, 67}) {
    for (int swap = 0; swap < 2; ++swap) {
      Case c;
      c.name = "Kmajor-NOSWZ OVERLAP q=" + std::to_string(q) + (swap ? " roles SWAPPED" : " roles H1");
      c.M = 128; c.N = 64; c.ksteps = 3;
      uint32_t al = swap ? 128 : 16, as = swap ? 16 : 128;
      uint32_t bl = swap ? 256 : 128, bs = swap ? 128 : 256;
      c.args = Args{desc_rest(al, as, 0, 0), desc_rest(bl, bs, 0, 0), (uint32_t)q * 16, BOFF, 32, 8 * 256, make_idesc(128, 64, 0, 0), 3, 64};
      c.pa = [=](int m, int k) { return (uint32_t)(q * 16 + 16 * m + 4 * k); };
      c.pb = [=](int n, int k) { int ks = k / 8, kk = k % 8; return (uint32_t)(BOFF + ks * 8 * 256 + (n % 8) * 16 + (n / 8) * 256 + (kk / 4) * 128 + (kk % 4) * 4); };
      cases.push_back(c);
    }
  }
  // C4: MN-major no-swizzle.  hypothesis H2: addr(mn, k) = (mn%4)*4 + (mn/4)*SBO + (k%8)*16 + (k/8)*LBO
  //   A canonical [k/8][m/4][k%8][m%4]: SBO = 128, LBO = (M/4)*128;   B OVERLAPPING: B(n, k) = F[b0 + 16 k + 4 n] (SBO = 16, LBO = 128)
  for (int M : {128, 64}) {
    for (int swap = 0; swap < 2; ++swap) {
      for (int q : {0, 3}) {
        Case c;
        c.name = "MNmajor-NOSWZ A canonical M=" + std::to_string(M) + ", B OVERLAP N=32 q=" + std::to_string(q) + (swap ? " roles SWAPPED" : " roles H2");
        c.M = M; c.N = 32; c.ksteps = 3;
        uint32_t a_sbo = 128, a_lbo = (uint32_t)(M / 4) * 128, b_sbo = 16, b_lbo = 128;
        c.args = Args{swap ? desc_rest(a_sbo, a_lbo, 0, 0) : desc_rest(a_lbo, a_sbo, 0, 0),
                      swap ? desc_rest(b_sbo, b_lbo, 0, 0) : desc_rest(b_lbo, b_sbo, 0, 0),
                      0, BOFF + (uint32_t)q * 16, a_lbo, 128, make_idesc(M, 32, 1, 1), 3, 32};
        c.pa = [=](int m, int k) { return (uint32_t)((m % 4) * 4 + (m / 4) * a_sbo + (k % 8) * 16 + (k / 8) * a_lbo); };
        c.pb = [=](int n, int k) { return (uint32_t)(BOFF + q * 16 + 16 * k + 4 * n); };
        cases.push_back(c);
      }
    }
  }
  // C5: A = MN-major SW128_BASE32B (TMA 128B_ATOM_32B image: [pixel][32 ch] rows of 128 B, 4-row atoms 512 B apart),
  //     M = 64 / 128 (2 / 4 channel blocks LBO apart), B = no-swizzle overlapping view (roles per H2).
  for (int M : {128, 64}) {
    Case c;
    c.name = "A MNmajor-SW128_BASE32B M=" + std::to_string(M) + " x B MNmajor-NOSWZ OVERLAP N=32";
    c.M = M; c.N = 32; c.ksteps = 4;
    const uint32_t blk = 16384;  // bytes between 32-channel blocks
    c.args = Args{desc_rest(blk, 512, 1, 0), desc_rest(128, 16, 0, 0), 0, BOFF, 1024, 128, make_idesc(M, 32, 1, 1), 4, 32};
    c.pa = [=](int m, int k) { return swz128_32((uint32_t)((m / 32) * blk + k * 128 + (m % 32) * 4)); };
    c.pb = [=](int n, int k) { return (uint32_t)(BOFF + 16 * k + 4 * n); };
    cases.push_back(c);
  }
  // C6: K-major SW128 A shifted view with N = 32 / 48 B tiles and an M = 64 A (TMEM row placement of M = 64)
  {
    Case c;
    c.name = "Kmajor-SW128 M=64 N=32 shift=5";
    c.M = 64; c.N = 32; c.ksteps = 4;
    c.args = Args{desc_rest(16, 1024, 2, 0), desc_rest(16, 1024, 2, 0), 5 * 128, BOFF, 32, 32, make_idesc(64, 32, 0, 0), 4, 32};
    c.pa = [=](int m, int k) { return swz128((uint32_t)(5 * 128 + m * 128 + k * 4)); };
    c.pb = [=](int n, int k) { return swz128(BOFF + n * 128 + k * 4); };
    cases.push_back(c);
  }

  std::vector<float> hD(128 * 256);
  for (auto &c : cases) {
    CK(cudaMemset(dD, 0, 128 * 256 * 4));
    probe<<<1, 128, smem>>>(dT, dD, c.args);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-90s : LAUNCH FAILED %s\n", c.name.c_str(), cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(hD.data(), dD, 128 * 256 * 4, cudaMemcpyDeviceToHost));
    const int K = 8 * c.ksteps;
    // expected
    std::vector<double> ref((size_t)c.M * c.N);
    bool oob = false;
    for (int m = 0; m < c.M; ++m)
      for (int n = 0; n < c.N; ++n) {
        double s = 0;
        for (int k = 0; k < K; ++k) {
          uint32_t ia = c.pa(m, k) / 4, ib = c.pb(n, k) / 4;
          if (ia >= (uint32_t)SM_FLOATS || ib >= (uint32_t)SM_FLOATS) { oob = true; continue; }
          s += (double)tab[ia] * tab[ib];
        }
        ref[(size_t)m * c.N + n] = s;
      }
    // lane maps to try: identity; M=64: rows 0..31 -> lanes 0..31, rows 32..63 -> lanes 64..95 etc.
    const char *lm_names[3] = {"lane=m", "lane=m%32+64*(m/32)", "lane=m%16+32*(m/16)"};
    int best_bad = 1 << 30, best_lm = -1;
    for (int lm = 0; lm < 3; ++lm) {
      int bad = 0;
      for (int m = 0; m < c.M; ++m) {
        int lane = lm == 0 ? m : (lm == 1 ? (m % 32) + 64 * (m / 32) : (m % 16) + 32 * (m / 16));
        if (lane >= 128) { bad += c.N; continue; }
        for (int n = 0; n < c.N; ++n)
          if (fabs(ref[(size_t)m * c.N + n] - hD[lane * 256 + n]) > 1e-3) bad++;
      }
      if (bad < best_bad) { best_bad = bad; best_lm = lm; }
    }
    printf("%-100s : %s (bad %d / %d, %s%s)\n", c.name.c_str(), best_bad == 0 ? "MATCH" : "mismatch", best_bad, c.M * c.N,
           lm_names[best_lm], oob ? ", OOB!" : "");
    if (best_bad && best_bad < c.M * c.N) {
      // which rows are wrong?
      printf("    wrong rows:");
      int shown = 0;
      for (int m = 0; m < c.M && shown < 24; ++m) {
        int bad = 0;
        for (int n = 0; n < c.N; ++n) if (fabs(ref[(size_t)m * c.N + n] - hD[m * 256 + n]) > 1e-3) bad++;
        if (bad) { printf(" %d", m); shown++; }
      }
      printf("\n");
    }
  }
  return 0;
}
