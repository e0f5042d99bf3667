// tcgen05 (5th-gen tensor core) implicit-GEMM convolution for sm_100a: fprop and stride-1 dgrad.
//
//   D[m, n] = sum_{tap, c} A_tap[m, c] * B_tap[n, c]       m = output pixel, n = output channel
//
// * A (activations, NHWC fp32 holding tf32-rounded values) is never im2col'ed in global memory: for
//   every filter tap the TMA engine fetches the shifted TH x TW pixel patch x 32 channels straight
//   into 128B-swizzled shared memory (4-D tiled tensor map; out-of-image rows/cols are zero-filled by
//   the TMA unit, which is the conv padding).  The im2col matrix exists only as SMEM stages.
// * B (weights) is repacked per call into [tap][Cout_pad][Cin] (tf32-rounded, K-major) by a tiny
//   kernel and fetched by TMA per (tap, 32-channel chunk); one B stage is shared by up to 4 M-tiles.
// * tcgen05.mma kind::tf32, M=128 x N<=256 x K=8, fp32 accumulators in TMEM (MT x N columns).
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer, warps 2..5 = epilogue
//   (tcgen05.ld -> bias -> ReLU/PReLU/LeakyReLU -> +residual -> store at the pixel-shuffled address).
#include "srb_common.cuh"
#include <cuda.h>

namespace srb {

namespace {

constexpr int kThreads = 192;
constexpr int kChunkC = 32;             // channels per K-block: 32 fp32 = 128 B = one swizzle row
constexpr int kABytes = 128 * 128;      // one A stage tile: 128 pixels x 128 B
constexpr int kMaxSmem = 227 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

__device__ __forceinline__ void tma_load_4d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), version 1 (sm_100).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);  // start address
  d |= (uint64_t)1 << 16;                   // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset
  d |= (uint64_t)1 << 46;                   // descriptor version
  d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

struct TcArgs {
  // geometry of the gather conv this launch computes
  int N, Cin, Ho, Wo, Co;       // Co = real number of output channels (GEMM N before padding)
  int kh, kw, pad;
  int TH, TW;                   // pixel patch of one M-tile, TH*TW == 128
  int tiles_h, tiles_w;         // tiles per image
  int num_tiles;                // N * tiles_h * tiles_w
  int MT;                       // M-tiles per CTA (share each B stage)
  int NT;                       // GEMM N tile (multiple of 16, <= 256)
  int stages;
  int tmem_cols;                // power of two >= MT*NT
  int ps;                       // pixel-shuffle factor for the output addressing
  T4 out;
  Epi epi;
};

__global__ void __launch_bounds__(kThreads, 1)
k_tc_conv(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][MT*16KB A | NT*128 B] then barriers
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = a.MT * kABytes + a.NT * 128;
  uint64_t *full_bar = (uint64_t *)(smem + (size_t)a.stages * stage_bytes);
  uint64_t *empty_bar = full_bar + a.stages;
  uint64_t *accum_bar = empty_bar + a.stages;
  uint32_t *tmem_slot = (uint32_t *)(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile0 = blockIdx.x * a.MT;
  const int n0 = blockIdx.y * a.NT;
  int mt_valid = a.num_tiles - tile0;
  if (mt_valid > a.MT) mt_valid = a.MT;
  const int chunks = a.Cin / kChunkC;
  const int kblocks = a.kh * a.kw * chunks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int tn[4], toh[4], tow[4];
      for (int t = 0; t < mt_valid; ++t) {
        int tile = tile0 + t;
        int tw_i = tile % a.tiles_w;
        int q = tile / a.tiles_w;
        int th_i = q % a.tiles_h;
        tn[t] = q / a.tiles_h;
        toh[t] = th_i * a.TH - a.pad;
        tow[t] = tw_i * a.TW - a.pad;
      }
      const uint32_t tx_bytes = (uint32_t)(mt_valid * kABytes + a.NT * 128);
      int kb = 0;
      for (int tap = 0; tap < a.kh * a.kw; ++tap) {
        const int r = tap / a.kw, s = tap - r * a.kw;
        for (int c = 0; c < chunks; ++c, ++kb) {
          const int st = kb % a.stages;
          const uint32_t ph = (uint32_t)(kb / a.stages) & 1u;
          mbar_wait(&empty_bar[st], ph ^ 1u);
          uint8_t *sa = smem + (size_t)st * stage_bytes;
          mbar_expect_tx(&full_bar[st], tx_bytes);
          for (int t = 0; t < mt_valid; ++t)
            tma_load_4d(&mapA, &full_bar[st], sa + t * kABytes, c * kChunkC, tow[t] + s, toh[t] + r, tn[t]);
          tma_load_3d(&mapB, &full_bar[st], sa + a.MT * kABytes, c * kChunkC, n0, tap);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=tf32, both K-major, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.NT >> 3) << 17) | ((128u >> 4) << 24);
      for (int kb = 0; kb < kblocks; ++kb) {
        const int st = kb % a.stages;
        const uint32_t ph = (uint32_t)(kb / a.stages) & 1u;
        mbar_wait(&full_bar[st], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(smem + (size_t)st * stage_bytes);
        const uint64_t bdesc0 = make_kmajor_sw128_desc(sa + a.MT * kABytes);
        for (int t = 0; t < mt_valid; ++t) {
          const uint64_t adesc0 = make_kmajor_sw128_desc(sa + t * kABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 4 x (K = 8 tf32 = 32 B) inside the 128 B swizzle row
            umma_tf32(tmem_base + (uint32_t)(t * a.NT), adesc0 + (uint64_t)(k * 2), bdesc0 + (uint64_t)(k * 2), idesc,
                      (kb | k) ? 1u : 0u);
        }
        umma_commit(&empty_bar[st]);  // frees this smem stage once the MMAs above have read it
      }
      umma_commit(accum_bar);  // all accumulators complete
    }
  } else {
    // ===================== epilogue: warps 2..5 -> TMEM lanes 32*(warp%4) .. +31 =====================
    const int lane_grp = warp & 3;
    mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const float slope = (a.epi.act == SRB_ACT_PRELU) ? __ldg(a.epi.alpha) : a.epi.slope;
    const int m = lane_grp * 32 + lane;  // row of the M-tile == pixel in the patch
    const int th = m / a.TW, tw = m - th * a.TW;
    for (int t = 0; t < mt_valid; ++t) {
      const int tile = tile0 + t;
      const int tw_i = tile % a.tiles_w;
      const int q = tile / a.tiles_w;
      const int th_i = q % a.tiles_h;
      const int n = q / a.tiles_h;
      const int oy = th_i * a.TH + th, ox = tw_i * a.TW + tw;
      const bool pix_ok = (oy < a.Ho) && (ox < a.Wo);
      const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(t * a.NT);
      for (int j0 = 0; j0 < a.NT; j0 += 16) {
        uint32_t v[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr + (uint32_t)j0)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (!pix_ok) continue;
        const int cbase = n0 + j0;
        if (cbase >= a.Co) continue;
        const bool vec = (a.ps == 1) && (a.out.sc == 1) && (cbase + 16 <= a.Co) && ((a.Co & 3) == 0) &&
                         (!a.epi.residual.p || a.epi.residual.sc == 1) && (!a.epi.preact.p || a.epi.preact.sc == 1);
        if (vec) {
          float *op = a.out.p + n * a.out.sn + (long long)oy * a.out.sh + (long long)ox * a.out.sw + cbase;
          const float *rp = a.epi.residual.p ? a.epi.residual.p + n * a.epi.residual.sn +
                                                   (long long)oy * a.epi.residual.sh + (long long)ox * a.epi.residual.sw + cbase
                                             : nullptr;
          float *pp = a.epi.preact.p ? a.epi.preact.p + n * a.epi.preact.sn + (long long)oy * a.epi.preact.sh +
                                           (long long)ox * a.epi.preact.sw + cbase
                                     : nullptr;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 z;
            z.x = __uint_as_float(v[j + 0]); z.y = __uint_as_float(v[j + 1]);
            z.z = __uint_as_float(v[j + 2]); z.w = __uint_as_float(v[j + 3]);
            if (a.epi.bias) {
              float4 b = __ldg((const float4 *)(a.epi.bias + cbase + j));
              z.x += b.x; z.y += b.y; z.z += b.z; z.w += b.w;
            }
            if (pp) *(float4 *)(pp + j) = z;
            float4 y;
            y.x = apply_act(z.x, a.epi.act, slope); y.y = apply_act(z.y, a.epi.act, slope);
            y.z = apply_act(z.z, a.epi.act, slope); y.w = apply_act(z.w, a.epi.act, slope);
            if (rp) {
              float4 rr = __ldg((const float4 *)(rp + j));
              y.x += rr.x; y.y += rr.y; y.z += rr.z; y.w += rr.w;
            }
            if (a.epi.round_tf32) { y.x = round_tf32(y.x); y.y = round_tf32(y.y); y.z = round_tf32(y.z); y.w = round_tf32(y.w); }
            *(float4 *)(op + j) = y;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int co = cbase + j;
            if (co < a.Co) {
              float z = __uint_as_float(v[j]) + (a.epi.bias ? __ldg(a.epi.bias + co) : 0.f);
              if (a.epi.preact.p) a.epi.preact.p[ps_offset(a.epi.preact, a.ps, n, co, oy, ox)] = z;
              float y = apply_act(z, a.epi.act, slope);
              if (a.epi.residual.p) y += __ldg(a.epi.residual.p + ps_offset(a.epi.residual, a.ps, n, co, oy, ox));
              if (a.epi.round_tf32) y = round_tf32(y);
              a.out.p[ps_offset(a.out, a.ps, n, co, oy, ox)] = y;
            }
          }
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols)
                 : "memory");
  }
}

// Repack weights (Conv2d OIHW, tf32 RN) into the GEMM B operand [tap][Co_pad][Cin]:
//   flip == 0 (fprop):  B[r*kw+s][co][ci] = w[co][ci][r][s]                       (N = Co, K = Ci)
//   flip == 1 (dgrad):  B[(kh-1-r)*kw + (kw-1-s)][ci][co] = w[co][ci][r][s]        (N = Ci, K = Co)
// Rows n >= N_real are zero.
__global__ void k_pack_weights(const float *__restrict__ w, float *__restrict__ out, int Co, int Ci, int kh, int kw,
                               int Npad, int flip) {
  const int N = flip ? Ci : Co, K = flip ? Co : Ci;
  const long long total = (long long)kh * kw * Npad * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int k = (int)(i % K);
    long long q = i / K;
    int n = (int)(q % Npad);
    int tap = (int)(q / Npad);
    float v = 0.f;
    if (n < N) {
      int r = tap / kw, s = tap - r * kw;
      if (flip) { r = kh - 1 - r; s = kw - 1 - s; }
      int co = flip ? k : n, ci = flip ? n : k;
      v = round_tf32(__ldg(w + (((long long)co * Ci + ci) * kh + r) * kw + s));
    }
    out[i] = v;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_tiled(CUtensorMap *map, void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides_bytes,
                 const cuuint32_t *box) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = cuTensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, base, dims, strides_bytes,
                                      box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    const char *msg = nullptr;
    cuGetErrorString(r, &msg);
    set_error("cuTensorMapEncodeTiled failed: %s", msg ? msg : "?");
    return SRB_ECUDA;
  }
  return SRB_OK;
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct Plan {
  int TH, TW, tiles_h, tiles_w, MT, NT, n_tiles_n, Npad, stages, tmem_cols;
  size_t smem;
};

bool make_plan(const Geom &g, Plan *p) {
  // N tile: multiple of 16, <= 256
  int Npad = round_up(g.Co, 16);
  int NT = Npad;
  if (NT > 256) {
    NT = 256;
    while (Npad % NT) NT -= 16;  // largest multiple of 16 dividing Npad (>= 16)
  }
  // pixel patch minimising padded work
  int bestTW = 0;
  long long best = -1;
  for (int TW = 8; TW <= 128; TW *= 2) {
    int TH = 128 / TW;
    long long work = (long long)round_up(g.Wo, TW) * round_up(g.Ho, TH);
    if (best < 0 || work < best || (work == best && TW > bestTW)) { best = work; bestTW = TW; }
  }
  p->TW = bestTW;
  p->TH = 128 / bestTW;
  p->tiles_w = (g.Wo + p->TW - 1) / p->TW;
  p->tiles_h = (g.Ho + p->TH - 1) / p->TH;
  p->NT = NT;
  p->Npad = Npad;
  p->n_tiles_n = Npad / NT;
  long long num_tiles = (long long)g.N * p->tiles_h * p->tiles_w;
  int MT = 512 / NT;
  if (MT > 4) MT = 4;
  if (MT < 1) MT = 1;
  // do not starve the 148 SMs on small problems
  while (MT > 1 && (num_tiles + MT - 1) / MT * p->n_tiles_n < 148 * 2) MT >>= 1;
  p->MT = MT;
  int cols = MT * NT, tc = 32;
  while (tc < cols) tc <<= 1;
  p->tmem_cols = tc;
  size_t stage = (size_t)MT * kABytes + (size_t)NT * 128;
  int stages = (int)((kMaxSmem - 2048) / stage);
  if (stages > 8) stages = 8;
  p->stages = stages;
  p->smem = (size_t)stages * stage + 1024 /*align slack*/ + (2 * stages + 1) * 8 + 16;
  return stages >= 2 && tc <= 512;
}

}  // namespace

bool tc_conv_supported(const Geom &g, const T4 &in, const T4 &out, bool /*dgrad*/) {
  if (g.st != 1 || g.N <= 0) return false;
  if (g.Ci % kChunkC != 0) return false;
  if (in.sc != 1) return false;                                       // channels_last activations
  if ((in.sw % 4) || (in.sh % 4) || (in.sn % 4)) return false;       // TMA strides: multiples of 16 B
  if (((uintptr_t)in.p) & 15) return false;
  if (g.kh > 16 || g.kw > 16) return false;
  if ((long long)g.N * g.Hi * g.Wi * g.Ci >= (1LL << 40)) return false;
  Plan p;
  return make_plan(g, &p);
}

size_t tc_conv_ws_bytes(const Geom &g) {
  int a = round_up(g.Co, 16), b = round_up(g.Ci, 16);
  int npad = a > b ? a : b;
  int k = g.Co > g.Ci ? g.Co : g.Ci;
  return (size_t)g.kh * g.kw * npad * k * sizeof(float) + 256;
}

int tc_conv_gather(const Geom &g, const T4 &in, const float *w, bool flip_transpose, const T4 &out, const Epi &epi,
                   void *ws, size_t ws_bytes, cudaStream_t st) {
  Plan p;
  SRB_REQUIRE(make_plan(g, &p), SRB_EUNSUPPORTED, "tc_conv: no tile plan");
  size_t need = (size_t)g.kh * g.kw * p.Npad * g.Ci * sizeof(float);
  uintptr_t wsp = ((uintptr_t)ws + 255) & ~(uintptr_t)255;
  SRB_REQUIRE(ws && wsp + need <= (uintptr_t)ws + ws_bytes, SRB_EWORKSPACE, "tc_conv workspace: need %zu bytes", need + 256);
  float *wp = (float *)wsp;

  // 1. pack weights.  In dgrad orientation the original filter is (Co_orig = g.Ci ... ) -- see caller:
  //    g here is already the *gather* geometry, so the stored filter is w[co_orig][ci_orig] with
  //    co_orig = g.Ci (K side) and ci_orig = g.Co (N side) when flip_transpose is set.
  {
    long long total = (long long)g.kh * g.kw * p.Npad * g.Ci;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (!flip_transpose)
      k_pack_weights<<<blocks, 256, 0, st>>>(w, wp, g.Co, g.Ci, g.kh, g.kw, p.Npad, 0);
    else
      k_pack_weights<<<blocks, 256, 0, st>>>(w, wp, /*Co_orig=*/g.Ci, /*Ci_orig=*/g.Co, g.kh, g.kw, p.Npad, 1);
    count_launch();
    SRB_CHECK_CUDA(cudaGetLastError());
  }

  // 2. tensor maps
  CUtensorMap mapA, mapB;
  {
    cuuint64_t dims[4] = {(cuuint64_t)g.Ci, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)in.sw * 4, (cuuint64_t)in.sh * 4, (cuuint64_t)in.sn * 4};
    cuuint32_t box[4] = {(cuuint32_t)kChunkC, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1};
    int rc = encode_tiled(&mapA, in.p, 4, dims, strides, box);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)g.Ci, (cuuint64_t)p.Npad, (cuuint64_t)(g.kh * g.kw)};
    cuuint64_t strides[2] = {(cuuint64_t)g.Ci * 4, (cuuint64_t)g.Ci * p.Npad * 4};
    cuuint32_t box[3] = {(cuuint32_t)kChunkC, (cuuint32_t)p.NT, 1};
    int rc = encode_tiled(&mapB, wp, 3, dims, strides, box);
    if (rc) return rc;
  }

  TcArgs a;
  a.N = g.N; a.Cin = g.Ci; a.Ho = g.Ho; a.Wo = g.Wo; a.Co = g.Co;
  a.kh = g.kh; a.kw = g.kw; a.pad = g.pad;
  a.TH = p.TH; a.TW = p.TW; a.tiles_h = p.tiles_h; a.tiles_w = p.tiles_w;
  a.num_tiles = g.N * p.tiles_h * p.tiles_w;
  a.MT = p.MT; a.NT = p.NT; a.stages = p.stages; a.tmem_cols = p.tmem_cols;
  a.ps = g.ps; a.out = out; a.epi = epi;

  static bool attr_set = false;
  if (!attr_set) {
    SRB_CHECK_CUDA(cudaFuncSetAttribute(k_tc_conv, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    attr_set = true;
  }
  dim3 grid((a.num_tiles + a.MT - 1) / a.MT, p.n_tiles_n);
  k_tc_conv<<<grid, kThreads, p.smem, st>>>(mapA, mapB, a);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

// ---- wgrad on tensor cores: not built yet (falls to the fp32 CUDA-core wgrad) ----
bool tc_wgrad_supported(const Geom &, const T4 &, const T4 &) { return false; }
size_t tc_wgrad_ws_bytes(const Geom &) { return 0; }
int tc_conv_wgrad(const Geom &, const T4 &, const T4 &, float *, float *, float, int, void *, size_t, cudaStream_t) {
  set_error("tensor-core wgrad not built");
  return SRB_EUNSUPPORTED;
}

}  // namespace srb
