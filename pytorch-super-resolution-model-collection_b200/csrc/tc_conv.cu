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
#include "tc_common.cuh"

namespace srb {

namespace {

constexpr int kThreads = 192;
constexpr int kChunkC = 32;             // channels per K-block: 32 fp32 = 128 B = one swizzle row
constexpr int kABytes = 128 * 128;      // one A stage tile: 128 pixels x 128 B
constexpr int kMaxSmem = 227 * 1024;

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
  const int chunks = (a.Cin + kChunkC - 1) / kChunkC;  // a ragged last chunk is zero-filled by TMA (A) / packing (B)
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
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====================
    {
      int tn[4], toh[4], tow[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        int tile = tile0 + (t < mt_valid ? t : 0);
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
          if (elect_one()) {
            mbar_expect_tx(&full_bar[st], tx_bytes);
#pragma unroll
            for (int t = 0; t < 4; ++t)
              if (t < mt_valid)
                tma_load_4d(&mapA, &full_bar[st], sa + t * kABytes, c * kChunkC, tow[t] + s, toh[t] + r, tn[t]);
            tma_load_3d(&mapB, &full_bar[st], sa + a.MT * kABytes, c * kChunkC, n0, tap);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    {
      // instruction descriptor: D=f32, A=B=tf32, both K-major, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.NT >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t d_hi = (uint32_t)(make_kmajor_sw128_desc(0) >> 32);
      for (int kb = 0; kb < kblocks; ++kb) {
        const int st = kb % a.stages;
        const uint32_t ph = (uint32_t)(kb / a.stages) & 1u;
        mbar_wait(&full_bar[st], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(smem + (size_t)st * stage_bytes);
        const uint32_t a_lo0 = ((sa >> 4) & 0x3FFF) | (1u << 16);
        const uint32_t b_lo0 = (((sa + a.MT * kABytes) >> 4) & 0x3FFF) | (1u << 16);
        if (elect_one()) {
          const uint32_t acc0 = kb ? 1u : 0u;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            if (t < mt_valid) {
              const uint32_t a_lo = a_lo0 + (uint32_t)(t * (kABytes >> 4));
              const uint32_t tcol = tmem_base + (uint32_t)(t * a.NT);
              // 4 x (K = 8 tf32 = 32 B) inside the 128 B swizzle row
              umma_tf32(tcol, ((uint64_t)d_hi << 32) | (uint64_t)(a_lo + 0), ((uint64_t)d_hi << 32) | (uint64_t)(b_lo0 + 0), idesc, acc0);
              umma_tf32(tcol, ((uint64_t)d_hi << 32) | (uint64_t)(a_lo + 2), ((uint64_t)d_hi << 32) | (uint64_t)(b_lo0 + 2), idesc, 1u);
              umma_tf32(tcol, ((uint64_t)d_hi << 32) | (uint64_t)(a_lo + 4), ((uint64_t)d_hi << 32) | (uint64_t)(b_lo0 + 4), idesc, 1u);
              umma_tf32(tcol, ((uint64_t)d_hi << 32) | (uint64_t)(a_lo + 6), ((uint64_t)d_hi << 32) | (uint64_t)(b_lo0 + 6), idesc, 1u);
            }
          }
          umma_commit(&empty_bar[st]);  // frees this smem stage once the MMAs above have read it
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(accum_bar);  // all accumulators complete
      __syncwarp();
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
        // Store modes (16 accumulator columns = conv channels cbase .. cbase+15 of this pixel):
        //   0: NHWC, no shuffle           -> 4 x float4 at channel cbase + 4q
        //   1: PixelShuffle(4) into NCHW  -> one output channel c = cbase/16; quad q = sub-row i: 4 contiguous j
        //   2: PixelShuffle(2) into NHWC  -> 4 output channels cbase/4 ..+3; quad q = sub-pixel (i,j)
        //   3: anything else              -> scalar stores through ps_offset()
        const bool full = (cbase + 16 <= a.Co);
        const bool lay0 = a.out.sc == 1 && (!a.epi.residual.p || a.epi.residual.sc == 1) &&
                          (!a.epi.preact.p || a.epi.preact.sc == 1);
        const bool lay1 = a.out.sw == 1 && (!a.epi.residual.p || a.epi.residual.sw == 1) &&
                          (!a.epi.preact.p || a.epi.preact.sw == 1);
        int mode = 3;
        if (full && a.ps == 1 && lay0 && (a.Co & 3) == 0) mode = 0;
        else if (full && a.ps == 4 && lay1 && (a.out.sh & 3) == 0) mode = 1;
        else if (full && a.ps == 2 && lay0 && ((a.Co >> 2) & 3) == 0) mode = 2;
        if (mode != 3) {
          float z[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 b = a.epi.bias ? __ldg((const float4 *)(a.epi.bias + cbase + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            z[j] = __uint_as_float(v[j]) + b.x; z[j + 1] = __uint_as_float(v[j + 1]) + b.y;
            z[j + 2] = __uint_as_float(v[j + 2]) + b.z; z[j + 3] = __uint_as_float(v[j + 3]) + b.w;
          }
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            long long off_o, off_r = 0, off_p = 0;
            float4 zq;
            if (mode == 0) {
              zq = make_float4(z[4 * q4], z[4 * q4 + 1], z[4 * q4 + 2], z[4 * q4 + 3]);
              off_o = n * a.out.sn + (long long)oy * a.out.sh + (long long)ox * a.out.sw + cbase + 4 * q4;
              if (a.epi.residual.p)
                off_r = n * a.epi.residual.sn + (long long)oy * a.epi.residual.sh + (long long)ox * a.epi.residual.sw + cbase + 4 * q4;
              if (a.epi.preact.p)
                off_p = n * a.epi.preact.sn + (long long)oy * a.epi.preact.sh + (long long)ox * a.epi.preact.sw + cbase + 4 * q4;
            } else if (mode == 1) {
              zq = make_float4(z[4 * q4], z[4 * q4 + 1], z[4 * q4 + 2], z[4 * q4 + 3]);
              const int c = cbase >> 4;
              const long long yy = (long long)oy * 4 + q4, xx = (long long)ox * 4;
              off_o = n * a.out.sn + c * a.out.sc + yy * a.out.sh + xx;
              if (a.epi.residual.p) off_r = n * a.epi.residual.sn + c * a.epi.residual.sc + yy * a.epi.residual.sh + xx;
              if (a.epi.preact.p) off_p = n * a.epi.preact.sn + c * a.epi.preact.sc + yy * a.epi.preact.sh + xx;
            } else {
              zq = make_float4(z[q4], z[4 + q4], z[8 + q4], z[12 + q4]);
              const int c = cbase >> 2;
              const long long yy = (long long)oy * 2 + (q4 >> 1), xx = (long long)ox * 2 + (q4 & 1);
              off_o = n * a.out.sn + yy * a.out.sh + xx * a.out.sw + c;
              if (a.epi.residual.p) off_r = n * a.epi.residual.sn + yy * a.epi.residual.sh + xx * a.epi.residual.sw + c;
              if (a.epi.preact.p) off_p = n * a.epi.preact.sn + yy * a.epi.preact.sh + xx * a.epi.preact.sw + c;
            }
            if (a.epi.preact.p) *(float4 *)(a.epi.preact.p + off_p) = zq;
            float4 y;
            y.x = apply_act(zq.x, a.epi.act, slope); y.y = apply_act(zq.y, a.epi.act, slope);
            y.z = apply_act(zq.z, a.epi.act, slope); y.w = apply_act(zq.w, a.epi.act, slope);
            if (a.epi.residual.p) {
              const float4 rr = __ldg((const float4 *)(a.epi.residual.p + off_r));
              y.x += rr.x; y.y += rr.y; y.z += rr.z; y.w += rr.w;
            }
            if (a.epi.round_tf32) { y.x = round_tf32(y.x); y.y = round_tf32(y.y); y.z = round_tf32(y.z); y.w = round_tf32(y.w); }
            *(float4 *)(a.out.p + off_o) = y;
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
// Rows n >= N_real and columns k >= K_real (K is padded to a multiple of 32) are zero.
__global__ void k_pack_weights(const float *__restrict__ w, float *__restrict__ out, int Co, int Ci, int kh, int kw,
                               int Npad, int Kpad, int flip) {
  const int N = flip ? Ci : Co, K = flip ? Co : Ci;
  const long long total = (long long)kh * kw * Npad * Kpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int k = (int)(i % Kpad);
    long long q = i / Kpad;
    int n = (int)(q % Npad);
    int tap = (int)(q / Npad);
    float v = 0.f;
    if (n < N && k < K) {
      int r = tap / kw, s = tap - r * kw;
      if (flip) { r = kh - 1 - r; s = kw - 1 - s; }
      int co = flip ? k : n, ci = flip ? n : k;
      v = round_tf32(__ldg(w + (((long long)co * Ci + ci) * kh + r) * kw + s));
    }
    out[i] = v;
  }
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
  if (g.Ci % 4 != 0 || g.Ci < 16) return false;                        // TMA: 16-byte pixel stride; ragged K chunk is zero-filled
  if (in.sc != 1) return false;                                       // channels_last activations
  if ((in.sw % 4) || (in.sh % 4) || (in.sn % 4)) return false;       // TMA strides: multiples of 16 B
  if (((uintptr_t)in.p) & 15) return false;
  if (g.kh > 16 || g.kw > 16) return false;
  if ((long long)g.N * g.Hi * g.Wi * g.Ci >= (1LL << 40)) return false;
  Plan p;
  return make_plan(g, &p);
}

size_t tc_conv_ws_bytes(const Geom &g) {
  int a = round_up(g.Co, 32), b = round_up(g.Ci, 32);
  int m = a > b ? a : b;
  return (size_t)g.kh * g.kw * m * m * sizeof(float) + 512;
}

int tc_conv_gather(const Geom &g, const T4 &in, const float *w, bool flip_transpose, const T4 &out, const Epi &epi,
                   void *ws, size_t ws_bytes, cudaStream_t st) {
  Plan p;
  SRB_REQUIRE(make_plan(g, &p), SRB_EUNSUPPORTED, "tc_conv: no tile plan");
  const int Kpad = round_up(g.Ci, kChunkC);
  size_t need = (size_t)g.kh * g.kw * p.Npad * Kpad * sizeof(float);
  uintptr_t wsp = ((uintptr_t)ws + 255) & ~(uintptr_t)255;
  SRB_REQUIRE(ws && wsp + need <= (uintptr_t)ws + ws_bytes, SRB_EWORKSPACE, "tc_conv workspace: need %zu bytes", need + 256);
  float *wp = (float *)wsp;

  // 1. pack weights.  In dgrad orientation the original filter is (Co_orig = g.Ci ... ) -- see caller:
  //    g here is already the *gather* geometry, so the stored filter is w[co_orig][ci_orig] with
  //    co_orig = g.Ci (K side) and ci_orig = g.Co (N side) when flip_transpose is set.
  {
    long long total = (long long)g.kh * g.kw * p.Npad * Kpad;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (!flip_transpose)
      k_pack_weights<<<blocks, 256, 0, st>>>(w, wp, g.Co, g.Ci, g.kh, g.kw, p.Npad, Kpad, 0);
    else
      k_pack_weights<<<blocks, 256, 0, st>>>(w, wp, /*Co_orig=*/g.Ci, /*Ci_orig=*/g.Co, g.kh, g.kw, p.Npad, Kpad, 1);
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
    cuuint64_t dims[3] = {(cuuint64_t)Kpad, (cuuint64_t)p.Npad, (cuuint64_t)(g.kh * g.kw)};
    cuuint64_t strides[2] = {(cuuint64_t)Kpad * 4, (cuuint64_t)Kpad * p.Npad * 4};
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

}  // namespace srb
