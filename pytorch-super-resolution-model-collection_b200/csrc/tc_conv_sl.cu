// "Slot-linear" tcgen05 implicit-GEMM convolution for sm_100a (stride-1 fprop, and dgrad as a flipped fprop).
//
//   D[m, n] = sum_{r,s,c} X[pixel(m) + (r,s), c] * W[n, c, r, s]        m = output pixel slot, n = output channel
//
// The input halo of a band of TH x TW output pixels -- (TH+kh-1) x BW pixel "slots", BW = TW+kw-1 -- is fetched
// ONCE per 32-channel chunk by a single TMA box load into shared memory (out-of-image slots are zero-filled by the
// TMA unit = the conv padding).  Output pixel (ty,tx) is slot q = ty*BW + tx of the flattened band, and filter tap
// (r,s) is nothing but the same shared-memory tile viewed (r*BW + s) slots later: the A operand of every tap is a
// SHIFTED VIEW, expressed purely through the start address of the tcgen05 shared-memory descriptor (the 128B swizzle
// is a function of the absolute smem address, so any 128-byte shift stays consistent with what TMA wrote --
// tools/probe_desc.cu).  An M-tile is 128 consecutive slots; the kw-1 wrap-around slots per row compute garbage that
// the epilogue never stores.  Compared with per-tap im2col loads this cuts L2->SMEM traffic by ~kh*kw.
//
// Two operand flavours share the kernel:
//   generic (Cin % 4 == 0, NHWC): slot = 128 B = 32 channels, K-major SWIZZLE_128B; one K-block = (chunk, tap) = 4 MMAs
//            of K = 8; weights stream through a TMA ring as [chunk][tap][Npad][32].
//   c4      (Cin <= 4): the image is packed to NHWC4 (16 B / pixel); slot = 16 B.  With the NO-swizzle K-major layout,
//            leading-byte-offset 16 B and stride-byte-offset 128 B, row m of the operand is the 32 bytes starting at
//            slot m: an OVERLAPPING (Toeplitz) view -- one K = 8 MMA covers two horizontal taps x 4 channels straight
//            from the raw image tile.  All weights ([kb][N/8][2][8][4]) stay resident in shared memory.
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issue (warp-uniform loops, one elected lane issues),
// warps 2..5 = epilogue (tcgen05.ld -> bias -> act -> +residual -> tf32 round -> store at the pixel-shuffled address).
// CTAs are sized for two per SM (<= 112 KB smem, <= 256 TMEM columns) so one CTA's epilogue overlaps the other's MMAs.
#include "tc_common.cuh"
#include <string.h>
#include <mutex>
#include <vector>

namespace srb {

namespace {

constexpr int kThreads = 320;      // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two CTAs per SM)
constexpr int kThreadsWide = 576;  // the same with 16 epilogue warps: one CTA per SM, four epilogue warps per TMEM lane quadrant.
                                   // The store-bound layers (Cin <= 4 edges writing 64 channels, bf16 bodies) were limited by the
                                   // number of stores in flight: VDSR 3->64 at 128^2 ran at 1.6 TB/s with 8 epilogue warps per SM
constexpr int kRsEpiPerQuadFwd = 4;  // == kRsEpiPerQuad of k_conv_rs (declared further down)
constexpr int kMaxLossCtas = 1024;  // fused loss: per-warp partial sums of at most this many CTAs
constexpr int kMaxBStages = 8;

struct SlArgs {
  int N, Ho, Wo, Co;
  int kh, kw, pad;
  int c4;
  int TH, TW, BW, BH;
  int bands_h, bands_w;
  int MTB, NT;
  int chunks;        // A loads per band: generic ceil(Cin/32); c4: 1
  int spairs;        // c4: ceil(kw/2)
  int a_bufs, a_buf_bytes, a_tx_bytes;
  int b_stages, b_stage_bytes;  // c4: one stage holding every K-block of this N tile
  int b_resident;               // generic: b_stages == K-blocks, every weight tile is loaded once per CTA and stays
  int t_bufs, tmem_cols;  // accumulator buffers in TMEM (2: epilogue of band i overlaps the MMAs of band i+1)
  int ps;
  int in_bf16;       // operands are bf16 (slot = 128 B = 64 channels, kind::f16 MMAs of K = 16); generic flavour only
  int out_bf16;      // out / residual / preact are bf16 tensors
  int chunk_elems;   // channels per A chunk: 32 (tf32) or 64 (bf16)
  int v8h;           // bf16 outputs: rows are 32-byte aligned (256-bit accesses of 16 bf16 channels)
  int cl;            // thread-block cluster size along grid.x (1 or 2).  2: the streamed weight ring is SHARED by the CTA pair --
                     // each CTA fetches half of every weight stage and TMA-multicasts it into both CTAs' shared memory, so the
                     // L2 -> SM weight traffic per SM halves; the pair walks its bands in lock-step (same K-block sequence)
  int in_ps;         // > 1: the input tensor is PixelShuffle_r of the logical input (dgrad of a PSBlock conv): chunk c is
  int in_cpb;        //      sub-pixel phase c / in_cpb, channel block c % in_cpb, fetched through a stride-r TMA traversal
  int kb_valid;      // K-blocks (chunk, tap) that exist: chunks * taps minus the taps the phase masks exclude
  int pad_w;         // horizontal padding (== pad unless a phase launch of a strided / transposed conv says otherwise)
  int rs;            // row-stacked kernel (k_conv_rs): an M tile is ONE output row of up to 128 / BW images side by side --
                     // slot q of the tile is pixel q % BW of image n + q / BW (the epilogues read this mapping, nothing else)
  unsigned short phase_mask[16];  // in_ps > 1: bit (r*kw + s) set = tap (r, s) of that input phase exists (others are skipped)
  const float *wpack;  // c4: packed weights (bulk-copied); generic: unused (TMA map)
  int v8;              // out / residual / preact / mask rows are 32-byte aligned: mode-0 epilogue uses 256-bit global accesses
  int dbg;             // debug knobs (srb_debug_set_flags): 1 = epilogue does nothing, 2 = A tiles are loaded only once per buffer
  long long *trace;    // debug (srb_debug_set_trace): 8 timestamps per CTA, null in production
  T4 out;
  Epi epi;
};

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define SL_TRACE(slot)                                                                           \
  do {                                                                                           \
    if (a.trace) a.trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + (slot)] = gtime();  \
  } while (0)

__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// tf32 round-to-nearest (ties away) on the bit pattern: 2 integer ops instead of cvt.rna's 4-instruction expansion
// (the Inf guard is dropped: +-Inf would become NaN, which only matters for a diverged run).
__device__ __forceinline__ float round_tf32_fast(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr));
  return r;
}

// 256-bit global accesses (sm_100): one full 32-byte sector per thread and instruction -- half the LSU wavefronts of float4.
// Macros on purpose: the operands are elements of register-resident arrays (taking their address would demote them to local memory).
#define STG256(ptr, v, o)                                                                                                  \
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "f"(v[(o)]), "f"(v[(o) + 1]),       \
               "f"(v[(o) + 2]), "f"(v[(o) + 3]), "f"(v[(o) + 4]), "f"(v[(o) + 5]), "f"(v[(o) + 6]), "f"(v[(o) + 7])         \
               : "memory")
#define LDG256(ptr, v, o)                                                                                                  \
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"                                              \
               : "=f"(v[(o)]), "=f"(v[(o) + 1]), "=f"(v[(o) + 2]), "=f"(v[(o) + 3]), "=f"(v[(o) + 4]), "=f"(v[(o) + 5]),  \
                 "=f"(v[(o) + 6]), "=f"(v[(o) + 7])                                                                        \
               : "l"(ptr))

// One epilogue warp's share of the band: work items (M-tile, 16-column group), item = half, half+2, ...
//   MODE 0: NHWC, no shuffle           -> 4 x float4 at channel cbase + 4q
//   MODE 1: PixelShuffle(4) into NCHW  -> one output channel c = cbase/16; quad q = sub-row i: 4 contiguous j
//   MODE 2: PixelShuffle(2) into NHWC  -> 4 output channels cbase/4 ..+3; quad q = sub-pixel (i,j)
//   MODE 3: anything else              -> scalar stores through ps_offset()
// Modes 0..2 need Co % 16 == 0.  EXTRA = a residual and/or pre-activation tensor is present; the lean variant
// (plain conv + bias + act) is ~100 instructions per item, which matters: the epilogue is issue-bound.
template <int MODE, bool EXTRA>
__device__ __forceinline__ void epilogue_items(const SlArgs &a, uint32_t trow, uint32_t bias_saddr, int half, int mtb, int m,
                                               int n, int n0, int oy0, int ox0, int rows_valid, int cols_valid, int istep = 2) {
  const int act = a.epi.act;
  const float slope = (act == SRB_ACT_PRELU) ? __ldg(a.epi.alpha) : a.epi.slope;
  const bool has_bias = a.epi.bias != nullptr;
  const bool rnd = a.epi.round_tf32 != 0;
  const bool has_res = EXTRA && a.epi.residual.p != nullptr, has_pre = EXTRA && a.epi.preact.p != nullptr;
  const bool has_mask = EXTRA && a.epi.mask.p != nullptr;  // MODE 0 / 3 only (dgrad never shuffles)
  const int ngroups = a.NT >> 4;
  int t_cur = -1, oy = 0, ox = 0;
  bool pix_ok = false;
  float *po = nullptr, *pp = nullptr;
  const float *pr = nullptr, *pm = nullptr;
  int pixw = 0;  // first bit word of this thread's pixel (host guarantees N*Ho*Wo*Co/16 < 2^31)
  int t_nx = half / ngroups, g_nx = half - t_nx * ngroups;  // (tile, column group) of the next item, advanced without a division
#pragma unroll 1
  for (int item = half; item < mtb * ngroups; item += istep) {
    const int t = t_nx, j0 = g_nx << 4;
    g_nx += istep;
    while (g_nx >= ngroups) { g_nx -= ngroups; ++t_nx; }
    if (t != t_cur) {
      t_cur = t;
      const int q = t * 128 + m;  // slot of this thread's accumulator row
      const int ty = q / a.BW, tx = q - ty * a.BW;
      if (a.rs) { n += ty; oy = oy0; }  // row-stacked tiles (mtb == 1): ty counts images, rows_valid = images present
      else oy = oy0 + ty;
      ox = ox0 + tx;
      pix_ok = (ty < rows_valid) && (tx < cols_valid);
      pixw = ((n * a.Ho + oy) * a.Wo + ox) * (a.Co >> 4);
      if (MODE != 3) {
        const long long yy = (long long)oy * a.ps, xx = (long long)ox * a.ps;
        po = a.out.p + (n * a.out.sn + yy * a.out.sh + xx * a.out.sw);
        if (EXTRA) {
          pr = a.epi.residual.p + (n * a.epi.residual.sn + yy * a.epi.residual.sh + xx * a.epi.residual.sw);
          pp = a.epi.preact.p + (n * a.epi.preact.sn + yy * a.epi.preact.sh + xx * a.epi.preact.sw);
          pm = a.epi.mask.p + (n * a.epi.mask.sn + yy * a.epi.mask.sh + xx * a.epi.mask.sw);
        }
      }
    }
    const int cbase = n0 + j0;
    if (cbase >= a.Co) continue;  // warp-uniform
    uint32_t v[16];
    tmem_ld16(trow + (uint32_t)(t * a.NT + j0), v);
    if (!pix_ok) continue;
    float z[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) z[j] = __uint_as_float(v[j]);
    if (has_bias) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 b = lds128(bias_saddr + (uint32_t)(j0 + j) * 4u);
        z[j] += b.x; z[j + 1] += b.y; z[j + 2] += b.z; z[j + 3] += b.w;
      }
    }
    if (MODE != 3 && a.epi.bits_out) {
      uint32_t mbits = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) mbits |= (z[j] > 0.f ? 1u : 0u) << j;
      a.epi.bits_out[pixw + (cbase >> 4)] = (unsigned short)mbits;
    }
    if (MODE == 3) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int co = cbase + j;
        if (co < a.Co) {
          if (a.epi.preact.p) st_any(a.epi.preact, ps_offset(a.epi.preact, a.ps, n, co, oy, ox), z[j]);
          float y = act == SRB_ACT_NONE ? z[j] : (act == SRB_ACT_RELU ? fmaxf(z[j], 0.f) : (z[j] > 0.f ? z[j] : z[j] * slope));
          if (a.epi.residual.p) y += ld_any(a.epi.residual, ps_offset(a.epi.residual, a.ps, n, co, oy, ox));
          if (a.epi.mask.p && !(ld_any(a.epi.mask, ps_offset(a.epi.mask, a.ps, n, co, oy, ox)) > 0.f)) y = 0.f;
          if (rnd) y = round_tf32_fast(y);
          st_any(a.out, ps_offset(a.out, a.ps, n, co, oy, ox), y);
        }
      }
      continue;
    }
    // group base pointers and quad strides (floats)
    float *pg, *ppg = nullptr;
    const float *prg = nullptr;
    long long qs_o, qs_o2 = 0, qs_r = 0, qs_r2 = 0, qs_p = 0, qs_p2 = 0;
    if (MODE == 0) {
      pg = po + cbase; qs_o = 4;
      if (EXTRA) { prg = pr + cbase; ppg = pp + cbase; qs_r = qs_p = 4; }
    } else if (MODE == 1) {
      const int c = cbase >> 4;
      pg = po + c * a.out.sc; qs_o = a.out.sh;
      if (EXTRA) { prg = pr + c * a.epi.residual.sc; ppg = pp + c * a.epi.preact.sc; qs_r = a.epi.residual.sh; qs_p = a.epi.preact.sh; }
    } else {
      const int c = cbase >> 2;
      pg = po + c; qs_o = a.out.sw; qs_o2 = a.out.sh;
      if (EXTRA) { prg = pr + c; ppg = pp + c; qs_r = a.epi.residual.sw; qs_r2 = a.epi.residual.sh; qs_p = a.epi.preact.sw; qs_p2 = a.epi.preact.sh; }
    }
    if (MODE == 0 && has_pre && a.v8) {
      STG256(ppg, z, 0);
      STG256(ppg + 8, z, 8);
    } else if (has_pre) {
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const float4 zq = MODE == 2 ? make_float4(z[q4], z[4 + q4], z[8 + q4], z[12 + q4])
                                    : make_float4(z[4 * q4], z[4 * q4 + 1], z[4 * q4 + 2], z[4 * q4 + 3]);
        *(float4 *)(ppg + (MODE == 2 ? (q4 >> 1) * qs_p2 + (q4 & 1) * qs_p : q4 * qs_p)) = zq;
      }
    }
    if (act == SRB_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 16; ++j) z[j] = fmaxf(z[j], 0.f);
    } else if (act != SRB_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 16; ++j) z[j] = z[j] > 0.f ? z[j] : z[j] * slope;
    }
    if (MODE == 0 && has_res && a.v8) {
      float rr[16];
      LDG256(prg, rr, 0);
      LDG256(prg + 8, rr, 8);
#pragma unroll
      for (int j = 0; j < 16; ++j) z[j] += rr[j];
    } else if (has_res) {
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const float4 rr = __ldg((const float4 *)(prg + (MODE == 2 ? (q4 >> 1) * qs_r2 + (q4 & 1) * qs_r : q4 * qs_r)));
        if (MODE == 2) { z[q4] += rr.x; z[4 + q4] += rr.y; z[8 + q4] += rr.z; z[12 + q4] += rr.w; }
        else { z[4 * q4] += rr.x; z[4 * q4 + 1] += rr.y; z[4 * q4 + 2] += rr.z; z[4 * q4 + 3] += rr.w; }
      }
    }
    if (a.epi.bits_in) {
      const uint32_t mbits = __ldg(a.epi.bits_in + pixw + (cbase >> 4));
#pragma unroll
      for (int j = 0; j < 16; ++j) z[j] = ((mbits >> j) & 1u) ? z[j] : 0.f;
    }
    if (MODE == 0 && has_mask) {
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const float4 mm = __ldg((const float4 *)(pm + cbase + 4 * q4));
        z[4 * q4] = mm.x > 0.f ? z[4 * q4] : 0.f; z[4 * q4 + 1] = mm.y > 0.f ? z[4 * q4 + 1] : 0.f;
        z[4 * q4 + 2] = mm.z > 0.f ? z[4 * q4 + 2] : 0.f; z[4 * q4 + 3] = mm.w > 0.f ? z[4 * q4 + 3] : 0.f;
      }
    }
    if (rnd) {
#pragma unroll
      for (int j = 0; j < 16; ++j) z[j] = round_tf32_fast(z[j]);
    }
    if (MODE == 0 && a.v8) {
      STG256(pg, z, 0);
      STG256(pg + 8, z, 8);
      continue;
    }
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const float4 y = MODE == 2 ? make_float4(z[q4], z[4 + q4], z[8 + q4], z[12 + q4])
                                 : make_float4(z[4 * q4], z[4 * q4 + 1], z[4 * q4 + 2], z[4 * q4 + 3]);
      *(float4 *)(pg + (MODE == 2 ? (q4 >> 1) * qs_o2 + (q4 & 1) * qs_o : q4 * qs_o)) = y;
    }
  }
}


// one pixel byte as torchvision's ToTensor maps it (byte / 255, correctly rounded; k_image_to_tensor computes the same)
__device__ __forceinline__ float u8_unit(const unsigned char *p) { return byte_over_255((unsigned int)__ldg(p)); }

// ---- loss-fused epilogue of the network's last conv (no activation, no residual; fp32 outputs) --------------------------------
//   MODE 1: PixelShuffle(4) into NCHW (ESPCN): y/target addressed like epilogue_items MODE 1; the gradient is written
//           un-shuffled -- 16 consecutive conv channels of this pixel = 64 contiguous bytes of the (N,Ho,Wo,Co) tensor
//   MODE 3: anything with ps == 1: scalar accesses, gradient in y's layout
template <int MODE>
__device__ __forceinline__ void epilogue_items_loss(const SlArgs &a, uint32_t trow, uint32_t bias_saddr, int half, int mtb, int m,
                                                    int n_in, int n0, int oy0, int ox0, int rows_valid, int cols_valid, float &lsum,
                                                    int istep = 2) {
  const bool has_bias = a.epi.bias != nullptr;
  const bool l1 = a.epi.loss_kind == 2;
  const bool rnd = a.epi.round_tf32 != 0;  // here: round the GRADIENT (it feeds the tensor-core dgrad / wgrad)
  const float coef = a.epi.loss_coef * (l1 ? 1.f : 2.f);
  const int ngroups = a.NT >> 4;
  int t_nx = half / ngroups, g_nx = half - t_nx * ngroups;  // (tile, column group) of the next item, advanced without a division
#pragma unroll 1
  for (int item = half; item < mtb * ngroups; item += istep) {
    const int t = t_nx, j0 = g_nx << 4;
    g_nx += istep;
    while (g_nx >= ngroups) { g_nx -= ngroups; ++t_nx; }
    const int q = t * 128 + m;
    const int ty = q / a.BW, tx = q - ty * a.BW;
    const int n = a.rs ? n_in + ty : n_in;  // row-stacked tiles: see SlArgs::rs
    const int oy = a.rs ? oy0 : oy0 + ty, ox = ox0 + tx;
    const bool pix_ok = (ty < rows_valid) && (tx < cols_valid);
    const int cbase = n0 + j0;
    if (cbase >= a.Co) continue;  // warp-uniform
    uint32_t v[16];
    tmem_ld16(trow + (uint32_t)(t * a.NT + j0), v);
    if (!pix_ok) continue;
    float z[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) z[j] = __uint_as_float(v[j]);
    if (has_bias) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 b = lds128(bias_saddr + (uint32_t)(j0 + j) * 4u);
        z[j] += b.x; z[j + 1] += b.y; z[j + 2] += b.z; z[j + 3] += b.w;
      }
    }
    if (MODE == 1) {
      const int c = cbase >> 4;
      const long long yy = (long long)oy * 4, xx = (long long)ox * 4;
      const long long toff = n * a.epi.target.sn + c * a.epi.target.sc + yy * a.epi.target.sh + xx * a.epi.target.sw;
      const float *pt = a.epi.target.p + toff;
      const unsigned char *pb = (const unsigned char *)a.epi.target.p + toff;  // SRB_U8 target: the strides count bytes
      const bool tu8 = a.epi.target.dt == SRB_U8;
      if (a.rs && oy + kRsEpiPerQuadFwd < a.Ho) {
        // row-stacked kernel: this warp's next output row is oy + kRsEpiPerQuad.  Its target rows are pulled into L2 now: the four
        // dependent 16-byte loads per item below would otherwise expose the full DRAM latency with ~16 KB in flight per SM.
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          if (tu8) asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + (4 * kRsEpiPerQuadFwd + q4) * a.epi.target.sh));
          else asm volatile("prefetch.global.L2 [%0];" ::"l"(pt + (4 * kRsEpiPerQuadFwd + q4) * a.epi.target.sh));
        }
      }
      float *po = a.out.p ? a.out.p + (n * a.out.sn + c * a.out.sc + yy * a.out.sh + xx * a.out.sw) : nullptr;
      float g[16];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        float4 tt;
        if (tu8) {  // raw image bytes (any pixel stride: HWC as decoded, or planar); t = byte / 255 as ToTensor computes it
          const unsigned char *pr = pb + q4 * a.epi.target.sh;
          tt = make_float4(u8_unit(pr), u8_unit(pr + a.epi.target.sw), u8_unit(pr + 2 * a.epi.target.sw), u8_unit(pr + 3 * a.epi.target.sw));
        } else {
          tt = __ldg((const float4 *)(pt + q4 * a.epi.target.sh));
        }
        const float d0 = z[4 * q4] - tt.x, d1 = z[4 * q4 + 1] - tt.y, d2 = z[4 * q4 + 2] - tt.z, d3 = z[4 * q4 + 3] - tt.w;
        if (l1) {
          lsum += (fabsf(d0) + fabsf(d1)) + (fabsf(d2) + fabsf(d3));
          g[4 * q4] = d0 > 0.f ? coef : (d0 < 0.f ? -coef : 0.f); g[4 * q4 + 1] = d1 > 0.f ? coef : (d1 < 0.f ? -coef : 0.f);
          g[4 * q4 + 2] = d2 > 0.f ? coef : (d2 < 0.f ? -coef : 0.f); g[4 * q4 + 3] = d3 > 0.f ? coef : (d3 < 0.f ? -coef : 0.f);
        } else {
          lsum += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
          g[4 * q4] = coef * d0; g[4 * q4 + 1] = coef * d1; g[4 * q4 + 2] = coef * d2; g[4 * q4 + 3] = coef * d3;
        }
        if (po) *(float4 *)(po + q4 * a.out.sh) = make_float4(z[4 * q4], z[4 * q4 + 1], z[4 * q4 + 2], z[4 * q4 + 3]);
      }
      if (rnd) {
#pragma unroll
        for (int j = 0; j < 16; ++j) g[j] = round_tf32_fast(g[j]);
      }
      float *pd = a.epi.dz_unshuf + ((long long)(n * a.Ho + oy) * a.Wo + ox) * a.Co + cbase;
      STG256(pd, g, 0);
      STG256(pd + 8, g, 8);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int co = cbase + j;
        if (co < a.Co) {
          const long long toff = ps_offset(a.epi.target, 1, n, co, oy, ox);
          const float tt = a.epi.target.dt == SRB_U8 ? u8_unit((const unsigned char *)a.epi.target.p + toff) : __ldg(a.epi.target.p + toff);
          const float d = z[j] - tt;
          float g;
          if (l1) { lsum += fabsf(d); g = d > 0.f ? coef : (d < 0.f ? -coef : 0.f); }
          else    { lsum += d * d; g = coef * d; }
          if (rnd) g = round_tf32_fast(g);
          if (a.out.p) a.out.p[ps_offset(a.out, 1, n, co, oy, ox)] = z[j];
          a.epi.dz.p[ps_offset(a.epi.dz, 1, n, co, oy, ox)] = g;
        }
      }
    }
  }
}

// ---- bf16-output epilogue (out / residual / preact are bf16 NHWC tensors; accumulators, bias and the math stay fp32) ----
//   MODE 0: NHWC, no shuffle: 16 channels = 32 B per thread and item -> one 256-bit store
//   MODE 2: PixelShuffle(2) into NHWC: an item is 32 accumulator columns = 8 output channels x 4 sub-pixels -> four 16-B stores
#define STG256U(ptr, v, o)                                                                                                 \
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(v[(o)]), "r"(v[(o) + 1]),       \
               "r"(v[(o) + 2]), "r"(v[(o) + 3]), "r"(v[(o) + 4]), "r"(v[(o) + 5]), "r"(v[(o) + 6]), "r"(v[(o) + 7])         \
               : "memory")
#define LDG256U(ptr, v, o)                                                                                                 \
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"                                              \
               : "=r"(v[(o)]), "=r"(v[(o) + 1]), "=r"(v[(o) + 2]), "=r"(v[(o) + 3]), "=r"(v[(o) + 4]), "=r"(v[(o) + 5]),  \
                 "=r"(v[(o) + 6]), "=r"(v[(o) + 7])                                                                        \
               : "l"(ptr))

template <int MODE>
__device__ __forceinline__ void epilogue_items_h(const SlArgs &a, uint32_t trow, uint32_t bias_saddr, int half, int mtb, int m,
                                                 int n, int n0, int oy0, int ox0, int rows_valid, int cols_valid, int istep = 2) {
  typedef unsigned short bf;
  const int act = a.epi.act;
  const float slope = (act == SRB_ACT_PRELU) ? __ldg(a.epi.alpha) : a.epi.slope;
  const bool has_bias = a.epi.bias != nullptr;
  const bool has_res = a.epi.residual.p != nullptr, has_pre = a.epi.preact.p != nullptr;
  constexpr int COLS = MODE == 2 ? 32 : 16;
  const int ngroups = a.NT / COLS;
  int t_cur = -1, oy = 0, ox = 0;
  bool pix_ok = false;
  bf *po = nullptr, *pp = nullptr;
  const bf *pr = nullptr;
  int pixw = 0;
  int t_nx = half / ngroups, g_nx = half - t_nx * ngroups;  // (tile, column group) of the next item, advanced without a division
#pragma unroll 1
  for (int item = half; item < mtb * ngroups; item += istep) {
    const int t = t_nx, j0 = g_nx * COLS;
    g_nx += istep;
    while (g_nx >= ngroups) { g_nx -= ngroups; ++t_nx; }
    if (t != t_cur) {
      t_cur = t;
      const int q = t * 128 + m;
      const int ty = q / a.BW, tx = q - ty * a.BW;
      oy = oy0 + ty; ox = ox0 + tx;
      pix_ok = (ty < rows_valid) && (tx < cols_valid);
      pixw = ((n * a.Ho + oy) * a.Wo + ox) * (a.Co >> 4);
      const long long yy = (long long)oy * a.ps, xx = (long long)ox * a.ps;
      po = (bf *)a.out.p + (n * a.out.sn + yy * a.out.sh + xx * a.out.sw);
      pr = (const bf *)a.epi.residual.p + (n * a.epi.residual.sn + yy * a.epi.residual.sh + xx * a.epi.residual.sw);
      pp = (bf *)a.epi.preact.p + (n * a.epi.preact.sn + yy * a.epi.preact.sh + xx * a.epi.preact.sw);
    }
    const int cbase = n0 + j0;
    if (cbase >= a.Co) continue;  // warp-uniform
    float z[COLS];
    {
      uint32_t v[16];
      tmem_ld16(trow + (uint32_t)(t * a.NT + j0), v);
#pragma unroll
      for (int j = 0; j < 16; ++j) z[j] = __uint_as_float(v[j]);
      if (MODE == 2) {
        tmem_ld16(trow + (uint32_t)(t * a.NT + j0 + 16), v);
#pragma unroll
        for (int j = 0; j < 16; ++j) z[16 + j] = __uint_as_float(v[j]);
      }
    }
    if (!pix_ok) continue;
    if (has_bias) {
#pragma unroll
      for (int j = 0; j < COLS; j += 4) {
        const float4 b = lds128(bias_saddr + (uint32_t)(j0 + j) * 4u);
        z[j] += b.x; z[j + 1] += b.y; z[j + 2] += b.z; z[j + 3] += b.w;
      }
    }
    if (MODE == 0 && a.epi.bits_out) {
      uint32_t mbits = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) mbits |= (z[j] > 0.f ? 1u : 0u) << j;
      a.epi.bits_out[pixw + (cbase >> 4)] = (unsigned short)mbits;
    }
    if (MODE == 0) {
      bf *pg = po + cbase;
      if (has_pre) {
        uint32_t u[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) u[j] = pack_bf16x2(z[2 * j], z[2 * j + 1]);
        STG256U(pp + cbase, u, 0);
      }
      if (act == SRB_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 16; ++j) z[j] = fmaxf(z[j], 0.f);
      } else if (act != SRB_ACT_NONE) {
#pragma unroll
        for (int j = 0; j < 16; ++j) z[j] = z[j] > 0.f ? z[j] : z[j] * slope;
      }
      if (has_res) {
        uint32_t u[8];
        LDG256U(pr + cbase, u, 0);
#pragma unroll
        for (int j = 0; j < 8; ++j) { z[2 * j] += bf16_lo(u[j]); z[2 * j + 1] += bf16_hi(u[j]); }
      }
      if (a.epi.bits_in) {
        const uint32_t mbits = __ldg(a.epi.bits_in + pixw + (cbase >> 4));
#pragma unroll
        for (int j = 0; j < 16; ++j) z[j] = ((mbits >> j) & 1u) ? z[j] : 0.f;
      }
      uint32_t u[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) u[j] = pack_bf16x2(z[2 * j], z[2 * j + 1]);
      STG256U(pg, u, 0);
    } else {
      // conv channel cbase + 4*cc + ij -> output channel c0 + cc (c0 = cbase / 4), sub-pixel ij = 2*i + j
      const int c0 = cbase >> 2;
#pragma unroll
      for (int ij = 0; ij < 4; ++ij) {
        const long long off = (long long)(ij >> 1) * a.out.sh + (long long)(ij & 1) * a.out.sw + c0;
        float y[8];
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) y[cc] = z[4 * cc + ij];
        if (has_pre) {
          const long long offp = (long long)(ij >> 1) * a.epi.preact.sh + (long long)(ij & 1) * a.epi.preact.sw + c0;
          *(uint4 *)(pp + offp) = make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]),
                                             pack_bf16x2(y[6], y[7]));
        }
        if (act == SRB_ACT_RELU) {
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) y[cc] = fmaxf(y[cc], 0.f);
        } else if (act != SRB_ACT_NONE) {
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) y[cc] = y[cc] > 0.f ? y[cc] : y[cc] * slope;
        }
        if (has_res) {
          const long long offr = (long long)(ij >> 1) * a.epi.residual.sh + (long long)(ij & 1) * a.epi.residual.sw + c0;
          const uint4 r4 = __ldg((const uint4 *)(pr + offr));
          y[0] += bf16_lo(r4.x); y[1] += bf16_hi(r4.x); y[2] += bf16_lo(r4.y); y[3] += bf16_hi(r4.y);
          y[4] += bf16_lo(r4.z); y[5] += bf16_hi(r4.z); y[6] += bf16_lo(r4.w); y[7] += bf16_hi(r4.w);
        }
        *(uint4 *)(po + off) = make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]),
                                          pack_bf16x2(y[6], y[7]));
      }
    }
  }
}

// Ring positions of the MMA warp (buffer index + phase bit each; no divisions in the issue loop).
struct MmaState {
  uint32_t a_buf, a_phase, b_st, b_phase, t_buf, t_phase;
};

// (Called by the ONE elected thread of the MMA warp: every wait, MMA and commit of the band is issued by it.)
// All MMAs of one band, generic operands: for every 32-channel chunk and filter tap, K = 4 x 8 over MTB_ M-tiles.
// K-step outer / M-tile inner so that consecutive MMAs write different accumulators.
template <int MTB_, bool BF>
__device__ __forceinline__ void mma_band_generic(const SlArgs &a, MmaState &ms, uint64_t *a_full, uint64_t *a_empty,
                                                 uint64_t *b_full, uint64_t *b_empty, uint32_t a_base, uint32_t b_base,
                                                 uint32_t tacc, uint32_t idesc) {
  // K-major SWIZZLE_128B: SBO 1024 B (8 slots), layout type 2; LBO field = 1 (unused)
  const uint64_t d_hi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
  const uint32_t lbo = 1u << 16;
  const uint32_t row_step = (uint32_t)(a.BW - a.kw) * 8u;  // desc units (16 B) from the end of one filter row to the next
  const uint32_t b_stage16 = (uint32_t)a.b_stage_bytes >> 4;
  uint32_t tcol[MTB_];
#pragma unroll
  for (int j = 0; j < MTB_; ++j) tcol[j] = tacc + (uint32_t)(j * a.NT);
  uint32_t b_res = ((b_base >> 4) & 0x3FFF) | lbo;  // resident weights: K-block kb lives at stage kb
  uint32_t first_kb = 0;                             // 0 for the very first K-block of the band: overwrite the accumulators
  for (int c = 0; c < a.chunks; ++c) {
    mbar_wait(&a_full[ms.a_buf], ms.a_phase);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t a_tap = (((a_base + ms.a_buf * (uint32_t)a.a_buf_bytes) >> 4) & 0x3FFF) | lbo;
    // phase masks only exist for interleaved inputs (<= 16 taps); plain convs have up to 16 x 16 taps and no mask
    const bool masked = a.in_ps > 1;
    const uint32_t tmask = masked ? (uint32_t)a.phase_mask[c / a.in_cpb] : 0u;
    for (int r = 0; r < a.kh; ++r) {
      for (int s = 0; s < a.kw; ++s) {
        uint32_t b_lo;
        if (masked && !((tmask >> (r * a.kw + s)) & 1u)) {  // this tap does not exist for this input phase (strided conv)
          if (a.b_resident) b_res += b_stage16;
          a_tap += 8u;
          continue;
        }
        if (a.b_resident) {
          b_lo = b_res;
          b_res += b_stage16;
        } else {
          mbar_wait(&b_full[ms.b_st], ms.b_phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          b_lo = (((b_base >> 4) + ms.b_st * b_stage16) & 0x3FFF) | lbo;
        }
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
#pragma unroll
          for (int j = 0; j < MTB_; ++j) {
            // one K step = 32 bytes of every slot: 8 tf32 or 16 bf16 channels -- the descriptor arithmetic is byte-identical
            if (BF) umma_bf16_ss(tcol[j], d_hi | (uint64_t)(a_tap + (uint32_t)(j * 1024 + 2 * k4)), d_hi | (uint64_t)(b_lo + 2u * k4), idesc,
                                 k4 ? 1u : first_kb);
            else umma_tf32_ss(tcol[j], d_hi | (uint64_t)(a_tap + (uint32_t)(j * 1024 + 2 * k4)), d_hi | (uint64_t)(b_lo + 2u * k4), idesc,
                              k4 ? 1u : first_kb);
          }
        }
        if (!a.b_resident) {  // frees this weight stage (in both CTAs of a sharing pair) once its MMAs have read it
          if (a.cl > 1) umma_commit_arrive_mc(&b_empty[ms.b_st], 3);
          else umma_commit_arrive(&b_empty[ms.b_st]);
        }
        first_kb = 1u;
        a_tap += 8u;  // next tap in the row: one slot (128 B) later
        if (!a.b_resident && ++ms.b_st == (uint32_t)a.b_stages) { ms.b_st = 0; ms.b_phase ^= 1u; }
      }
      a_tap += row_step;
    }
    umma_commit_arrive(&a_empty[ms.a_buf]);
    if (++ms.a_buf == (uint32_t)a.a_bufs) { ms.a_buf = 0; ms.a_phase ^= 1u; }
  }
}

// All MMAs of one band, c4 operands: one K = 8 MMA per (filter row, tap pair) and M-tile.
//   A: no swizzle, LBO 16 B (next 4 floats along K = next pixel), SBO 128 B (next 8 rows = next 8 slots)
//   B: no swizzle canonical [N/8][2][8 rows][16 B]: LBO 128 B, SBO 256 B
template <int MTB_>
__device__ __forceinline__ void mma_band_c4(const SlArgs &a, MmaState &ms, uint64_t *a_full, uint64_t *a_empty, uint32_t a_base,
                                            uint32_t b_base, uint32_t tacc, uint32_t idesc) {
  const uint64_t a_hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32, b_hi = (uint64_t)((256u >> 4) | (1u << 14)) << 32;
  const uint32_t a_lbo = (16u >> 4) << 16, b_lbo = (128u >> 4) << 16;
  uint32_t tcol[MTB_];
#pragma unroll
  for (int j = 0; j < MTB_; ++j) tcol[j] = tacc + (uint32_t)(j * a.NT);
  mbar_wait(&a_full[ms.a_buf], ms.a_phase);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t a_tap = (((a_base + ms.a_buf * (uint32_t)a.a_buf_bytes) >> 4) & 0x3FFF) | a_lbo;
  uint32_t b_lo = ((b_base >> 4) & 0x3FFF) | b_lbo;
  const uint32_t kb16 = (uint32_t)a.NT * 2u;                          // NT * 32 B per K-block, in 16-byte units
  const uint32_t row_step = (uint32_t)(a.BW - 2 * a.spairs);          // slots (16 B) to the next filter row
  uint32_t acc = 0;
  for (int r = 0; r < a.kh; ++r) {
    for (int sp = 0; sp < a.spairs; ++sp) {
#pragma unroll
      for (int j = 0; j < MTB_; ++j)  // 128 slots x 16 B = 2048 B per M-tile
        umma_tf32_ss(tcol[j], a_hi | (uint64_t)(a_tap + (uint32_t)(j * 128)), b_hi | (uint64_t)b_lo, idesc, acc);
      acc = 1u;
      a_tap += 2u;   // two pixels (32 B) to the right
      b_lo += kb16;
    }
    a_tap += row_step;
  }
  umma_commit_arrive(&a_empty[ms.a_buf]);
  if (++ms.a_buf == (uint32_t)a.a_bufs) { ms.a_buf = 0; ms.a_phase ^= 1u; }
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Persistent: CTA (x, y) walks bands x, x + gridDim.x, ... of N tile y.  Every ring (A chunk buffers, B stages, TMEM
// accumulator buffers) is tracked by a running counter, so the TMA producer runs ahead into the next band while the
// MMAs of the current one are still in flight, and the epilogue of band i overlaps the MMAs of band i+1.
__global__ void __launch_bounds__(kThreadsWide)
k_conv_sl(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, SlArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *a_smem = smem;
  uint8_t *b_smem = smem + (size_t)a.a_bufs * a.a_buf_bytes;
  uint64_t *a_full = (uint64_t *)(b_smem + (size_t)a.b_stages * a.b_stage_bytes);
  uint64_t *a_empty = a_full + 2;
  uint64_t *b_full = a_empty + 2;
  uint64_t *b_empty = b_full + kMaxBStages;
  uint64_t *t_full = b_empty + kMaxBStages;
  uint64_t *t_empty = t_full + 2;
  uint32_t *tmem_slot = (uint32_t *)(t_empty + 2);
  float *bias_s = (float *)(((uintptr_t)(tmem_slot + 1) + 15) & ~(uintptr_t)15);  // NT floats

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  if (threadIdx.x == 0) {
    SL_TRACE(0);
    if (a.trace) {
      uint32_t smid;
      asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
      a.trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + 7] = smid;
    }
  }
  const int num_bands = a.N * a.bands_h * a.bands_w;
  const int n0 = blockIdx.y * a.NT;
  const int taps = a.kh * a.kw;
  const int acc_cols = a.MTB * a.NT;  // TMEM columns of one accumulator buffer

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    if (!a.c4) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1);
      mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], (blockDim.x - 64) / 32);
    }
    for (int s = 0; s < kMaxBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], (uint32_t)a.cl); }
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
  if (a.cl > 1) cluster_sync_all();  // the peer's barriers must be initialised before anything is multicast into them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();  // everything above overlapped the previous kernel; from here on global memory is touched
  if (threadIdx.x == 0) SL_TRACE(1);
  // every CTA walks the same number of band slots; a slot past the last band only keeps a shared weight ring in lock-step
  const int iters = (num_bands + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t crank = a.cl > 1 ? cluster_ctarank() : 0u;

// band index -> image, origin and number of M-tiles holding at least one real pixel
#define SL_BAND_GEOM(band)                                                                 \
  const int bw_i = (band) % a.bands_w;                                                     \
  const int bq_ = (band) / a.bands_w;                                                      \
  const int bh_i = bq_ % a.bands_h;                                                        \
  const int img = bq_ / a.bands_h;                                                         \
  const int oy0 = bh_i * a.TH, ox0 = bw_i * a.TW;                                          \
  const int rows_valid = min(a.TH, a.Ho - oy0), cols_valid = min(a.TW, a.Wo - ox0);        \
  const int mtb = ((rows_valid - 1) * a.BW + cols_valid + 127) >> 7;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (a.c4) {
      if (elect_one()) {  // every K-block of this N tile stays resident
        const uint32_t wbytes = (uint32_t)a.b_stage_bytes;
        mbar_expect_tx(&b_full[0], wbytes);
        const uint8_t *src = (const uint8_t *)a.wpack + (size_t)blockIdx.y * wbytes;
        for (uint32_t off = 0; off < wbytes; off += 16384u)
          bulk_g2s(b_smem + off, src + off, min(16384u, wbytes - off), &b_full[0]);
      }
      __syncwarp();
    }
    if (!a.c4 && a.b_resident) {
      if (elect_one()) {
        mbar_expect_tx(&b_full[0], (uint32_t)(a.b_stages * a.b_stage_bytes));
        for (int kb = 0; kb < a.b_stages; ++kb)
          tma_load_3d(&mapB, &b_full[0], b_smem + (size_t)kb * a.b_stage_bytes, 0, n0, kb);
      }
      __syncwarp();
    }
    uint32_t ac = 0, kb = 0;  // running A-chunk / K-block counters
    const uint32_t half_rows = (uint32_t)(a.NT / a.cl), half_bytes = half_rows * 128u;
    for (int it = 0; it < iters; ++it) {
      const int band = (int)blockIdx.x + it * (int)gridDim.x;
      const bool valid = band < num_bands;
      if (!valid && (a.cl == 1 || a.c4 || a.b_resident)) break;
      SL_BAND_GEOM(valid ? band : 0)
      (void)mtb;
      const int ix0 = ox0 - a.pad_w, iy0 = oy0 - a.pad;
      for (int c = 0; c < a.chunks; ++c) {
        const uint32_t buf = ac % (uint32_t)a.a_bufs;
        if (valid) mbar_wait(&a_empty[buf], ((ac / (uint32_t)a.a_bufs) & 1u) ^ 1u);
        if (valid && elect_one()) {
          if ((a.dbg & 2) && ac >= (uint32_t)a.a_bufs) {
            mbar_arrive(&a_full[buf]);
          } else {
            mbar_expect_tx(&a_full[buf], (uint32_t)a.a_tx_bytes);
            int cc = c, ix = ix0, iy = iy0;
            if (a.in_ps > 1) {  // pixel-un-shuffle folded into the load: phase (i, j) of the shuffled gradient
              const int ij = c / a.in_cpb;
              cc = c - ij * a.in_cpb;
              iy = iy0 * a.in_ps + ij / a.in_ps;
              ix = ix0 * a.in_ps + ij % a.in_ps;
            }
            tma_load_4d(&mapA, &a_full[buf], a_smem + (size_t)buf * a.a_buf_bytes, cc * a.chunk_elems, ix, iy, img);
          }
        }
        __syncwarp();
        if (valid) ++ac;
        if (!a.c4 && !a.b_resident) {
          const bool masked = a.in_ps > 1;
          const uint32_t tmask = masked ? (uint32_t)a.phase_mask[c / a.in_cpb] : 0u;
          for (int tap = 0; tap < taps; ++tap) {
            if (masked && !((tmask >> tap) & 1u)) continue;
            const uint32_t st = kb++ % (uint32_t)a.b_stages;
            mbar_wait(&b_empty[st], ((((kb - 1u) / (uint32_t)a.b_stages)) & 1u) ^ 1u);  // released by every CTA that shares the ring
            if (elect_one()) {
              mbar_expect_tx(&b_full[st], (uint32_t)a.b_stage_bytes);
              if (a.cl > 1)  // my half of the stage, into both CTAs
                tma_load_3d_mc(&mapB, &b_full[st], b_smem + (size_t)st * a.b_stage_bytes + crank * half_bytes, 0,
                               n0 + (int)(crank * half_rows), c * taps + tap, (uint16_t)3);
              else
                tma_load_3d(&mapB, &b_full[st], b_smem + (size_t)st * a.b_stage_bytes, 0, n0, c * taps + tap);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: D=f32, A=B=tf32, both K-major, N>>3 at bit 17, M>>4 at bit 24
    // (kind::f16 with bf16 operands: a/b format 1 instead of 2)
    const uint32_t fmt = a.in_bf16 ? 1u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(a.NT >> 3) << 17) | ((128u >> 4) << 24);
    const long long clk0 = clock64();
    const long long gt0 = gtime();
    MmaState ms;
    ms.a_buf = 0; ms.a_phase = 0; ms.b_st = 0; ms.b_phase = 0; ms.t_buf = 0; ms.t_phase = 0;
    if (a.c4 || a.b_resident) mbar_wait(&b_full[0], 0);
    bool first = true;
    if (elect_one())
    for (int it = 0; it < iters; ++it) {
      const int band = (int)blockIdx.x + it * (int)gridDim.x;
      if (band >= num_bands) {
        if (a.cl == 1 || a.c4 || a.b_resident) break;
        // idle slot of a sharing pair: consume the weight stages without issuing MMAs so that the peer's ring keeps moving
        for (int q = 0; q < a.kb_valid; ++q) {
          mbar_wait(&b_full[ms.b_st], ms.b_phase);
          umma_commit_arrive_mc(&b_empty[ms.b_st], 3);
          if (++ms.b_st == (uint32_t)a.b_stages) { ms.b_st = 0; ms.b_phase ^= 1u; }
        }
        continue;
      }
      SL_BAND_GEOM(band)
      (void)img;
      mbar_wait(&t_empty[ms.t_buf], ms.t_phase ^ 1u);  // epilogue has drained this accumulator buffer
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc = tmem_base + ms.t_buf * (uint32_t)acc_cols;
      const uint32_t a_base = smem_u32(a_smem), b_base = smem_u32(b_smem);
      if (first) SL_TRACE(2);
      first = false;
      // straight-line issue code per M-tile count (no per-MMA predicates, descriptors stay in uniform registers)
      if (a.c4) {
        if (mtb == 1) mma_band_c4<1>(a, ms, a_full, a_empty, a_base, b_base, tacc, idesc);
        else if (mtb == 2) mma_band_c4<2>(a, ms, a_full, a_empty, a_base, b_base, tacc, idesc);
        else if (mtb == 3) mma_band_c4<3>(a, ms, a_full, a_empty, a_base, b_base, tacc, idesc);
        else mma_band_c4<4>(a, ms, a_full, a_empty, a_base, b_base, tacc, idesc);
      } else if (a.in_bf16) {
        if (mtb == 1) mma_band_generic<1, true>(a, ms, a_full, a_empty, b_full, b_empty, a_base, b_base, tacc, idesc);
        else if (mtb == 2) mma_band_generic<2, true>(a, ms, a_full, a_empty, b_full, b_empty, a_base, b_base, tacc, idesc);
        else if (mtb == 3) mma_band_generic<3, true>(a, ms, a_full, a_empty, b_full, b_empty, a_base, b_base, tacc, idesc);
        else mma_band_generic<4, true>(a, ms, a_full, a_empty, b_full, b_empty, a_base, b_base, tacc, idesc);
      } else {
        if (mtb == 1) mma_band_generic<1, false>(a, ms, a_full, a_empty, b_full, b_empty, a_base, b_base, tacc, idesc);
        else if (mtb == 2) mma_band_generic<2, false>(a, ms, a_full, a_empty, b_full, b_empty, a_base, b_base, tacc, idesc);
        else if (mtb == 3) mma_band_generic<3, false>(a, ms, a_full, a_empty, b_full, b_empty, a_base, b_base, tacc, idesc);
        else mma_band_generic<4, false>(a, ms, a_full, a_empty, b_full, b_empty, a_base, b_base, tacc, idesc);
      }
      umma_commit_arrive(&t_full[ms.t_buf]);  // this band's accumulators are complete
      if (++ms.t_buf == (uint32_t)a.t_bufs) { ms.t_buf = 0; ms.t_phase ^= 1u; }
    }
    __syncwarp();
    if (lane == 0 && a.trace) {  // slot 3: MMA-loop duration in ns (low 32 bits) and in SM cycles (high 32 bits)
      const long long dc = clock64() - clk0, dt = gtime() - gt0;
      a.trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + 3] = (dc << 32) | (dt & 0xffffffffLL);
    }
  } else {
    // ===================== epilogue: warps 2..9; warp w reads TMEM lanes 32*(w%4) .. +31 =====================
    // Two warps share each lane quadrant and alternate over the (M-tile, 16-column group) work items.
    const int lane_grp = warp & 3, half = (warp - 2) >> 2, halves = ((int)blockDim.x - 64) / 128;  // epilogue warps per lane quadrant
    // bias -> shared memory while the first MMAs run (zero when absent or beyond Co)
    for (int j = threadIdx.x - 64; j < a.NT; j += (int)blockDim.x - 64) {
      const int co = n0 + j;
      bias_s[j] = (a.epi.bias && co < a.Co) ? __ldg(a.epi.bias + co) : 0.f;
    }
    const bool lay0 = a.out.sc == 1 && (!a.epi.residual.p || a.epi.residual.sc == 1) &&
                      (!a.epi.preact.p || a.epi.preact.sc == 1) && (!a.epi.mask.p || a.epi.mask.sc == 1);
    const bool lay1 = a.out.sw == 1 && (!a.epi.residual.p || a.epi.residual.sw == 1) &&
                      (!a.epi.preact.p || a.epi.preact.sw == 1);
    int fmode = 3;  // see epilogue_items
    int hmode = 3;  // bf16 outputs: epilogue_items_h mode (0, 2) or the generic scalar path (3)
    if (a.out_bf16 && (a.Co & 15) == 0 && lay0 && a.v8h && !a.epi.mask.p) {
      if (a.ps == 1) hmode = 0;
      else if (a.ps == 2 && (a.NT & 31) == 0 && (a.Co & 31) == 0 && !a.epi.bits_out && !a.epi.bits_in) hmode = 2;
    }
    if (!a.out_bf16 && (a.Co & 15) == 0) {
      if (a.ps == 1 && lay0) fmode = 0;
      else if (a.epi.mask.p) fmode = 3;  // masks only exist on the vector path of mode 0
      else if (a.ps == 4 && lay1 && (a.out.sh & 3) == 0) fmode = 1;
      else if (a.ps == 2 && lay0 && ((a.Co >> 2) & 3) == 0) fmode = 2;
    }
    const bool extra = a.epi.residual.p != nullptr || a.epi.preact.p != nullptr || a.epi.mask.p != nullptr;
    asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x - 64) : "memory");  // bias_s visible to all epilogue warps
    const uint32_t bsa = smem_u32(bias_s);
    const int m = lane_grp * 32 + lane;
    float lsum = 0.f;  // this thread's share of the fused loss sum
    const int lmode = a.epi.loss_kind ? ((a.ps == 4 && a.epi.dz_unshuf) ? 1 : 3) : 0;
    uint32_t wi = 0;
    for (int band = blockIdx.x; band < num_bands; band += gridDim.x, ++wi) {
      SL_BAND_GEOM(band)
      const uint32_t tb = wi % (uint32_t)a.t_bufs;
      mbar_wait(&t_full[tb], (wi / (uint32_t)a.t_bufs) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (wi == 0 && warp == 2 && lane == 0) SL_TRACE(4);
      const uint32_t trow = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + tb * (uint32_t)acc_cols;
#define SL_EPI(MODE, EXTRA) epilogue_items<MODE, EXTRA>(a, trow, bsa, half, mtb, m, img, n0, oy0, ox0, rows_valid, cols_valid, halves)
#define SL_EPIH(MODE) epilogue_items_h<MODE>(a, trow, bsa, half, mtb, m, img, n0, oy0, ox0, rows_valid, cols_valid, halves)
      if (a.dbg & 1) {}
      else if (lmode == 1) epilogue_items_loss<1>(a, trow, bsa, half, mtb, m, img, n0, oy0, ox0, rows_valid, cols_valid, lsum, halves);
      else if (lmode == 3) epilogue_items_loss<3>(a, trow, bsa, half, mtb, m, img, n0, oy0, ox0, rows_valid, cols_valid, lsum, halves);
      else if (a.out_bf16 && hmode == 0) SL_EPIH(0);
      else if (a.out_bf16 && hmode == 2) SL_EPIH(2);
      else if (fmode == 0) { if (extra) SL_EPI(0, true); else SL_EPI(0, false); }
      else if (fmode == 1) { if (extra) SL_EPI(1, true); else SL_EPI(1, false); }
      else if (fmode == 2) { if (extra) SL_EPI(2, true); else SL_EPI(2, false); }
      else SL_EPI(3, true);
#undef SL_EPI
#undef SL_EPIH
      // all TMEM reads of this warp are complete (tcgen05.wait::ld inside tmem_ld16): hand the buffer back
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[tb]);
      if (wi == 0 && warp == 2 && lane == 0) SL_TRACE(5);
    }
    if (lmode) {  // one partial per epilogue warp, summed over the lanes in a fixed (butterfly) order
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
      if (lane == 0) a.epi.loss_part[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * ((blockDim.x - 64) / 32) + (warp - 2)] = lsum;
    }
  }
#undef SL_BAND_GEOM

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (a.cl > 1) cluster_sync_all();  // no CTA of the pair may exit while the other can still signal its barriers
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols)
                 : "memory");
  }
  if (threadIdx.x == 0) SL_TRACE(6);
}

// loss = inv_n * sum of the per-warp partials, in a fixed order (one block)
__global__ void __launch_bounds__(256) k_sl_loss_finish(const float *__restrict__ partial, int n, float inv_n, float *loss) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = red[0] * inv_n;
}

// Filter element for GEMM row n (output channel of this launch), K index k (input channel of this launch), tap (r,s):
//   flip == 0 (fprop):  w[n][k][r][s]                      (Conv2d OIHW, Ci = Kk)
//   flip == 1 (dgrad):  w[k][n][kh-1-r][kw-1-s]            (the launch's output channel is the filter's input channel)
__device__ __forceinline__ float wval(const float *__restrict__ w, int n, int k, int r, int s, int Nn, int Kk, int kh,
                                      int kw, int flip) {
  if (!flip) return __ldg(w + (((long long)n * Kk + k) * kh + r) * kw + s);
  return __ldg(w + (((long long)k * Nn + n) * kh + (kh - 1 - r)) * kw + (kw - 1 - s));
}

// generic B operand: out[chunk][tap][Npad][32] (tf32 RN), zero for n >= Nn or k >= Kk
// perm_C > 0 (input un-shuffle fold): launch K index k = ij * perm_C + c stands for the filter's channel c * perm_rr + ij
__device__ __forceinline__ int perm_k(int k, int perm_C, int perm_rr) {
  if (perm_C <= 0) return k;
  const int ij = k / perm_C;
  return (k - ij * perm_C) * perm_rr + ij;
}

// Filter addressing of the phase launches of strided / transposed convolutions (ConvOpt::wmode 1 and 2, srb_common.cuh).
struct WMap {
  int wmode, st, pad0, kh0, kw0, dmin_r, dmin_s, ra, rb, tmax_a, tmax_b, C;  // C: channels per input phase (wmode 1)
};
// value of launch filter element (n, k, tap (tr, ts)); Nn / Kk: launch output / reduction channel counts
__device__ __forceinline__ float wval_phase(const float *__restrict__ w, const WMap &m, int n, int k, int tr, int ts, int Nn, int Kk) {
  if (m.wmode == 1) {
    const int ph = k / m.C, c = k - ph * m.C, a = ph / m.st, b = ph - a * m.st;
    const int r = m.st * (tr + m.dmin_r) + a + m.pad0, s = m.st * (ts + m.dmin_s) + b + m.pad0;
    if (r < 0 || r >= m.kh0 || s < 0 || s >= m.kw0) return 0.f;
    return __ldg(w + (((long long)n * m.C + c) * m.kh0 + r) * m.kw0 + s);
  }
  const int r = m.ra + m.st * (m.tmax_a - tr), s = m.rb + m.st * (m.tmax_b - ts);
  if (r < 0 || r >= m.kh0 || s < 0 || s >= m.kw0) return 0.f;
  return __ldg(w + (((long long)k * Nn + n) * m.kh0 + r) * m.kw0 + s);
}

__global__ void k_pack_w_sl(const float *__restrict__ w, float *__restrict__ out, int Nn, int Kk, int kh, int kw, int Npad,
                            int chunks, int flip, int perm_C, int perm_rr, WMap wm) {
  pdl_trigger();
  pdl_wait();
  const int taps = kh * kw;
  const long long total = (long long)chunks * taps * Npad * 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(i & 31);
    long long q = i >> 5;
    const int n = (int)(q % Npad); q /= Npad;
    const int tap = (int)(q % taps);
    const int c = (int)(q / taps);
    const int k = c * 32 + kk;
    float v = 0.f;
    if (n < Nn && k < Kk) {
      const int r = tap / kw, s = tap - r * kw;
      v = wm.wmode ? round_tf32(wval_phase(w, wm, n, k, r, s, Nn, Kk))
                   : round_tf32(wval(w, n, perm_k(k, perm_C, perm_rr), r, s, Nn, Kk, kh, kw, flip));
    }
    out[i] = v;
  }
}

// generic B operand for bf16 operands: out[chunk][tap][Npad][64] (bf16 RN), zero for n >= Nn or k >= Kk
__global__ void k_pack_w_sl_h(const float *__restrict__ w, unsigned short *__restrict__ out, int Nn, int Kk, int kh, int kw,
                              int Npad, int chunks, int flip, int perm_C, int perm_rr, WMap wm) {
  pdl_trigger();
  pdl_wait();
  const int taps = kh * kw;
  const long long total = (long long)chunks * taps * Npad * 64;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(i & 63);
    long long q = i >> 6;
    const int n = (int)(q % Npad); q /= Npad;
    const int tap = (int)(q % taps);
    const int c = (int)(q / taps);
    const int k = c * 64 + kk;
    float v = 0.f;
    if (n < Nn && k < Kk) {
      const int r = tap / kw, s = tap - r * kw;
      v = wm.wmode ? wval_phase(w, wm, n, k, r, s, Nn, Kk) : wval(w, n, perm_k(k, perm_C, perm_rr), r, s, Nn, Kk, kh, kw, flip);
    }
    out[i] = (unsigned short)(pack_bf16x2(v, 0.f) & 0xffffu);
  }
}

// c4 B operand: out[ntile][kb = r*spairs+sp][NT/8][2][8][4]; element (n, k): pixel offset sl = k/4 -> s = 2*sp+sl, channel k%4
__global__ void k_pack_w_c4(const float *__restrict__ w, float *__restrict__ out, int Nn, int Kk, int kh, int kw, int NT,
                            int ntiles, int spairs, int flip) {
  pdl_trigger();
  pdl_wait();
  const int kblocks = kh * spairs;
  const long long total = (long long)ntiles * kblocks * NT * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k4 = (int)(i & 3);
    const int n8 = (int)((i >> 2) & 7);
    const int kh2 = (int)((i >> 5) & 1);
    long long q = i >> 6;
    const int ng = (int)(q % (NT / 8)); q /= (NT / 8);
    const int kb = (int)(q % kblocks);
    const int t = (int)(q / kblocks);
    const int n = t * NT + ng * 8 + n8;
    const int r = kb / spairs, sp = kb - r * spairs;
    const int s = 2 * sp + kh2, ci = k4;
    float v = 0.f;
    if (n < Nn && ci < Kk && s < kw) v = round_tf32(wval(w, n, ci, r, s, Nn, Kk, kh, kw, flip));
    out[i] = v;
  }
}

// x (N, C<=4, H, W; any strides) -> NHWC4 image (16 B per pixel, channel >= C zero), tf32-rounded
__global__ void k_pack_nhwc4(T4 x, float4 *__restrict__ xp, int N, int C, int H, int W) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)N * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % W);
    long long q = i / W;
    const int yy = (int)(q % H);
    const int n = (int)(q / H);
    const float *p = x.p + n * x.sn + (long long)yy * x.sh + (long long)xx * x.sw;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < C) v[c] = round_tf32(__ldg(p + c * x.sc));
    xp[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}


// ================================================================================================================================
// Row-stacked variant (k_conv_rs) for narrow layers (kh * NT <= 256, all weights resident).
//
// The slot-linear kernel above issues one MMA per filter tap with N = Cout: at Cout = 32..64 an SS-mode MMA is bound by
// re-reading its 4 KB A tile from shared memory (32 + N/4 cycles against N/2 of math, tools/bench_umma.cu).  Here an M tile is
// ONE INPUT ROW -- 128 pixel slots: G images side by side, BW slots each -- and one MMA multiplies it with the taps of ALL kh
// filter rows at once: the B operand of horizontal tap s is [kh][NT] filter rows stacked along N (stored r = kh-1 .. 0), and
// the accumulator of output row o is the TMEM column block o*NT, so input row j accumulates into the CONTIGUOUS column
// window of output rows j-(kh-1) .. j.  One A read now feeds kh taps: 53 / 60 / 69 cycles per MMA at N = 96 / 144 / 192
// instead of kh x (40 / 44 / 48) (tools/bench_umma_rows.cu).  Horizontal taps remain shifted views of the same row.
//
// Rows stream top to bottom: a CTA owns a contiguous range of the global (image group, column strip, output row) sequence,
// loads every input row once per 32-channel chunk into a shared-memory ring, and the TMEM column blocks form a ring as well:
// output row o is complete when input row o+kh-1 has been multiplied, is drained by the epilogue warps and its block is
// reused R rows later -- no bands, no vertical halo except at the start of a CTA's range.
constexpr int kRsMaxStages = 8;
constexpr int kRsMaxBlocks = 16;
constexpr int kRsPrefetchRows = 6;
constexpr int kRsEpiPerQuad = kRsEpiPerQuadFwd;        // epilogue warps per TMEM lane quadrant
// Warps 0-1: TMA producers, warps 2-3: MMA issuers (one of each per row STREAM), then the epilogue warps.  With two streams a CTA
// walks two halves of its row range concurrently -- each stream has its own operand ring, half of the TMEM ring and half of the
// epilogue warps; the resident weights are shared -- so that one issuing thread's per-row bookkeeping (barrier waits, commits,
// descriptor updates: ~0.25 us per row that the tensor pipe otherwise idles through) overlaps the other stream's MMAs.
constexpr int kRsThreads = 128 + 128 * kRsEpiPerQuad;

__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap *map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

struct RsArgs {
  SlArgs e;               // what the epilogues read: N, Ho, Wo, Co, NT, BW (= slots per image), ps, out, epi, v8, rs = 1
  int chunks, kv_last;    // 32-channel chunks of the input; K steps (of 8 channels) that exist in the last chunk
  int G, TW, CT;          // images per M tile, output columns per strip, column strips per image
  int R;                  // TMEM ring: accumulator blocks of NT columns (per stream)
  int streams;            // 1 or 2 concurrent row streams per CTA (S and R are per stream)
  int S, stage_bytes, stage_tx;  // A ring: one stage = one input row, all chunks (chunk_bytes each; stage_tx = TMA bytes per chunk)
  int chunk_bytes;
  int bcs_bytes;          // B: bytes of one (chunk, s) operand = kh * NT * 128
  long long total_rows;   // image groups x column strips x Ho
};

// debug trace (srb_debug_set_trace): CTA 0 writes globaltimer stamps, 4 roles x 128 events
#define RS_TRACE(role, idx)                                                                                       \
  do {                                                                                                            \
    if (a.e.trace && blockIdx.x == 0 && (idx) < 128) a.e.trace[(role) * 128 + (idx)] = gtime();                   \
  } while (0)

struct RsSeg {  // a run of output rows inside one (image group, column strip)
  int img0, ox0, oy0, cnt, imgs_valid, cols_valid;
};
__device__ __forceinline__ RsSeg rs_segment(const RsArgs &a, long long row, long long r1) {
  RsSeg sg;
  const long long strip = row / a.e.Ho;
  sg.oy0 = (int)(row - strip * a.e.Ho);
  const long long left = r1 - row;
  sg.cnt = (int)(left < (long long)(a.e.Ho - sg.oy0) ? left : (long long)(a.e.Ho - sg.oy0));
  const int ig = (int)(strip / a.CT), ct = (int)(strip - (long long)ig * a.CT);
  sg.img0 = ig * a.G;
  sg.ox0 = ct * a.TW;
  sg.imgs_valid = min(a.G, a.e.N - sg.img0);
  sg.cols_valid = min(a.TW, a.e.Wo - sg.ox0);
  return sg;
}

struct RsPart { uint32_t tcol, boff, idesc; };  // one MMA: D column, B row offset (16-byte units), instruction descriptor
// All MMAs of one input row in the steady state (see k_conv_rs): for every chunk, kw horizontal taps x 4 K-steps accumulate into
// the window of kh output-row blocks starting at TMEM column tc_lo; the very first MMA of the row is split -- the kh-1 older
// blocks accumulate, the new block (tc_hi, met by filter row 0 = B row block kh-1) is overwritten.  Straight-line issue code.
template <int KW>
__device__ __forceinline__ void rs_issue_row_steady(int chunks, int kv_last, uint32_t a_st, uint32_t chunk16, uint32_t b_lo0, uint32_t bcs16,
                                                    uint32_t tc_lo, uint32_t tc_hi, uint32_t id_full, uint32_t id_win1, uint32_t id_one,
                                                    uint32_t boff_new) {
  const uint64_t d_hi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;  // SWIZZLE_128B, SBO 1024 B
  umma_tf32_ss(tc_lo, d_hi | (uint64_t)a_st, d_hi | (uint64_t)b_lo0, id_win1, 1u);
  umma_tf32_ss(tc_hi, d_hi | (uint64_t)a_st, d_hi | (uint64_t)(b_lo0 + boff_new), id_one, 0u);
#pragma unroll 1
  for (int c = 0; c < chunks; ++c) {
    const uint32_t ac = a_st + (uint32_t)c * chunk16, bc = b_lo0 + (uint32_t)(c * KW) * bcs16;
    const int kv = c == chunks - 1 ? kv_last : 4;  // K steps (8 channels each) that exist in this chunk
#pragma unroll
    for (int s = 0; s < KW; ++s) {
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        if (k4 >= kv) continue;
        if (s == 0 && k4 == 0) {
          if (c) umma_tf32_ss(tc_lo, d_hi | (uint64_t)ac, d_hi | (uint64_t)bc, id_full, 1u);
        } else {
          umma_tf32_ss(tc_lo, d_hi | (uint64_t)(ac + (uint32_t)(8 * s + 2 * k4)), d_hi | (uint64_t)(bc + (uint32_t)s * bcs16 + (uint32_t)(2 * k4)), id_full,
                       1u);
        }
      }
    }
  }
}

// General version for one (input row, chunk): band edges (partial windows), TMEM ring wrap-around (two column ranges), partial
// last chunk, any filter width.  `first` (chunk 0, tap 0, K-step 0 of a fresh row): fa0 / fa1 accumulate, fo overwrites.
__device__ __forceinline__ void rs_issue_chunk(bool chunk0, bool fresh, int kv, int kw, uint32_t a_row, uint32_t b_cs, uint32_t bcs16,
                                               const RsPart &acc0, const RsPart &acc1, bool two, const RsPart &fa0, const RsPart &fa1,
                                               uint32_t n1f, uint32_t nf, const RsPart &fo) {
  const uint64_t d_hi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
  const bool first = chunk0 && fresh;
  if (first) {
    const uint64_t da = d_hi | (uint64_t)a_row;
    if (n1f) umma_tf32_ss(fa0.tcol, da, d_hi | (uint64_t)(b_cs + fa0.boff), fa0.idesc, 1u);
    if (nf > n1f) umma_tf32_ss(fa1.tcol, da, d_hi | (uint64_t)(b_cs + fa1.boff), fa1.idesc, 1u);
    umma_tf32_ss(fo.tcol, da, d_hi | (uint64_t)(b_cs + fo.boff), fo.idesc, 0u);
  }
  for (int s = 0; s < kw; ++s) {
    for (int k4 = 0; k4 < kv; ++k4) {
      if (s == 0 && k4 == 0 && first) continue;
      const uint64_t da = d_hi | (uint64_t)(a_row + (uint32_t)(8 * s + 2 * k4));
      const uint32_t bk = b_cs + (uint32_t)s * bcs16 + (uint32_t)(2 * k4);
      umma_tf32_ss(acc0.tcol, da, d_hi | (uint64_t)(bk + acc0.boff), acc0.idesc, 1u);
      if (two) umma_tf32_ss(acc1.tcol, da, d_hi | (uint64_t)(bk + acc1.boff), acc1.idesc, 1u);
    }
  }
}

__global__ void __launch_bounds__(kRsThreads)
k_conv_rs(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, RsArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *a_smem = smem;
  uint8_t *b_smem = smem + (size_t)a.streams * a.S * a.stage_bytes + 1024;  // 1 KB gap: the shifted views of the last stage over-read kw-1 slots
  const int kh = a.e.kh, kw = a.e.kw, NT = a.e.NT;
  uint64_t *a_full0 = (uint64_t *)(b_smem + (size_t)a.chunks * kw * a.bcs_bytes);  // [stream][kRsMaxStages]
  uint64_t *a_empty0 = a_full0 + 2 * kRsMaxStages;
  uint64_t *t_full0 = a_empty0 + 2 * kRsMaxStages;                                  // [stream][kRsMaxBlocks]
  uint64_t *t_empty0 = t_full0 + 2 * kRsMaxBlocks;
  uint64_t *b_full = t_empty0 + 2 * kRsMaxBlocks;
  uint32_t *tmem_slot = (uint32_t *)(b_full + 1);
  float *bias_s = (float *)(((uintptr_t)(tmem_slot + 1) + 15) & ~(uintptr_t)15);  // NT floats

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    for (int s = 0; s < 2 * kRsMaxStages; ++s) { mbar_init(&a_full0[s], 1); mbar_init(&a_empty0[s], 1); }
    for (int s = 0; s < 2 * kRsMaxBlocks; ++s) { mbar_init(&t_full0[s], 1); mbar_init(&t_empty0[s], 4); }
    mbar_init(b_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base0 = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();  // everything above overlapped the previous kernel; from here on global memory is touched

  // this CTA's share of the global output-row sequence, and this warp's stream's share of that
  const long long c0 = a.total_rows * (long long)blockIdx.x / (long long)gridDim.x;
  const long long c1 = a.total_rows * (long long)(blockIdx.x + 1) / (long long)gridDim.x;
  // role -> stream: producers 0/1, issuers 2/3; epilogue warp e = warp - 4: with two streams e / 4 alternates between them
  const int q = a.streams == 1 ? 0 : (warp < 4 ? (warp & 1) : (((warp - 4) >> 2) & 1));
  const long long cm = a.streams == 1 ? c1 : c0 + (c1 - c0 + 1) / 2;
  const long long r0 = q == 0 ? c0 : cm, r1 = q == 0 ? cm : c1;
  uint64_t *a_full = a_full0 + q * kRsMaxStages, *a_empty = a_empty0 + q * kRsMaxStages;
  uint64_t *t_full = t_full0 + q * kRsMaxBlocks, *t_empty = t_empty0 + q * kRsMaxBlocks;
  a_smem += (size_t)q * a.S * a.stage_bytes;
  const uint32_t tmem_base = tmem_base0 + (uint32_t)(q * a.R * NT);  // this stream's half of the accumulator ring
  const bool idle_role = a.streams == 1 && (warp == 1 || warp == 3);

  if (idle_role) {
  } else if (warp < 2) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const int ncs = a.chunks * kw;
      if (warp == 0) {  // the resident weights, shared by both streams
        mbar_expect_tx(b_full, (uint32_t)(ncs * a.bcs_bytes));
        for (int cs = 0; cs < ncs; ++cs) tma_load_3d(&mapB, b_full, b_smem + (size_t)cs * a.bcs_bytes, 0, 0, cs);
      }
      uint32_t st = 0, ph = 0, nloaded = 0;
      int tr_i = 0;
      for (long long row = r0; row < r1;) {
        const RsSeg sg = rs_segment(a, row, r1);
        const int nin = sg.cnt + kh - 1;
        for (int j = 0; j < nin; ++j) {
          const int iy = sg.oy0 - a.e.pad + j;
          mbar_wait(&a_empty[st], ph ^ 1u);
          if (q == 0) RS_TRACE(0, tr_i);
          ++tr_i;
          if ((a.e.dbg & 2) && nloaded >= (uint32_t)a.S) {
            mbar_arrive(&a_full[st]);  // debug: operands are loaded only once per stage
          } else {
            mbar_expect_tx(&a_full[st], (uint32_t)(a.stage_tx * a.chunks));
            for (int c = 0; c < a.chunks; ++c)
              tma_load_4d(&mapA, &a_full[st], a_smem + (size_t)st * a.stage_bytes + (size_t)c * a.chunk_bytes, c * 32, sg.ox0 - a.e.pad_w,
                          sg.img0, iy);
          }
          ++nloaded;
          if (++st == (uint32_t)a.S) { st = 0; ph ^= 1u; }
        }
        row += sg.cnt;
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    // ===================== MMA issuer: one elected lane runs the whole loop =====================
    // (scalar state only and straight-line issue code: ptxas then keeps descriptors, TMEM addresses and instruction
    // descriptors in UNIFORM registers, which is what UTCHMMA reads; per-lane guards or indexed structs cost an R2UR per operand)
    if (elect_one()) {
    const uint32_t idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 4) << 24);  // D=f32, A=B=tf32 K-major, M=128; N added per MMA
    const uint32_t tmem_u = tmem_base;
    const uint32_t a_base16 = smem_u32(a_smem) >> 4, stage16 = (uint32_t)a.stage_bytes >> 4;
    const uint32_t b_base16 = smem_u32(b_smem) >> 4, bcs16 = (uint32_t)a.bcs_bytes >> 4;
    const uint32_t R = (uint32_t)a.R, nt8 = (uint32_t)NT * 8u, chunk16 = (uint32_t)a.chunk_bytes >> 4;
    const uint32_t b_lo0 = (b_base16 & 0x3FFF) | (1u << 16);
    const uint32_t id_full = idesc0 | ((((uint32_t)(kh * NT)) >> 3) << 17), id_win1 = idesc0 | ((((uint32_t)((kh - 1) * NT)) >> 3) << 17),
                   id_one = idesc0 | (((uint32_t)NT >> 3) << 17);
    mbar_wait(b_full, 0);
    uint32_t st = 0, ph = 0;
    uint32_t blk_hi = 0, par_hi = 0;  // ring block / use parity of the next output row to be started
    uint32_t blk_lo = 0;              // ring block of the oldest incomplete output row
    int tr_m = 0;
    for (long long row = r0; row < r1;) {
      const RsSeg sg = rs_segment(a, row, r1);
      const int nin = sg.cnt + kh - 1;
      for (int j = 0; j < nin; ++j) {
        if (q == 0) RS_TRACE(1, tr_m);
        ++tr_m;
        const bool fresh = j < sg.cnt;  // output row j is touched for the first time: its block is overwritten
        if (fresh) {
          mbar_wait(&t_empty[blk_hi], par_hi ^ 1u);  // drained by the epilogue
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        mbar_wait(&a_full[st], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_st = ((a_base16 + st * stage16) & 0x3FFF) | (1u << 16);
        // steady state: a full window of kh blocks that does not wrap around the TMEM ring, no partial last chunk
        const bool steady = fresh && j >= kh - 1 && blk_lo + (uint32_t)kh <= R && kh > 1;
        if (a.e.dbg & 4) {
        } else if (steady && kw == 3) {
          rs_issue_row_steady<3>(a.chunks, a.kv_last, a_st, chunk16, b_lo0, bcs16, tmem_u + blk_lo * (uint32_t)NT, tmem_u + blk_hi * (uint32_t)NT,
                                 id_full, id_win1, id_one, (uint32_t)(kh - 1) * nt8);
        } else if (steady && kw == 5) {
          rs_issue_row_steady<5>(a.chunks, a.kv_last, a_st, chunk16, b_lo0, bcs16, tmem_u + blk_lo * (uint32_t)NT, tmem_u + blk_hi * (uint32_t)NT,
                                 id_full, id_win1, id_one, (uint32_t)(kh - 1) * nt8);
        } else {
          const int lo = j - (kh - 1) > 0 ? j - (kh - 1) : 0;
          const uint32_t nb = (uint32_t)((fresh ? j : sg.cnt - 1) - lo + 1);
          const uint32_t rb_lo = (uint32_t)((kh - 1) - (j - lo));  // B row block (filter row kh-1-rb) that meets output row lo
          // accumulate window: blocks blk_lo .. blk_lo+nb-1, split where the ring wraps
          RsPart acc0, acc1, fa0, fa1, fo;
          const uint32_t n1 = nb < R - blk_lo ? nb : R - blk_lo;
          acc0.tcol = tmem_u + blk_lo * (uint32_t)NT; acc0.boff = rb_lo * nt8; acc0.idesc = idesc0 | (((n1 * (uint32_t)NT) >> 3) << 17);
          acc1.tcol = tmem_u; acc1.boff = (rb_lo + n1) * nt8; acc1.idesc = idesc0 | ((((nb - n1) * (uint32_t)NT) >> 3) << 17);
          const bool two = n1 < nb;
          // first MMA of a fresh row: the window without the new block accumulates, the new block is overwritten
          const uint32_t nf = fresh ? nb - 1u : nb;
          const uint32_t n1f = nf < R - blk_lo ? nf : R - blk_lo;
          fa0 = acc0; fa0.idesc = idesc0 | (((n1f * (uint32_t)NT) >> 3) << 17);
          fa1 = acc1; fa1.boff = (rb_lo + n1f) * nt8; fa1.idesc = idesc0 | ((((nf - n1f) * (uint32_t)NT) >> 3) << 17);
          fo.tcol = tmem_u + blk_hi * (uint32_t)NT; fo.boff = (uint32_t)(kh - 1) * nt8; fo.idesc = id_one;
          for (int c = 0; c < a.chunks; ++c) {
            const int kv = (c == a.chunks - 1) ? a.kv_last : 4;
            rs_issue_chunk(c == 0, fresh, kv, kw, a_st + (uint32_t)c * chunk16, b_lo0 + (uint32_t)(c * kw) * bcs16, bcs16, acc0, acc1, two, fa0,
                           fa1, n1f, nf, fo);
          }
        }
        if (a.e.dbg & 32) mbar_arrive(&a_empty[st]);  // debug (only meaningful with dbg & 4): plain arrive instead of tcgen05.commit
        else umma_commit_arrive(&a_empty[st]);
        if (++st == (uint32_t)a.S) { st = 0; ph ^= 1u; }
        if (j >= kh - 1) {  // output row j-(kh-1) is complete
          if (a.e.dbg & 32) mbar_arrive(&t_full[blk_lo]);
          else umma_commit_arrive(&t_full[blk_lo]);
          if (++blk_lo == R) blk_lo = 0;
        }
        if (fresh && ++blk_hi == R) { blk_hi = 0; par_hi ^= 1u; }
      }
      row += sg.cnt;
    }
    }
    __syncwarp();
  } else {
    // ===================== epilogue: warps 2..; warp w reads TMEM lanes 32*(w%4) .. +31; the kRsEpiPerQuad warps of a lane
    // quadrant take output rows round-robin (several rows in flight hide the latency of the epilogue's global loads) =====================
    const SlArgs &e = a.e;
    const int lane_grp = warp & 3;
    // rows of a stream go round-robin over its `halves` epilogue warps per lane quadrant
    const int halves = kRsEpiPerQuad / a.streams, half = a.streams == 1 ? (warp - 4) >> 2 : (warp - 4) >> 3;
    for (int j = threadIdx.x - 128; j < NT; j += kRsThreads - 128) bias_s[j] = (e.epi.bias && j < e.Co) ? __ldg(e.epi.bias + j) : 0.f;
    const bool lay0 = e.out.sc == 1 && (!e.epi.residual.p || e.epi.residual.sc == 1) &&
                      (!e.epi.preact.p || e.epi.preact.sc == 1) && (!e.epi.mask.p || e.epi.mask.sc == 1);
    const bool lay1 = e.out.sw == 1 && (!e.epi.residual.p || e.epi.residual.sw == 1) && (!e.epi.preact.p || e.epi.preact.sw == 1);
    int fmode = 3;
    if ((e.Co & 15) == 0) {
      if (e.ps == 1 && lay0) fmode = 0;
      else if (e.epi.mask.p) fmode = 3;
      else if (e.ps == 4 && lay1 && (e.out.sh & 3) == 0) fmode = 1;
      else if (e.ps == 2 && lay0 && ((e.Co >> 2) & 3) == 0) fmode = 2;
    }
    const bool extra = e.epi.residual.p != nullptr || e.epi.preact.p != nullptr || e.epi.mask.p != nullptr;
    asm volatile("bar.sync 1, %0;" ::"r"(kRsThreads - 128) : "memory");
    const uint32_t bsa = smem_u32(bias_s);
    const int m = lane_grp * 32 + lane;
    float lsum = 0.f;
    const int lmode = e.epi.loss_kind ? ((e.ps == 4 && e.epi.dz_unshuf) ? 1 : 3) : 0;
    uint32_t ctr = 0;
    for (long long row = r0; row < r1;) {
      const RsSeg sg = rs_segment(a, row, r1);
      for (int o = 0; o < sg.cnt; ++o, ++ctr) {
        if ((int)(ctr % (uint32_t)halves) != half) continue;
        const uint32_t blk = ctr % (uint32_t)a.R;
        mbar_wait(&t_full[blk], (ctr / (uint32_t)a.R) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane_grp == 2 && lane == 0 && half < 2 && q == 0) RS_TRACE(2 + half, (int)(ctr / (uint32_t)halves));
        const uint32_t trow = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + blk * (uint32_t)NT;
        const int oy = sg.oy0 + o;
#define RS_EPI(MODE, EXTRA) epilogue_items<MODE, EXTRA>(e, trow, bsa, 0, 1, m, sg.img0, 0, oy, sg.ox0, sg.imgs_valid, sg.cols_valid, 1)
        if (e.dbg & 1) {}
        else if (lmode == 1) epilogue_items_loss<1>(e, trow, bsa, 0, 1, m, sg.img0, 0, oy, sg.ox0, sg.imgs_valid, sg.cols_valid, lsum, 1);
        else if (lmode == 3) epilogue_items_loss<3>(e, trow, bsa, 0, 1, m, sg.img0, 0, oy, sg.ox0, sg.imgs_valid, sg.cols_valid, lsum, 1);
        else if (fmode == 0) { if (extra) RS_EPI(0, true); else RS_EPI(0, false); }
        else if (fmode == 1) { if (extra) RS_EPI(1, true); else RS_EPI(1, false); }
        else if (fmode == 2) { if (extra) RS_EPI(2, true); else RS_EPI(2, false); }
        else RS_EPI(3, true);
#undef RS_EPI
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[blk]);
      }
      row += sg.cnt;
    }
    if (lmode) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
      if (lane == 0) e.epi.loss_part[(size_t)blockIdx.x * (4 * kRsEpiPerQuad) + (warp - 4)] = lsum;  // one slot per epilogue warp
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base0), "r"(512u) : "memory");
  }
}

// B operand of k_conv_rs: out[chunk][s][rr][Npad][32] (tf32 RN), rr = kh-1-r (filter rows stacked along N, last row first)
__global__ void k_pack_w_rs(const float *__restrict__ w, float *__restrict__ out, int Nn, int Kk, int kh, int kw, int Npad, int chunks,
                            int flip) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)chunks * kw * kh * Npad * 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(i & 31);
    long long q = i >> 5;
    const int n = (int)(q % Npad); q /= Npad;
    const int rr = (int)(q % kh); q /= kh;
    const int s = (int)(q % kw);
    const int c = (int)(q / kw);
    const int k = c * 32 + kk;
    float v = 0.f;
    if (n < Nn && k < Kk) v = round_tf32(wval(w, n, k, kh - 1 - rr, s, Nn, Kk, kh, kw, flip));
    out[i] = v;
  }
}


// ================================================================================================================================
// Opt-in packed-weight cache (srb_weight_cache_enable / srb_weight_cache_repack).
//
// Every fprop / dgrad launch needs its filter in an operand layout that depends on the kernel flavour; by default a small pack
// kernel in front of the conv writes it into the call's workspace (3-8 us per call plus a dependent-launch gap: 18 % of the
// EDSR-256 bf16 step, 15 % of EDSR-64, 9 % of the SRGAN iteration).  With the cache enabled the packed copies persist in buffers
// the LIBRARY allocates (the one exception to "torch owns all device memory"), conv calls skip their pack kernel, and the host
// promises to call srb_weight_cache_repack(stream) after every weight update (optimizer step, load_state_dict): ONE launch
// re-packs every cached filter from the current weights.  CUDA-graph friendly: entries are created during the eager warm-up
// (cudaMalloc + table upload are not capturable; a miss during capture falls back to the per-call pack), the repack launch is
// captured behind the optimizer.
struct PackJob {
  const float *w;
  void *out;
  int kind;  // 0: k_pack_w_sl (tf32)  1: k_pack_w_sl_h (bf16)  2: k_pack_w_rs  3: k_pack_w_c4
  int Nn, Kk, kh, kw, Npad, chunks, flip, perm_C, perm_rr;
  int NT, ntiles, spairs;
  WMap wm;
  long long total;  // elements of `out`
};

__device__ __forceinline__ void pack_job_element(const PackJob &j, long long i) {
  if (j.kind == 2) {
    const int kk = (int)(i & 31);
    long long q = i >> 5;
    const int n = (int)(q % j.Npad); q /= j.Npad;
    const int rr = (int)(q % j.kh); q /= j.kh;
    const int s = (int)(q % j.kw);
    const int c = (int)(q / j.kw);
    const int k = c * 32 + kk;
    float v = 0.f;
    if (n < j.Nn && k < j.Kk) v = round_tf32(wval(j.w, n, k, j.kh - 1 - rr, s, j.Nn, j.Kk, j.kh, j.kw, j.flip));
    ((float *)j.out)[i] = v;
  } else if (j.kind == 3) {
    const int kblocks = j.kh * j.spairs;
    const int k4 = (int)(i & 3), n8 = (int)((i >> 2) & 7), kh2 = (int)((i >> 5) & 1);
    long long q = i >> 6;
    const int ng = (int)(q % (j.NT / 8)); q /= (j.NT / 8);
    const int kb = (int)(q % kblocks);
    const int t = (int)(q / kblocks);
    const int n = t * j.NT + ng * 8 + n8;
    const int r = kb / j.spairs, sp = kb - r * j.spairs;
    const int s = 2 * sp + kh2;
    float v = 0.f;
    if (n < j.Nn && k4 < j.Kk && s < j.kw) v = round_tf32(wval(j.w, n, k4, r, s, j.Nn, j.Kk, j.kh, j.kw, j.flip));
    ((float *)j.out)[i] = v;
  } else {
    const int ce = j.kind == 1 ? 64 : 32, sh = j.kind == 1 ? 6 : 5;
    const int taps = j.kh * j.kw;
    const int kk = (int)(i & (ce - 1));
    long long q = i >> sh;
    const int n = (int)(q % j.Npad); q /= j.Npad;
    const int tap = (int)(q % taps);
    const int c = (int)(q / taps);
    const int k = c * ce + kk;
    float v = 0.f;
    if (n < j.Nn && k < j.Kk) {
      const int r = tap / j.kw, s = tap - r * j.kw;
      v = j.wm.wmode ? wval_phase(j.w, j.wm, n, k, r, s, j.Nn, j.Kk) : wval(j.w, n, perm_k(k, j.perm_C, j.perm_rr), r, s, j.Nn, j.Kk, j.kh, j.kw, j.flip);
    }
    if (j.kind == 1) ((unsigned short *)j.out)[i] = (unsigned short)(pack_bf16x2(v, 0.f) & 0xffffu);
    else ((float *)j.out)[i] = round_tf32(v);
  }
}

// Elements of one job a block of the multi-job kernel handles: sized per cache so that the launch has ~8 blocks per SM.  (A fixed
// 4096 left ESPCN's 80 k packed elements to 21 blocks of 16 dependent iterations each: 14 us at the tail of a 430 us step.)
constexpr int kPackChunkMin = 256, kPackChunkMax = 256 * 16;
// blk[b] = (job index, chunk index inside the job)
__global__ void __launch_bounds__(256) k_pack_multi(const PackJob *__restrict__ jobs, const int2 *__restrict__ blk, int chunk) {
  pdl_trigger();
  pdl_wait();
  const int2 bj = blk[blockIdx.x];
  const PackJob j = jobs[bj.x];
  const long long i0 = (long long)bj.y * chunk;
  const long long i1 = i0 + chunk < j.total ? i0 + chunk : j.total;
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) pack_job_element(j, i);
}
__global__ void __launch_bounds__(256) k_pack_one(PackJob j) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < j.total; i += (long long)gridDim.x * blockDim.x) pack_job_element(j, i);
}

long long *g_sl_trace = nullptr;
long long g_sl_trace_ctas = 0;
int g_sl_dbg = 0;

struct WCacheDev {
  std::vector<PackJob> jobs;
  PackJob *d_jobs = nullptr;
  int2 *d_blk = nullptr;
  int n_blk = 0;
  int chunk = kPackChunkMax;
};
std::mutex g_wc_mu;
bool g_wc_enabled = false;
WCacheDev g_wc[64];

inline bool wc_same(const PackJob &a, const PackJob &b) {
  return a.w == b.w && a.kind == b.kind && a.Nn == b.Nn && a.Kk == b.Kk && a.kh == b.kh && a.kw == b.kw && a.Npad == b.Npad &&
         a.chunks == b.chunks && a.flip == b.flip && a.perm_C == b.perm_C && a.perm_rr == b.perm_rr && a.NT == b.NT &&
         a.ntiles == b.ntiles && a.spairs == b.spairs && memcmp(&a.wm, &b.wm, sizeof(WMap)) == 0 && a.total == b.total;
}

// (re)build the device tables of one device's cache; not capturable (called from the eager warm-up only)
int wc_upload(WCacheDev &c) {
  std::vector<int2> blk;
  long long all = 0;
  for (const PackJob &j : c.jobs) all += j.total;
  long long chunk = (all / (148 * 8) + 255) / 256 * 256;
  c.chunk = (int)(chunk < kPackChunkMin ? kPackChunkMin : (chunk > kPackChunkMax ? kPackChunkMax : chunk));
  for (size_t j = 0; j < c.jobs.size(); ++j) {
    const long long nchunks = (c.jobs[j].total + c.chunk - 1) / c.chunk;
    for (long long q = 0; q < nchunks; ++q) blk.push_back(make_int2((int)j, (int)q));
  }
  if (c.d_jobs) SRB_CHECK_CUDA(cudaFree(c.d_jobs));
  if (c.d_blk) SRB_CHECK_CUDA(cudaFree(c.d_blk));
  c.d_jobs = nullptr; c.d_blk = nullptr; c.n_blk = 0;
  if (c.jobs.empty()) return SRB_OK;
  SRB_CHECK_CUDA(cudaMalloc(&c.d_jobs, c.jobs.size() * sizeof(PackJob)));
  SRB_CHECK_CUDA(cudaMalloc(&c.d_blk, blk.size() * sizeof(int2)));
  SRB_CHECK_CUDA(cudaMemcpy(c.d_jobs, c.jobs.data(), c.jobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice));
  SRB_CHECK_CUDA(cudaMemcpy(c.d_blk, blk.data(), blk.size() * sizeof(int2), cudaMemcpyHostToDevice));
  c.n_blk = (int)blk.size();
  return SRB_OK;
}

// Packed copy of job `t` (out unset): the cached buffer when the cache is on (created and packed now on a miss), else null.
// A miss while the stream is capturing also returns null: the caller packs into its workspace as without the cache.
int wc_lookup(PackJob t, size_t elem_bytes, cudaStream_t st, void **out) {
  *out = nullptr;
  if (!g_wc_enabled || !weight_cache_scope_allowed()) return SRB_OK;
  int dev = 0;
  SRB_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return SRB_OK;
  std::lock_guard<std::mutex> lk(g_wc_mu);
  if (!g_wc_enabled) return SRB_OK;
  WCacheDev &c = g_wc[dev];
  for (const PackJob &j : c.jobs)
    if (wc_same(j, t)) { *out = j.out; return SRB_OK; }
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  SRB_CHECK_CUDA(cudaStreamIsCapturing(st, &cs));
  if (cs != cudaStreamCaptureStatusNone) return SRB_OK;
  SRB_CHECK_CUDA(cudaMalloc(&t.out, (size_t)t.total * elem_bytes + 256));
  c.jobs.push_back(t);
  int rc = wc_upload(c);
  if (rc) return rc;
  int blocks = (int)((t.total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_pack_one<<<blocks, 256, 0, st>>>(t);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  *out = t.out;
  return SRB_OK;
}

struct SlPlan {
  SlArgs a;
  size_t smem;
  int n_tiles_n, Npad, ctas_per_sm, grid_x;
  size_t wpack_floats, xpack_floats;
};

inline double mma_cost(int N) {  // cycles of one SS-mode tf32 MMA, M=128 K=8 (tools/bench_umma.cu)
  double c1 = 32.0 + N / 4.0, c2 = N / 2.0;
  return c1 > c2 ? c1 : c2;
}

bool make_sl_plan(const Geom &g, SlPlan *pl, bool bf16 = false, int in_ps = 1) {
  SlArgs &a = pl->a;
  const bool c4 = g.Ci <= 4 && !bf16;
  const int celems = bf16 ? 64 : 32;  // channels per 128-byte slot
  const int sb = c4 ? 16 : 128;
  const int Npad = round_up_i(g.Co, 16);
  int NT0 = Npad;
  if (NT0 > 256) {
    NT0 = 256;
    while (Npad % NT0) NT0 -= 16;
  }
  const int spairs = (g.kw + 1) / 2;
  const int kextra = c4 ? 2 * spairs - 1 : g.kw - 1;
  const int chunks = c4 ? 1 : (g.Ci + celems - 1) / celems;
  const int kblocks = c4 ? g.kh * spairs : g.kh * g.kw * chunks;
  double best_t = -1.0;
  // N tile: the widest that fits.  Half-width tiles (two N tiles, more M tiles per band, half the streamed-weight traffic
  // per pixel) were measured SLOWER on the 256 -> 256 layers (101 vs 95 us, profiles/README.md), so kMaxNSplit stays 1.
  constexpr int kMaxNSplit = 1;
  for (int nsplit = 1; nsplit <= kMaxNSplit; ++nsplit) {
  if (nsplit == 2 && (c4 || NT0 < 128 || (NT0 / 2) % 16 != 0)) break;
  const int NT = NT0 / nsplit;
  const int n_tiles = Npad / NT;
  const double mma_per_tile = (c4 ? kblocks : kblocks * 4) * mma_cost(NT);
  for (int ctas = 2; ctas >= 1; --ctas) {
    const int smem_cap = ctas == 2 ? 111 * 1024 : 226 * 1024;
    const int col_cap = ctas == 2 ? 256 : 512;
    for (int t_bufs = 2; t_bufs >= 1; --t_bufs) {
      for (int MTB = (col_cap / (t_bufs * NT) < 4 ? col_cap / (t_bufs * NT) : 4); MTB >= 1; --MTB) {
        if ((g_sl_dbg & 8192) && MTB > 1) continue;   // debug: force the smallest bands
        if ((g_sl_dbg & 16384) && MTB > 2) continue;
        for (int wsplit = 1; wsplit <= 16; ++wsplit) {
          const int TW = (g.Wo + wsplit - 1) / wsplit;
          const int BW = TW + kextra;
          if (BW * in_ps > 256 || TW > 128 * MTB) continue;  // TMA box dims <= 256 (the un-shuffling traversal spans r x as many)
          int TH = (128 * MTB - TW) / BW + 1;
          if (TH > g.Ho) TH = g.Ho;
          const int bands_h = (g.Ho + TH - 1) / TH;
          TH = (g.Ho + bands_h - 1) / bands_h;
          const int BH = TH + g.kh - 1;
          if (BH * in_ps > 256) continue;
          const int bands_w = (g.Wo + TW - 1) / TW;
          int slots = BH * BW;
          const int need = 128 * MTB + (g.kh - 1) * BW + kextra;
          if (need > slots) slots = need;
          const int a_buf = round_up_i(slots * sb, 1024);
          int b_stage, b_stages, a_bufs = 2;  // two A buffers: the next chunk / next band loads under the current MMAs
          int b_res = 0;
          if (c4) {
            b_stage = kblocks * NT * 32; b_stages = 1;
          } else {
            b_stage = NT * 128;
            b_stages = (smem_cap - 3072 - a_bufs * a_buf) / b_stage;
            if (b_stages >= kblocks) { b_stages = kblocks; b_res = 1; }  // all weights stay resident
            else {
              if (b_stages < 4) { a_bufs = 1; b_stages = (smem_cap - 3072 - a_buf) / b_stage; }
              if (b_stages >= kblocks) { b_stages = kblocks; b_res = 1; }
              else if (b_stages > kMaxBStages) b_stages = kMaxBStages;
              if (b_stages < 2) continue;
            }
          }
          const size_t smem = (size_t)a_bufs * a_buf + (size_t)b_stages * b_stage + 1024 + 512 + 1024;
          if (smem > (size_t)smem_cap) continue;
          // time model per real output pixel (cycles on one SM)
          const double eff = (double)g.Ho * g.Wo / ((double)bands_h * bands_w * MTB * 128);
          double t_mma = mma_per_tile / 128.0 / eff;
          if (!c4 && !b_res) {  // a streamed weight ring delivers one K-block per (TMA round trip / stages)
            const double ring = 2400.0 / b_stages * kblocks / (MTB * 128.0) / eff;
            if (ring > t_mma) t_mma = ring;
          }
          const double a_bytes = (double)chunks * BH * BW * sb / ((double)TH * TW);
          const double b_bytes = (c4 || b_res) ? 0.0 : (double)kblocks * NT * 128 / (eff * MTB * 128.0);
          {  // TMA writes share the SMEM port with the MMAs' operand reads (measured: 50 -> 55..61 cycles per MMA in-kernel,
             // profiles/r1_trace_conv_sl.txt): stretch the MMA time by written / read bytes per pixel
            const double mma_smem_per_px = (c4 ? kblocks : kblocks * 4) * (4096.0 + NT * 32.0) / 128.0 / eff;
            t_mma *= 1.0 + (a_bytes + b_bytes) / mma_smem_per_px;
          }
          const double t_mem = (a_bytes + b_bytes + 4.0 * NT) / 48.0;  // ~48 B/clk/SM of L2->SM + store bandwidth
          double t = t_mma > t_mem ? t_mma : t_mem;
          t += 600.0 / ((double)TH * TW);  // per-band hand-offs (accumulator swap, first-MMA latency)
          if (t_bufs == 1) t *= (ctas == 2 ? 1.1 : 1.3);  // the epilogue is only hidden by a co-resident CTA, if any
          if (a_bufs == 1) t *= 1.15;                       // operand loads serialise with the MMAs
          // wave quantisation of the persistent grid: the slowest CTA walks ceil(work / slots) bands
          const long long work = (long long)g.N * bands_h * bands_w * (Npad / NT);
          const long long slots_sm = 148LL * ctas;
          const long long waves = (work + slots_sm - 1) / slots_sm;
          if (waves < 4) t *= (double)(waves * slots_sm) / (double)(work > 0 ? work : 1);  // many waves: tail effects are small
          // fixed cost per CTA (TMEM alloc, barrier init, first operand + resident weight fetch, drain of the last
          // epilogue), amortised over the pixels one CTA produces
          const long long waves_sm = (work + 147) / 148;  // bands per SM, however many CTAs share it
          t += (2500.0 + (b_res ? 0.02 * b_stage * b_stages : 0.0)) / ((double)waves_sm * TH * TW);
          t *= (double)n_tiles;  // every pixel is visited once per N tile
          if (best_t < 0 || t < best_t) {
            best_t = t;
            a.TH = TH; a.TW = TW; a.BW = BW; a.BH = BH; a.bands_h = bands_h; a.bands_w = bands_w;
            a.MTB = MTB; a.NT = NT; a.chunks = chunks; a.spairs = spairs;
            a.a_bufs = a_bufs; a.a_buf_bytes = a_buf; a.a_tx_bytes = BH * BW * sb;
            a.b_stages = b_stages; a.b_stage_bytes = b_stage; a.b_resident = b_res;
            a.t_bufs = t_bufs;
            int tc = 32;
            while (tc < t_bufs * MTB * NT) tc <<= 1;
            a.tmem_cols = tc;
            pl->smem = smem;
            pl->ctas_per_sm = ctas;
          }
        }
      }
    }
  }
  }  // nsplit
  if (best_t < 0) return false;
  const int NT = a.NT;
  a.N = g.N; a.Ho = g.Ho; a.Wo = g.Wo; a.Co = g.Co; a.kh = g.kh; a.kw = g.kw; a.pad = g.pad; a.c4 = c4 ? 1 : 0;
  a.ps = g.ps;
  a.in_bf16 = bf16 ? 1 : 0;
  a.out_bf16 = 0;
  a.chunk_elems = celems;
  a.v8h = 0;
  a.in_ps = in_ps;
  a.in_cpb = in_ps > 1 ? g.Ci / (in_ps * in_ps) / celems : 0;
  a.pad_w = g.pad;
  a.rs = 0;
  a.kb_valid = kblocks;
  for (int i = 0; i < 16; ++i) a.phase_mask[i] = 0xffff;
  pl->Npad = Npad;
  pl->n_tiles_n = Npad / NT;
  {  // persistent grid: as many CTAs as fit on the chip; CTA x walks bands x, x + grid, ...
    const long long num_bands = (long long)g.N * a.bands_h * a.bands_w;
    long long slots = (long long)pl->ctas_per_sm * 148 / pl->n_tiles_n;
    if (slots < 1) slots = 1;
    long long P = num_bands < slots ? num_bands : slots;
    if (P < 1) P = 1;
    pl->grid_x = (int)P;
  }
  // share the streamed weight ring between the two CTAs of a cluster (see SlArgs::cl)
  // MEASURED SLOWER (r2f: EDSR-256 bf16 fprop 60 -> 64 us, VDSR body 185 -> 203 us): the L2 slices run at 17 % and the limit
  // is the SMEM operand path of the MMAs, which multicast does not relieve; kept behind debug flag 64 for experiments.
  a.cl = (!c4 && !a.b_resident && pl->ctas_per_sm == 1 && pl->grid_x >= 2 && (g_sl_dbg & 64)) ? 2 : 1;
  if (a.cl == 2) pl->grid_x &= ~1;
  pl->wpack_floats = c4 ? (size_t)pl->n_tiles_n * kblocks * NT * 8 : (size_t)kblocks * Npad * 32;
  pl->xpack_floats = c4 ? (size_t)g.N * g.Hi * g.Wi * 4 : 0;
  return true;
}


// ---- row-stacked kernel: plan + launch ------------------------------------------------------------------------------------------
struct RsPlan {
  RsArgs a;
  size_t smem;
  int grid, Npad, BWs;
  double lane_eff;
};

// Layers the row-stacked kernel takes: tf32 operands from an NHWC fp32 tensor (Cin >= 8), stride 1, no input un-shuffle, a single
// N tile with kh * NT <= 256 and every weight resident in shared memory next to >= 3 row stages.
bool make_rs_plan(const Geom &g, RsPlan *pl) {
  if (g_sl_dbg & 128) return false;  // debug: force the slot-linear kernel
  if (g.st != 1 || g.Ci <= 4 || g.kh > 9 || g.kw > 9) return false;
  const int NT = round_up_i(g.Co, 16);
  if (g.kh * NT > 256) return false;
  const int chunks = (g.Ci + 31) / 32;
  RsArgs &a = pl->a;
  memset(&a, 0, sizeof(a));
  int CT = 1, TW = g.Wo;
  if (g.Wo > 128) { CT = (g.Wo + 127) / 128; TW = (g.Wo + CT - 1) / CT; }
  const int BWs = TW + g.kw - 1;
  int G = 1;
  if (CT == 1) {
    G = (128 + g.kw - 1) / BWs;
    if (G > g.N) G = g.N;
    if (G < 1) G = 1;
    while (G > 1 && (G - 1) * BWs + TW > 128) --G;
  }
  if (BWs > 256) return false;
  pl->lane_eff = (double)G * TW / 128.0;
  if (pl->lane_eff < 0.55) return false;
  const int slots = G * BWs > 128 ? G * BWs : 128;
  a.chunk_bytes = round_up_i(slots * 128, 1024);
  a.stage_bytes = chunks * a.chunk_bytes;
  a.stage_tx = G * BWs * 128;
  a.bcs_bytes = g.kh * NT * 128;
  const long long b_bytes = (long long)chunks * g.kw * a.bcs_bytes;
  const long long fixed = b_bytes + 1024 /*gap*/ + 1024 /*align*/ + 2048 /*barriers, bias*/;
  // two row streams when each still gets >= 2 row stages and a TMEM ring of >= 2 * kh blocks (else the window wraps too often)
  const int r_all = 512 / NT;
  a.streams = (!(g_sl_dbg & 65536) && (r_all / 2) >= 2 * g.kh && (226 * 1024 - fixed) / (2 * a.stage_bytes) >= 2) ? 2 : 1;
  long long S = (226 * 1024 - fixed) / ((long long)a.streams * a.stage_bytes);
  if (S < 2) return false;
  if (S > kRsMaxStages) S = kRsMaxStages;
  a.S = (int)S;
  a.chunks = chunks;
  a.kv_last = (g.Ci - (chunks - 1) * 32 + 7) / 8;
  a.G = G; a.TW = TW; a.CT = CT;
  a.R = r_all / a.streams < kRsMaxBlocks ? r_all / a.streams : kRsMaxBlocks;
  if (a.R < g.kh + 1) return false;
  const long long IG = (g.N + G - 1) / G;
  a.total_rows = IG * CT * g.Ho;
  // a CTA pays kh-1 halo rows and ~10 us of fixed cost per launch: small problems (EDSR-64 / SRGAN-G bodies: 2-3 rows per CTA)
  // measured faster on the slot-linear kernel (19.7 vs 20.9 us)
  if (a.total_rows < 148LL * 8 && !(g_sl_dbg & 1024)) return false;
  SlArgs &e = a.e;
  e.N = g.N; e.Ho = g.Ho; e.Wo = g.Wo; e.Co = g.Co; e.kh = g.kh; e.kw = g.kw; e.pad = g.pad; e.pad_w = g.pad;
  e.NT = NT; e.BW = BWs; e.ps = g.ps; e.rs = 1; e.MTB = 1; e.chunk_elems = 32;
  pl->smem = (size_t)a.streams * a.S * a.stage_bytes + (size_t)fixed;
  pl->grid = (int)(a.total_rows < 148 ? a.total_rows : 148);
  pl->Npad = NT;
  pl->BWs = BWs;
  return pl->grid >= 1;
}

bool rs_usable(const Geom &g, const T4 &in, const T4 &out, const Epi &epi, const ConvOpt &opt, RsPlan *pl) {
  if (in.dt != SRB_F32 || out.dt != SRB_F32) return false;
  if (opt.in_ps != 1 || opt.wmode != 0 || (opt.pad_w >= 0 && opt.pad_w != g.pad) || opt.in_h > 0 || opt.in_w > 0) return false;
  (void)epi;
  return make_rs_plan(g, pl);
}

int tc_conv_rs_launch(const Geom &g, const T4 &in, const float *w, bool flip_transpose, const T4 &out, const Epi &epi_in, void *ws,
                      size_t ws_bytes, cudaStream_t st, const ConvOpt &opt, RsPlan &pl) {
  Epi epi = epi_in;
  RsArgs &a = pl.a;
  const size_t wpack_floats = (size_t)a.chunks * g.kw * g.kh * pl.Npad * 32;
  const size_t loss_bytes = epi.loss_kind ? (size_t)kMaxLossCtas * 8 * sizeof(float) + 256 : 0;
  const size_t need = wpack_floats * sizeof(float) + 512 + loss_bytes;
  uintptr_t wsp = ((uintptr_t)ws + 255) & ~(uintptr_t)255;
  SRB_REQUIRE(ws && wsp + need <= (uintptr_t)ws + ws_bytes, SRB_EWORKSPACE, "tc_conv_rs workspace: need %zu bytes, have %zu", need + 256,
              ws_bytes);
  float *wp = (float *)wsp;
  if (epi.loss_kind) {
    SRB_REQUIRE(opt.loss_out != nullptr && pl.grid * 4 * kRsEpiPerQuad <= kMaxLossCtas * 8, SRB_EINVAL, "fused loss: bad arguments");
    SRB_REQUIRE(epi.act == SRB_ACT_NONE && !epi.residual.p && !epi.preact.p && !epi.mask.p && !epi.bits_out && !epi.bits_in &&
                    epi.target.p && (epi.target.dt == SRB_F32 || epi.target.dt == SRB_U8),
                SRB_EUNSUPPORTED, "fused loss: the last conv must have fp32 output, no activation and no residual");
    if (g.ps == 4 && epi.dz_unshuf) {
      auto ok = [](const T4 &t) {
        return !t.p || t.dt == SRB_U8 || (t.sw == 1 && (t.sh & 3) == 0 && (t.sc & 3) == 0 && (t.sn & 3) == 0 && (((uintptr_t)t.p) & 15) == 0);
      };
      SRB_REQUIRE((g.Co & 15) == 0 && ok(out) && ok(epi.target) && (((uintptr_t)epi.dz_unshuf) & 31) == 0, SRB_EUNSUPPORTED,
                  "fused loss with PixelShuffle(4): NCHW-contiguous y / target, Cout*16 channels");
    } else {
      SRB_REQUIRE(g.ps == 1 && epi.dz.p && epi.dz.dt == SRB_F32 && !epi.dz_unshuf, SRB_EUNSUPPORTED,
                  "fused loss: PixelShuffle(4) with an un-shuffled gradient, or no PixelShuffle with the gradient in y's layout");
    }
    epi.loss_part = (float *)(((uintptr_t)wp + wpack_floats * sizeof(float) + 255) & ~(uintptr_t)255);
  }
  {
    PackJob job;
    memset(&job, 0, sizeof(job));
    job.w = w; job.kind = 2; job.Nn = g.Co; job.Kk = g.Ci; job.kh = g.kh; job.kw = g.kw; job.Npad = pl.Npad; job.chunks = a.chunks;
    job.flip = flip_transpose ? 1 : 0; job.total = (long long)wpack_floats;
    void *cached = nullptr;
    int rc_c = wc_lookup(job, 4, st, &cached);
    if (rc_c) return rc_c;
    if (cached) {
      wp = (float *)cached;
    } else {
      const long long total = (long long)wpack_floats;
      int blocks = (int)((total + 255) / 256);
      if (blocks > 148 * 8) blocks = 148 * 8;
      launch_pdl(k_pack_w_rs, dim3(blocks), dim3(256), 0, st, w, wp, g.Co, g.Ci, g.kh, g.kw, pl.Npad, a.chunks, flip_transpose ? 1 : 0);
      count_launch();
      SRB_CHECK_CUDA(cudaGetLastError());
    }
  }
  CUtensorMap mapA, mapB;
  {  // (C, W, N, H): a box of {32, BWs, G, 1} lands as [image][slot][32 ch] = one M tile row
    cuuint64_t dims[4] = {(cuuint64_t)g.Ci, (cuuint64_t)g.Wi, (cuuint64_t)g.N, (cuuint64_t)g.Hi};
    cuuint64_t strides[3] = {(cuuint64_t)in.sw * 4, (cuuint64_t)in.sn * 4, (cuuint64_t)in.sh * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)pl.BWs, (cuuint32_t)a.G, 1};
    int rc = encode_tiled(&mapA, in.p, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[3] = {32, (cuuint64_t)(g.kh * pl.Npad), (cuuint64_t)(a.chunks * g.kw)};
    cuuint64_t strides[2] = {128, (cuuint64_t)g.kh * pl.Npad * 128};
    cuuint32_t box[3] = {32, (cuuint32_t)(g.kh * pl.Npad), 1};
    int rc = encode_tiled(&mapB, wp, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  a.e.out = out;
  a.e.epi = epi;
  a.e.dbg = g_sl_dbg;
  a.e.trace = g_sl_trace_ctas >= 64 ? g_sl_trace : nullptr;
  {
    auto ok32 = [](const T4 &t) {
      return !t.p || ((((uintptr_t)t.p) & 31) == 0 && (t.sn & 7) == 0 && (t.sh & 7) == 0 && (t.sw & 7) == 0);
    };
    a.e.v8 = (ok32(out) && ok32(epi.residual) && ok32(epi.preact)) ? 1 : 0;
  }
  {
    static std::atomic<unsigned long long> attr_done{0};
    int rc = ensure_kernel_attrs(k_conv_rs, attr_done, kMaxSmemBytes, true);
    if (rc) return rc;
  }
  SRB_CHECK_CUDA(launch_pdl(k_conv_rs, dim3(pl.grid), dim3(kRsThreads), pl.smem, st, mapA, mapB, a));
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  if (epi.loss_kind) {
    const double numel = (double)g.N * g.Ho * g.Wo * g.Co;
    launch_pdl(k_sl_loss_finish, dim3(1), dim3(256), 0, st, epi.loss_part, pl.grid * 4 * kRsEpiPerQuad, (float)(1.0 / numel), opt.loss_out);
    count_launch();
    SRB_CHECK_CUDA(cudaGetLastError());
  }
  return SRB_OK;
}

}  // namespace

void tc_conv_set_trace(long long *buf, long long max_ctas) { g_sl_trace = buf; g_sl_trace_ctas = max_ctas; }
void tc_conv_set_dbg(int flags) { g_sl_dbg = flags; }
int tc_conv_get_dbg() { return g_sl_dbg; }

int tc_weight_cache_enable(int on) {
  std::lock_guard<std::mutex> lk(g_wc_mu);
  g_wc_enabled = on != 0;
  if (!on) {
    int cur = 0;
    cudaGetDevice(&cur);
    for (int d = 0; d < 64; ++d) {
      WCacheDev &c = g_wc[d];
      if (c.jobs.empty() && !c.d_jobs) continue;
      cudaSetDevice(d);
      cudaDeviceSynchronize();
      for (PackJob &j : c.jobs) cudaFree(j.out);
      c.jobs.clear();
      if (c.d_jobs) cudaFree(c.d_jobs);
      if (c.d_blk) cudaFree(c.d_blk);
      c.d_jobs = nullptr; c.d_blk = nullptr; c.n_blk = 0;
    }
    cudaSetDevice(cur);
  }
  return SRB_OK;
}

int tc_weight_cache_repack(cudaStream_t st) {
  int dev = 0;
  SRB_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return SRB_OK;
  std::lock_guard<std::mutex> lk(g_wc_mu);
  WCacheDev &c = g_wc[dev];
  if (!g_wc_enabled || c.n_blk == 0) return SRB_OK;
  k_pack_multi<<<c.n_blk, 256, 0, st>>>(c.d_jobs, c.d_blk, c.chunk);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

int tc_weight_cache_entries() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  std::lock_guard<std::mutex> lk(g_wc_mu);
  return (int)g_wc[dev].jobs.size();
}

bool tc_conv_supported(const Geom &g, const T4 &in, const T4 &out, bool /*dgrad*/, int in_ps, int /*pad_w*/) {
  if (g.st != 1 || g.N <= 0) return false;
  if (in_ps > 1) {  // `in` is PixelShuffle_r of the logical input: needs whole channel chunks per sub-pixel phase
    const int celems = in.dt == SRB_BF16 ? 64 : 32, rr = in_ps * in_ps;
    if (in_ps > 8 || g.Ci % rr != 0 || (g.Ci / rr) % celems != 0) return false;
  }
  if (g.kh > 16 || g.kw > 16) return false;
  if ((long long)g.N * g.Hi * g.Wi * (g.Ci > 4 ? g.Ci : 4) >= (1LL << 40)) return false;
  if ((long long)g.N * ((g.Ho + 0) * (long long)g.Wo) >= (1LL << 31)) return false;
  if (in.dt == SRB_BF16) {
    if (g.Ci % 8 != 0 || g.Ci < 8) return false;                       // TMA: 16-byte pixel stride
    if (in.sc != 1) return false;                                       // channels_last activations
    if ((in.sw % 8) || (in.sh % 8) || (in.sn % 8)) return false;       // TMA strides: multiples of 16 B
    if (((uintptr_t)in.p) & 15) return false;
    SlPlan ph;
    return make_sl_plan(g, &ph, true, in_ps);
  }
  if (g.Ci > 4) {
    if (g.Ci % 4 != 0 || g.Ci < 8) return false;                       // TMA: 16-byte pixel stride
    if (in.sc != 1) return false;                                       // channels_last activations
    if ((in.sw % 4) || (in.sh % 4) || (in.sn % 4)) return false;       // TMA strides: multiples of 16 B
    if (((uintptr_t)in.p) & 15) return false;
  }
  SlPlan p;
  return make_sl_plan(g, &p, false, in_ps);
}

size_t tc_conv_ws_bytes(const Geom &g) {
  SlPlan p;
  if (g.st != 1 || !make_sl_plan(g, &p)) return 0;
  // dgrad of the same layer swaps Ci/Co: take the larger of both packings
  Geom gd = g;
  gd.Ci = g.Co; gd.Co = g.Ci; gd.Hi = g.Ho; gd.Wi = g.Wo; gd.Ho = g.Hi; gd.Wo = g.Wi; gd.pad = g.kh - 1 - g.pad; gd.ps = 1;
  size_t a = (p.wpack_floats + p.xpack_floats) * sizeof(float) + 1024;
  SlPlan pd;
  if (gd.pad >= 0 && make_sl_plan(gd, &pd)) {
    size_t b = (pd.wpack_floats + pd.xpack_floats) * sizeof(float) + 1024;
    if (b > a) a = b;
  }
  return a;
}

int tc_conv_describe(const Geom &g, char *buf, size_t n, bool bf16) {
  RsPlan rp;
  if (!bf16 && make_rs_plan(g, &rp)) {
    const RsArgs &r = rp.a;
    return snprintf(buf, n,
                    "conv_rs row-stacked: %d image(s) x %d slots per M tile (%d output columns, lane use %.2f), %d strip(s)/image, N = %d x %d = %d, "
                    "chunks %d (last: %d K-steps), %d stream(s) x %d row stages x %d B, weights %d B resident, TMEM ring %d blocks, smem %zu B, "
                    "%lld output rows on grid %d",
                    r.G, rp.BWs, r.TW, rp.lane_eff, r.CT, g.kh, r.e.NT, g.kh * r.e.NT, r.chunks, r.kv_last, r.streams, r.S, r.stage_bytes,
                    r.chunks * g.kw * r.bcs_bytes, r.R, rp.smem, r.total_rows, rp.grid);
  }
  SlPlan pl;
  if (!make_sl_plan(g, &pl, bf16)) return snprintf(buf, n, "conv_sl: no plan");
  const SlArgs &a = pl.a;
  return snprintf(buf, n,
                  "conv_sl %s: band %dx%d (halo %dx%d), %dx%d bands/img, MTB %d, NT %d x%d, chunks %d, a_bufs %d x %d B, "
                  "b_stages %d%s x %d B, smem %zu B, tmem %d cols (%d acc bufs), %d bands on grid %d x %d (%d CTA/SM)",
                  a.c4 ? "c4" : (a.in_bf16 ? "generic-bf16" : "generic"), a.TH, a.TW, a.BH, a.BW, a.bands_h, a.bands_w, a.MTB, a.NT, pl.n_tiles_n, a.chunks,
                  a.a_bufs, a.a_buf_bytes, a.b_stages, (a.c4 || a.b_resident) ? " (resident)" : "", a.b_stage_bytes, pl.smem, a.tmem_cols, a.t_bufs,
                  g.N * a.bands_h * a.bands_w, pl.grid_x, pl.n_tiles_n, pl.ctas_per_sm);
}

int tc_conv_gather(const Geom &g, const T4 &in, const float *w, bool flip_transpose, const T4 &out, const Epi &epi_in,
                   void *ws, size_t ws_bytes, cudaStream_t st, const ConvOpt &opt) {
  const PdlScope pdl_scope(2.0 * g.N * g.Ho * g.Wo * (double)g.Co * g.Ci * g.kh * g.kw < 2.0e10);
  {
    RsPlan rp;
    if (rs_usable(g, in, out, epi_in, opt, &rp)) return tc_conv_rs_launch(g, in, w, flip_transpose, out, epi_in, ws, ws_bytes, st, opt, rp);
  }
  SlPlan pl;
  Epi epi = epi_in;
  const int in_ps = opt.in_ps;
  float *loss_out = opt.loss_out;
  const bool bf_in = in.dt == SRB_BF16, bf_out = out.dt == SRB_BF16;
  SRB_REQUIRE(make_sl_plan(g, &pl, bf_in, in_ps), SRB_EUNSUPPORTED, "tc_conv: no band plan");
  SRB_REQUIRE(in_ps == 1 || !pl.a.c4, SRB_EUNSUPPORTED, "tc_conv: un-shuffled input needs the generic operand flavour");
  const int perm_C = in_ps > 1 ? g.Ci / (in_ps * in_ps) : 0, perm_rr = in_ps * in_ps;
  WMap wm;
  wm.wmode = opt.wmode; wm.st = opt.st; wm.pad0 = opt.pad0; wm.kh0 = opt.kh0; wm.kw0 = opt.kw0;
  wm.dmin_r = opt.dmin_r; wm.dmin_s = opt.dmin_s; wm.ra = opt.ra; wm.rb = opt.rb; wm.tmax_a = opt.tmax_a; wm.tmax_b = opt.tmax_b;
  wm.C = perm_C > 0 ? perm_C : g.Ci;
  if (opt.pad_w >= 0) pl.a.pad_w = opt.pad_w;
  if (opt.wmode == 1) {  // which taps exist for which input phase
    SRB_REQUIRE(in_ps > 1 && in_ps <= 4 && g.kh * g.kw <= 16 && !pl.a.c4, SRB_EUNSUPPORTED, "strided conv: unsupported phase geometry");
    int valid = 0;
    for (int ph = 0; ph < in_ps * in_ps; ++ph) {
      const int pa = ph / in_ps, pb = ph % in_ps;
      unsigned m = 0;
      for (int tr = 0; tr < g.kh; ++tr)
        for (int ts = 0; ts < g.kw; ++ts) {
          const int r = opt.st * (tr + opt.dmin_r) + pa + opt.pad0, s2 = opt.st * (ts + opt.dmin_s) + pb + opt.pad0;
          if (r >= 0 && r < opt.kh0 && s2 >= 0 && s2 < opt.kw0) m |= 1u << (tr * g.kw + ts);
        }
      pl.a.phase_mask[ph] = (unsigned short)m;
      valid += __builtin_popcount(m) * pl.a.in_cpb;
    }
    pl.a.kb_valid = valid;
  }
  SRB_REQUIRE((!epi.residual.p || epi.residual.dt == out.dt) && (!epi.preact.p || epi.preact.dt == out.dt), SRB_EINVAL,
              "residual / preact must have the dtype of the output");
  SRB_REQUIRE(!epi.mask.p || epi.mask.dt == SRB_F32 || !bf_out, SRB_EUNSUPPORTED, "float relu_mask with bf16 tensors (use relu_bits)");
  SlArgs &a = pl.a;
  const size_t loss_bytes = epi.loss_kind ? (size_t)kMaxLossCtas * 8 * sizeof(float) + 256 : 0;
  const size_t need = (pl.wpack_floats + pl.xpack_floats) * sizeof(float) + 512 + loss_bytes;
  uintptr_t wsp = ((uintptr_t)ws + 255) & ~(uintptr_t)255;
  SRB_REQUIRE(ws && wsp + need <= (uintptr_t)ws + ws_bytes, SRB_EWORKSPACE, "tc_conv workspace: need %zu bytes, have %zu",
              need + 256, ws_bytes);
  float *wp = (float *)wsp;
  float *xp = (float *)((wsp + pl.wpack_floats * sizeof(float) + 255) & ~(uintptr_t)255);
  if (epi.loss_kind) {
    SRB_REQUIRE(loss_out != nullptr && (long long)pl.grid_x * pl.n_tiles_n * 16 <= kMaxLossCtas * 8, SRB_EINVAL, "fused loss: bad arguments");
    SRB_REQUIRE(epi.act == SRB_ACT_NONE && !epi.residual.p && !epi.preact.p && !epi.mask.p && !epi.bits_out && !epi.bits_in &&
                    out.dt == SRB_F32 && epi.target.p && (epi.target.dt == SRB_F32 || epi.target.dt == SRB_U8),
                SRB_EUNSUPPORTED, "fused loss: the last conv must have fp32 output, no activation and no residual");
    if (g.ps == 4 && epi.dz_unshuf) {
      auto ok = [](const T4 &t) {
        return !t.p || t.dt == SRB_U8 || (t.sw == 1 && (t.sh & 3) == 0 && (t.sc & 3) == 0 && (t.sn & 3) == 0 && (((uintptr_t)t.p) & 15) == 0);
      };
      SRB_REQUIRE((g.Co & 15) == 0 && ok(out) && ok(epi.target) && (((uintptr_t)epi.dz_unshuf) & 31) == 0, SRB_EUNSUPPORTED,
                  "fused loss with PixelShuffle(4): NCHW-contiguous y / target, Cout*16 channels");
    } else {
      SRB_REQUIRE(g.ps == 1 && epi.dz.p && epi.dz.dt == SRB_F32 && !epi.dz_unshuf, SRB_EUNSUPPORTED,
                  "fused loss: PixelShuffle(4) with an un-shuffled gradient, or no PixelShuffle with the gradient in y's layout");
    }
    epi.loss_part = (float *)(((uintptr_t)xp + pl.xpack_floats * sizeof(float) + 255) & ~(uintptr_t)255);
  }

  // 1. operands: weights (and, for c4, the NHWC4 image)
  {
    PackJob job;
    memset(&job, 0, sizeof(job));
    job.w = w; job.kind = a.c4 ? 3 : (bf_in ? 1 : 0);
    job.Nn = g.Co; job.Kk = g.Ci; job.kh = g.kh; job.kw = g.kw; job.Npad = pl.Npad; job.chunks = a.chunks; job.flip = flip_transpose ? 1 : 0;
    job.perm_C = perm_C; job.perm_rr = perm_rr; job.NT = a.NT; job.ntiles = pl.n_tiles_n; job.spairs = a.spairs; job.wm = wm;
    job.total = a.c4 ? (long long)pl.wpack_floats : (long long)a.chunks * g.kh * g.kw * pl.Npad * (bf_in ? 64 : 32);
    void *cached = nullptr;
    int rc_c = wc_lookup(job, bf_in ? 2 : 4, st, &cached);
    if (rc_c) return rc_c;
    if (cached) wp = (float *)cached;
    const long long total = (long long)pl.wpack_floats;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (cached) {
    } else if (a.c4)
      launch_pdl(k_pack_w_c4, dim3(blocks), dim3(256), 0, st, w, wp, g.Co, g.Ci, g.kh, g.kw, a.NT, pl.n_tiles_n, a.spairs, flip_transpose ? 1 : 0);
    else if (bf_in)
      launch_pdl(k_pack_w_sl_h, dim3(blocks), dim3(256), 0, st, w, (unsigned short *)wp, g.Co, g.Ci, g.kh, g.kw, pl.Npad, a.chunks, flip_transpose ? 1 : 0,
                                            perm_C, perm_rr, wm);
    else
      launch_pdl(k_pack_w_sl, dim3(blocks), dim3(256), 0, st, w, wp, g.Co, g.Ci, g.kh, g.kw, pl.Npad, a.chunks, flip_transpose ? 1 : 0, perm_C, perm_rr, wm);
    if (!cached) count_launch();
    SRB_CHECK_CUDA(cudaGetLastError());
    if (a.c4) {
      const long long px = (long long)g.N * g.Hi * g.Wi;
      int pb = (int)((px + 255) / 256);
      if (pb > 148 * 16) pb = 148 * 16;
      launch_pdl(k_pack_nhwc4, dim3(pb), dim3(256), 0, st, in, (float4 *)xp, g.N, g.Ci, g.Hi, g.Wi);
      count_launch();
      SRB_CHECK_CUDA(cudaGetLastError());
    }
  }

  // 2. tensor maps
  CUtensorMap mapA, mapB;
  if (a.c4) {
    cuuint64_t dims[4] = {4, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {16, (cuuint64_t)g.Wi * 16, (cuuint64_t)g.Hi * g.Wi * 16};
    cuuint32_t box[4] = {4, (cuuint32_t)a.BW, (cuuint32_t)a.BH, 1};
    int rc = encode_tiled(&mapA, xp, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    mapB = mapA;  // unused
    a.wpack = wp;
  } else {
    {
      // in_ps > 1: the map describes the SHUFFLED tensor (C, W*r, H*r, N) and is traversed with element strides (1, r, r, 1):
      // a box of (BW*r) x (BH*r) delivers BW x BH pixels of one sub-pixel phase
      const cuuint64_t r = (cuuint64_t)in_ps;
      cuuint64_t dims[4] = {(cuuint64_t)(in_ps > 1 ? perm_C : g.Ci), (cuuint64_t)(opt.in_w > 0 ? opt.in_w : g.Wi * (int)r),
                            (cuuint64_t)(opt.in_h > 0 ? opt.in_h : g.Hi * (int)r), (cuuint64_t)g.N};
      const cuuint64_t es = bf_in ? 2 : 4;
      cuuint64_t strides[3] = {(cuuint64_t)in.sw * es, (cuuint64_t)in.sh * es, (cuuint64_t)in.sn * es};
      cuuint32_t box[4] = {(cuuint32_t)a.chunk_elems, (cuuint32_t)(a.BW * in_ps), (cuuint32_t)(a.BH * in_ps), 1};
      cuuint32_t estr[4] = {1, (cuuint32_t)in_ps, (cuuint32_t)in_ps, 1};
      int rc = encode_tiled(&mapA, in.p, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, bf_in, in_ps > 1 ? estr : nullptr);
      if (rc) return rc;
    }
    {
      cuuint64_t dims[3] = {(cuuint64_t)a.chunk_elems, (cuuint64_t)pl.Npad, (cuuint64_t)(a.chunks * g.kh * g.kw)};
      cuuint64_t strides[2] = {128, (cuuint64_t)pl.Npad * 128};
      cuuint32_t box[3] = {(cuuint32_t)a.chunk_elems, (cuuint32_t)(a.NT / a.cl), 1};
      int rc = encode_tiled(&mapB, wp, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, bf_in);
      if (rc) return rc;
    }
    a.wpack = nullptr;
  }
  a.out = out;
  a.epi = epi;
  {
    auto ok32 = [](const T4 &t) {
      return !t.p || ((((uintptr_t)t.p) & 31) == 0 && (t.sn & 7) == 0 && (t.sh & 7) == 0 && (t.sw & 7) == 0);
    };
    a.v8 = (ok32(out) && ok32(epi.residual) && ok32(epi.preact)) ? 1 : 0;
    auto ok32h = [](const T4 &t) {  // bf16: strides in 2-byte elements
      return !t.p || ((((uintptr_t)t.p) & 31) == 0 && (t.sn & 15) == 0 && (t.sh & 15) == 0 && (t.sw & 15) == 0);
    };
    a.out_bf16 = bf_out ? 1 : 0;
    a.v8h = (bf_out && ok32h(out) && ok32h(epi.residual) && ok32h(epi.preact)) ? 1 : 0;
  }
  a.dbg = g_sl_dbg;
  a.trace = ((long long)pl.grid_x * pl.n_tiles_n <= g_sl_trace_ctas) ? g_sl_trace : nullptr;

  {
    static std::atomic<unsigned long long> attr_done{0};
    int rc = ensure_kernel_attrs(k_conv_sl, attr_done, kMaxSmemBytes, true);
    if (rc) return rc;
  }
  dim3 grid((unsigned)pl.grid_x, (unsigned)pl.n_tiles_n);
  // one CTA per SM: 16 epilogue warps (more stores in flight); two per SM: 8 each (registers: 96 x 576 x 2 would not fit)
  const int sl_threads = (pl.ctas_per_sm == 1 && !(g_sl_dbg & 32768)) ? kThreadsWide : kThreads;
  if (a.cl > 1) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(sl_threads);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)a.cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SRB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k_conv_sl, mapA, mapB, a));
  } else {
    SRB_CHECK_CUDA(launch_pdl(k_conv_sl, grid, dim3(sl_threads), pl.smem, st, mapA, mapB, a));
  }
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  if (epi.loss_kind) {
    const double numel = (double)g.N * g.Ho * g.Wo * g.Co;
    launch_pdl(k_sl_loss_finish, dim3(1), dim3(256), 0, st, epi.loss_part, (int)(grid.x * grid.y * ((sl_threads - 64) / 32)), (float)(1.0 / numel), loss_out);
    count_launch();
    SRB_CHECK_CUDA(cudaGetLastError());
  }
  return SRB_OK;
}

}  // namespace srb
