// tcgen05 weight-gradient kernel for sm_100a (stride-1 Conv2d, NHWC fp32 activations holding tf32 values).
//
//   dW[co][ci][r][s] = sum_{n,oy,ox} dz[n,oy,ox,co] * x[n,oy+r-pad,ox+s-pad,ci]
//
// as a GEMM whose K dimension is the PIXEL index:  D[(s,ci), co] += sum_p X_{r,s}[p, ci] * dZ[p, co].
// Both operands are "MN-major" (channels contiguous, K = pixel rows of 128 B), which for tf32 needs the
// 128B swizzle with 32B atoms (TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B <-> UMMA layout type 1).
//
// * A band of TH output rows x the FULL padded width BW = Wo+kw-1 is loaded ONCE per 32-channel block as a
//   halo tile (TMA zero-fills the padding and everything outside the image); the kh*kw filter taps are not
//   separate loads but SHIFTED VIEWS of that tile: tap (r,s) starts (r*BW+s) pixel slots = (r*BW+s)*128 B
//   later.  One tcgen05.mma covers 4 consecutive s-taps at once (the 4 M-blocks of 32 channels are 128 B
//   apart: leading-byte-offset = 128 B), so a 3x3 filter needs 3 accumulators per 32-channel block.
// * dz is loaded with the same flattened pitch BW; its columns >= Wo are out of bounds => zero, which is
//   what keeps the wrapped-around x slots of the flattened view from contributing.
// * Each CTA walks a contiguous range of bands accumulating in TMEM (fp32), then dumps its partial dW to
//   a workspace; k_wgrad_finish sums the partials in a fixed order (deterministic), scales, and writes OIHW.
#include "tc_common.cuh"

namespace srb {

namespace {

constexpr int kWgThreads = 192;
constexpr int kMaxAcc = 8;  // accumulators (M=128 tiles) per CTA the unrolled issue loop supports

struct WgArgs {
  int N, Ho, Wo, Ci, Co;
  int kh, kw, pad;
  int TH, BW, BH;       // band rows; flattened pitch (= x box width); x box height = TH + RG - 1
  int TW, bands_w;      // output columns per band and column tiles per image (1: the band spans the full row)
  int dz_rowwise;       // bands_w > 1: dz is loaded one row per TMA (TW slots) at pitch BW; the kw-1 slots between rows stay zero
  int bands_per_img, num_bands, bands_per_cta;
  int CIB, RG, SG, NT;  // ci-blocks per CTA; filter rows per CTA; ceil(kw/4); co tile (multiple of 32)
  int n_cig, n_rg, n_cot;
  int stages, tmem_cols;
  int x_slots, dz_slots;  // allocated 128-B pixel slots per 32-channel block (zero tail included)
  int ksteps;             // ceil(TH*BW / 8)
  int c4;                 // Cin <= 4: x is the zero-padded NHWC4 image seen through an overlapping-stride TMA view whose
                          // 32 "channels" are the 8-pixel x 4-channel window starting at the slot; an accumulator's four
                          // 32-lane M-blocks are four consecutive FILTER ROWS (LBO = one slot row); RG = ceil(kh/4) accumulators
  int bf16;               // operands are bf16: 64-channel blocks (128 B slots), standard 128B swizzle, K = 16 pixels per MMA, and an
                          // accumulator's two 64-lane M-blocks are two consecutive s-taps (c2 == 0) or two ci-blocks (c2 == 1)
  int c2;                 // bf16 only: M = 2 ci-blocks x 1 tap (LBO = one x tile) instead of 1 ci-block x 2 taps (LBO = one slot);
                          // CIB then counts ci-block PAIRS and SG = kw
  int z_ps, z_cpb;        // z_ps > 1: dz is given as PixelShuffle_r(dz) (y's layout, C = Co / r^2 channels): N-block q of the GEMM is
                          // sub-pixel phase q / z_cpb, channel block q % z_cpb, fetched through a stride-r TMA traversal; the
                          // GEMM's output channel co' = ij * C + c is filter co = c * r^2 + ij (k_wgrad_finish undoes it)
  int rn;                 // tf32 generic only: the kh FILTER ROWS are stacked along N instead of being separate accumulators:
                          //   D[(s,ci), (r,co)] = sum_q X[q+s][ci] * dZ[q + (kh-1-r)*BW][co]
                          // q runs over a band of TH INPUT rows (x tile: TH rows, no halo); the dz tile holds the TH+kh-1 output rows
                          // those input rows meet (rows outside the image arrive as zeros), and the B operand's N-block kh-1-r is
                          // that tile viewed kh-1-r rows LATER (LBO = one slot row = BW*128 B) -- one A read feeds kh filter taps.
                          // An accumulator is (32-channel co block, s-group, ci block) and has acc_cols = kh*32 columns.
                          // rn == 2: the co blocks of the tile are stacked along N as well (N = kh * NT <= 256, e.g. 3 x 64 = 192: one
                          // MMA of cost max(32 + N/4, N/2) instead of NT/32 of them): the dz tile is laid out [row][co block][BW slots]
                          // (one TMA per row and co block), so that N-block (kh-1-r) * NT/32 + cob sits (that many) * BW slots after
                          // the view's start -- a single LBO.  K then walks the band row by row (BW % 8 == 0: no K-step straddles
                          // two rows) and B skips the other co blocks' slots at every row end.  Accumulator = (s-group, ci block).
  int acc_cols;           // TMEM columns (= MMA N) per accumulator: NT, kh*32 (rn == 1) or kh*NT (rn == 2)
  int Hi;                 // input rows (rn: the bands tile the INPUT rows)
  int ph_st, ph_a, ph_b, ph_pad0, ph_kh0, ph_kw0;  // ph_st > 0: this launch is one input phase of a strided conv (WgPhase): the
                          // finish kernel scatters its taps into the kh0 x kw0 filter
  int dbg;                // debug knobs (srb_debug_set_flags): 2 = stages are TMA-loaded only once, 4 = no MMAs are issued
  int acc_off[kMaxAcc];   // A-descriptor offset (16-byte units) of accumulator j relative to the stage's x tile (host-computed)
  int acc_boff[kMaxAcc];  // B-descriptor offset (16-byte units) of accumulator j relative to the stage's dz view (rn: its co block)
  float *partial;         // [gridDim.x][gridDim.y][ACC][128][NT]
  float *db_part;         // bias gradient partials [gridDim.x][4 warps][n_cot * NT], or null
};

// MN-major operand, 128B swizzle with 32B atoms: 4 K-rows x 128 B per atom (SBO = 512 B between atoms along K),
// LBO = byte distance between consecutive 32-channel blocks along M/N.
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version
  d |= (uint64_t)1 << 61;  // SWIZZLE_128B_BASE32B
  return d;
}

// tcgen05.mma with the descriptors given as (low, high) word pairs: the running low words are bumped by plain 32-bit adds
// in uniform registers, no 64-bit OR per MMA.
template <bool ACCUM, bool BF = false>
__device__ __forceinline__ void umma_tf32_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
  if (BF)
    asm volatile(
        "{ .reg .pred p; .reg .b64 da, db; setp.ne.b32 p, %5, 0; mov.b64 da, {%1, %3}; mov.b64 db, {%2, %3};\n"
        "  tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p; }" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(ACCUM ? 1u : 0u)
        : "memory");
  else
    asm volatile(
        "{ .reg .pred p; .reg .b64 da, db; setp.ne.b32 p, %5, 0; mov.b64 da, {%1, %3}; mov.b64 db, {%2, %3};\n"
        "  tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p; }" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(ACCUM ? 1u : 0u)
        : "memory");
}

// db partial sums of one bf16 dz stage: NB 64-channel blocks; lane l takes row q0 + l/8 and the 16-B chunk (8 channels) l%8
// (standard 128B swizzle: 16-B chunk index XOR (row & 7))
template <int NB>
__device__ __forceinline__ void db_rows_h(float (&dbs)[4][8], uint32_t sz, int dz_bytes, int rows, int lane_grp, int lrow, int lchunk) {
  for (int q0 = lane_grp * 4; q0 < rows; q0 += 16) {
    const int q = q0 + lrow;
    const uint32_t off = sz + (uint32_t)q * 128u + ((uint32_t)(lchunk ^ (q & 7)) << 4);
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      uint4 v;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(off + (uint32_t)(j * dz_bytes)));
      dbs[j][0] += bf16_lo(v.x); dbs[j][1] += bf16_hi(v.x); dbs[j][2] += bf16_lo(v.y); dbs[j][3] += bf16_hi(v.y);
      dbs[j][4] += bf16_lo(v.z); dbs[j][5] += bf16_hi(v.z); dbs[j][6] += bf16_lo(v.w); dbs[j][7] += bf16_hi(v.w);
    }
  }
}

// db partial sums of one dz stage: NB 32-channel blocks, 4 rows per LDS.128 (see the caller for the lane mapping)
template <int NB>
__device__ __forceinline__ void db_rows(float4 (&dbs)[8], uint32_t sz, int dz_bytes, int rows, int lane_grp, int lrow, int lchunk,
                                        int row_lo = 0) {
  for (int q0 = (row_lo & ~3) + lane_grp * 4; q0 < rows; q0 += 16) {
    const int q = q0 + lrow;
    const uint32_t off = sz + (uint32_t)q * 128u + ((uint32_t)(lchunk ^ ((q & 3) << 1)) << 4);
    const float keep = (q >= row_lo && q < rows) ? 1.f : 0.f;  // rows outside [row_lo, rows) belong to a neighbouring band (or are the zero tail)
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(off + (uint32_t)(j * dz_bytes)));
      dbs[j].x += keep * v.x; dbs[j].y += keep * v.y; dbs[j].z += keep * v.z; dbs[j].w += keep * v.w;
    }
  }
}

// All MMAs of one band (one pipeline stage) for NACC accumulators: straight-line issue code per K-step -- one add and one
// MMA per accumulator.  (The previous runtime-predicated loop spent ~24 instructions per MMA and was issue bound.)
template <int NACC, bool BF, bool ROWS = false>
__device__ __forceinline__ void wg_issue_band(const WgArgs &a, uint32_t x_lo, uint32_t z_lo, uint32_t hi, uint32_t idesc,
                                              uint32_t tmem_base, bool first) {
  if (ROWS) {  // rn == 2: K-steps row by row; at a row end the dz descriptor skips the other co blocks' slots of that row
    const uint32_t ks_row = (uint32_t)a.BW >> 3, b_skip = (uint32_t)((a.NT >> 5) - 1) * (uint32_t)a.BW * 8u;
    uint32_t al[NACC], tc[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) {
      al[j] = x_lo + (uint32_t)a.acc_off[j];
      tc[j] = tmem_base + (uint32_t)(j * a.acc_cols);
    }
    uint32_t bl = z_lo;
    for (int row = 0; row < a.TH; ++row) {
      uint32_t k = 0;
      if (first && row == 0) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
          umma_tf32_lohi<false, false>(tc[j], al[j], bl, hi, idesc);
          al[j] += 64u;
        }
        bl += 64u;
        k = 1;
      }
#pragma unroll 3
      for (; k < ks_row; ++k) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
          umma_tf32_lohi<true, false>(tc[j], al[j], bl, hi, idesc);
          al[j] += 64u;
        }
        bl += 64u;
      }
      bl += b_skip;
    }
    return;
  }
  constexpr uint32_t KADV = BF ? 128u : 64u;  // one K step = 16 (bf16) / 8 (tf32) pixel rows x 128 B, in 16-byte units
  uint32_t al[NACC], tc[NACC], bo[NACC];
#pragma unroll
  for (int j = 0; j < NACC; ++j) {
    al[j] = x_lo + (uint32_t)a.acc_off[j];
    bo[j] = (uint32_t)a.acc_boff[j];
    tc[j] = tmem_base + (uint32_t)(j * a.acc_cols);
  }
  uint32_t bl = z_lo;
  int ks = 0;
  if (first) {  // very first K-step of this CTA: overwrite the accumulators
#pragma unroll
    for (int j = 0; j < NACC; ++j) {
      umma_tf32_lohi<false, BF>(tc[j], al[j], bl + bo[j], hi, idesc);
      al[j] += KADV;
    }
    bl += KADV;
    ks = 1;
  }
#pragma unroll 2
  for (; ks < a.ksteps; ++ks) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) {
      umma_tf32_lohi<true, BF>(tc[j], al[j], bl + bo[j], hi, idesc);
      al[j] += KADV;
    }
    bl += KADV;
  }
}

__global__ void __launch_bounds__(kWgThreads, 1)
k_tc_wgrad(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapZ, WgArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int blk_ch = a.bf16 ? 64 : 32;              // channels per 128-byte slot
  const int nb = (a.NT + blk_ch - 1) / blk_ch;      // dz blocks per stage
  const int XT = a.CIB * (a.c2 ? 2 : 1);            // x tiles (channel blocks) per stage
  const int x_bytes = a.x_slots * 128, dz_bytes = a.dz_slots * 128;
  const int stage_bytes = XT * x_bytes + nb * dz_bytes;
  uint64_t *full_bar = (uint64_t *)(smem + (size_t)a.stages * stage_bytes);
  uint64_t *empty_bar = full_bar + a.stages;
  uint64_t *accum_bar = empty_bar + a.stages;
  uint32_t *tmem_slot = (uint32_t *)(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  // grid.y -> (ci group, filter-row group, co tile)
  int by = blockIdx.y;
  const int cot = by % a.n_cot; by /= a.n_cot;
  const int rgi = by % a.n_rg;
  const int cig = by / a.n_rg;
  const int r0 = rgi * a.RG;
  const int rg_valid = min(a.RG, a.kh - r0);
  const int band0 = blockIdx.x * a.bands_per_cta;
  const int band1 = min(band0 + a.bands_per_cta, a.num_bands);
  const int ACC = a.rn == 2 ? a.SG * a.CIB : (a.rn ? nb * a.SG * a.CIB : a.RG * a.SG * a.CIB);
  // the CTAs of the first (ci group, filter-row group) also reduce dz over pixels: db[co] = sum_p dz[p][co]
  const bool do_db = a.db_part != nullptr && cig == 0 && rgi == 0;

  // zero all stage buffers once: the tails past each TMA box must read as 0.0f forever
  {
    uint4 z = make_uint4(0, 0, 0, 0);
    uint4 *p = (uint4 *)smem;
    const int n16 = a.stages * stage_bytes / 16;
    for (int i = threadIdx.x; i < ((a.dbg & 16) ? 0 : n16); i += kWgThreads) p[i] = z;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapZ) : "memory");
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], do_db ? 5 : 1);  // MMA commit (+ the four column-sum warps)
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
  pdl_wait();  // the shared-memory zeroing and set-up above overlapped the previous kernel; global memory from here on

  if (warp == 0) {
    // ===================== TMA producer: one halo band per stage =====================
    {
      const int z_rows = a.rn ? a.TH + a.kh - 1 : a.TH;  // dz rows per band
      const uint32_t tx_bytes = (uint32_t)(XT * a.BH * a.BW * 128 + nb * z_rows * (a.dz_rowwise ? a.TW : a.BW) * 128);
      int it = 0;
      for (int band = band0; band < band1; ++band, ++it) {
        const int st = it % a.stages;
        const uint32_t ph = (uint32_t)(it / a.stages) & 1u;
        mbar_wait(&empty_bar[st], ph ^ 1u);
        const int n = band / a.bands_per_img;
        const int rem = band - n * a.bands_per_img;
        const int bh = rem / a.bands_w;
        const int oh0 = bh * a.TH, ox0 = (rem - bh * a.bands_w) * a.TW;
        uint8_t *sx = smem + (size_t)st * stage_bytes;
        uint8_t *sz = sx + XT * x_bytes;
        if ((a.dbg & 2) && it >= a.stages) {
          if (elect_one()) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[st])) : "memory");
          __syncwarp();
          continue;
        }
        if (elect_one()) {
          mbar_expect_tx(&full_bar[st], tx_bytes);
          if (a.c4) {
            tma_load_4d(&mapX, &full_bar[st], sx, 0, 0, oh0, n);  // padded image: no negative coordinates
          } else {
            for (int cb = 0; cb < XT; ++cb)
              tma_load_4d(&mapX, &full_bar[st], sx + cb * x_bytes, (cig * XT + cb) * blk_ch, ox0 - a.pad, a.rn ? oh0 : oh0 - a.pad + r0, n);
          }
          for (int j = 0; j < nb; ++j) {
            int zc = cot * a.NT + j * blk_ch, zx = a.dz_rowwise ? ox0 : 0, zy = a.rn ? oh0 + a.pad - (a.kh - 1) : oh0, zr = 1;
            if (a.z_ps > 1) {  // pixel-un-shuffle folded into the load (host guarantees NT % blk_ch == 0)
              const int q = zc / blk_ch, ij = q / a.z_cpb;
              zr = a.z_ps;
              zc = (q - ij * a.z_cpb) * blk_ch;
              zx = zx * zr + ij % zr;
              zy = zy * zr + ij / zr;
            }
            if (a.rn == 2) {
              for (int t = 0; t < z_rows; ++t)  // tile layout [row][co block][BW slots]
                tma_load_4d(&mapZ, &full_bar[st], sz + (size_t)((t * nb + j) * a.BW) * 128, zc, zx, zy + t * zr, n);
            } else if (a.dz_rowwise) {
              for (int t = 0; t < z_rows; ++t)  // rows past the image (and columns past Wo) arrive as zeros
                tma_load_4d(&mapZ, &full_bar[st], sz + j * dz_bytes + t * a.BW * 128, zc, zx, zy + t * zr, n);
            } else {
              tma_load_4d(&mapZ, &full_bar[st], sz + j * dz_bytes, zc, zx, zy, n);
            }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp walks the loop, one elected lane issues) =====================
    {
      // D=f32, A=B=tf32, both MN-major (bits 15,16), N>>3 at bit 17, M=128>>4 at bit 24
      const uint32_t fmt = a.bf16 ? 1u : 2u;  // kind::f16 bf16 operands / kind::tf32
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(a.acc_cols >> 3) << 17) | ((128u >> 4) << 24);
      // descriptor high words are loop invariant; low word = (addr >> 4) | LBO << 16
      // bf16: standard 128B swizzle (layout type 2), 8 K-rows x 128 B per atom -> SBO = 1024 B
      const uint32_t a_hi = a.bf16 ? (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29)) : (uint32_t)(make_mnmajor_desc(0, 128u) >> 32);
      const uint32_t a_lbo = a.c4 ? ((((uint32_t)a.BW * 128u) >> 4) & 0x3FFF) << 16
                                  : (a.c2 ? (((uint32_t)x_bytes >> 4) & 0x3FFF) << 16 : (128u >> 4) << 16);
      // B: consecutive 32-channel N-blocks are the co blocks of the dz stage, or (rn) the same co block one slot row later
      const uint32_t b_lbo = ((((uint32_t)(a.rn ? a.BW * 128 : dz_bytes)) >> 4) & 0x3FFF) << 16;
      int it = 0;
      for (int band = band0; band < band1; ++band, ++it) {
        const int st = it % a.stages;
        const uint32_t ph = (uint32_t)(it / a.stages) & 1u;
        mbar_wait(&full_bar[st], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sx = smem_u32(smem + (size_t)st * stage_bytes);
        const uint32_t x_lo = ((sx >> 4) & 0x3FFF) | a_lbo;
        const uint32_t z_lo = (((sx + XT * x_bytes) >> 4) & 0x3FFF) | b_lbo;
        if (elect_one()) {
          const int nacc = a.rn ? ACC : (a.c4 ? a.RG * a.SG : rg_valid * a.SG * a.CIB);  // a prefix of acc_off (filter-row major)
          const bool first = it == 0;
#define WG_ISSUE(NA)                                                                         \
  if (a.bf16) wg_issue_band<NA, true>(a, x_lo, z_lo, a_hi, idesc, tmem_base, first);       \
  else if (a.rn == 2) wg_issue_band<NA, false, true>(a, x_lo, z_lo, a_hi, idesc, tmem_base, first); \
  else wg_issue_band<NA, false>(a, x_lo, z_lo, a_hi, idesc, tmem_base, first);             \
  break
          switch ((a.dbg & 4) ? 0 : nacc) {
            case 0: break;
            case 1: WG_ISSUE(1);
            case 2: WG_ISSUE(2);
            case 3: WG_ISSUE(3);
            case 4: WG_ISSUE(4);
            case 5: WG_ISSUE(5);
            case 6: WG_ISSUE(6);
            case 7: WG_ISSUE(7);
            default: WG_ISSUE(8);
          }
#undef WG_ISSUE
          umma_commit_arrive(&empty_bar[st]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit_arrive(accum_bar);
      __syncwarp();
    }
  } else {
    // ===================== epilogue: dump the partial dW tile =====================
    const int lane_grp = warp & 3;
    if (do_db && a.bf16) {
      const int rows = a.TH * a.BW;
      float dbs[4][8];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) dbs[j][e] = 0.f;
      const int lrow = lane >> 3, lchunk = lane & 7;
      int it = 0;
      for (int band = band0; band < band1; ++band, ++it) {
        const int st = it % a.stages;
        mbar_wait(&full_bar[st], (uint32_t)(it / a.stages) & 1u);
        const uint32_t sz = smem_u32(smem + (size_t)st * stage_bytes + (size_t)XT * x_bytes);
        if (!(a.dbg & 8)) {
          switch (nb) {
            case 1: db_rows_h<1>(dbs, sz, dz_bytes, rows, lane_grp, lrow, lchunk); break;
            case 2: db_rows_h<2>(dbs, sz, dz_bytes, rows, lane_grp, lrow, lchunk); break;
            case 3: db_rows_h<3>(dbs, sz, dz_bytes, rows, lane_grp, lrow, lchunk); break;
            default: db_rows_h<4>(dbs, sz, dz_bytes, rows, lane_grp, lrow, lchunk); break;
          }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[st])) : "memory");
      }
      float *dp = a.db_part + ((size_t)blockIdx.x * 4 + lane_grp) * (size_t)(a.n_cot * a.NT) + (size_t)cot * a.NT;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < nb) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float v = dbs[j][e];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            const int ch = j * 64 + lchunk * 8 + e;
            if (lane < 8 && ch < a.NT) dp[ch] = v;
          }
        }
      }
    } else if (do_db) {
      // While the MMAs run, these warps walk the same stages and sum the dz tiles over pixels (db).  A warp reads four
      // 128-B rows per LDS.128: lane l takes row q0 + l/8 and the 16-B chunk holding channels 4*(l%8)..+3 of every
      // 32-channel block (128B_ATOM_32B swizzle: 32-byte atom index XOR (row & 3), i.e. 16-B chunk index XOR ((row & 3) << 1)).
      // rn: the dz tiles of neighbouring bands overlap by kh-1 rows; a band sums the TH rows with its own row indices
      const int row_lo = a.rn ? (a.kh - 1 - a.pad) * a.BW : 0;
      const int rows = row_lo + a.TH * a.BW;
      float4 dbs[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) dbs[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int lrow = lane >> 3, lchunk = lane & 7;
      int it = 0;
      for (int band = band0; band < band1; ++band, ++it) {
        const int st = it % a.stages;
        mbar_wait(&full_bar[st], (uint32_t)(it / a.stages) & 1u);
        const uint32_t sz = smem_u32(smem + (size_t)st * stage_bytes + (size_t)XT * x_bytes);
        if (a.rn == 2) {
          // tile [row][co block][BW slots]: the band owns the TH tile rows from kh-1-pad on (BW % 8 == 0: whole 4-slot groups)
          for (int t = a.kh - 1 - a.pad; t < a.kh - 1 - a.pad + a.TH; ++t) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (j < nb) {
                const int q_lo = (t * nb + j) * a.BW;
                for (int q0 = q_lo + lane_grp * 4; q0 < q_lo + a.BW; q0 += 16) {
                  const int q = q0 + lrow;
                  const uint32_t off = sz + (uint32_t)q * 128u + ((uint32_t)(lchunk ^ ((q & 3) << 1)) << 4);
                  float4 v;
                  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(off));
                  dbs[j].x += v.x; dbs[j].y += v.y; dbs[j].z += v.z; dbs[j].w += v.w;
                }
              }
            }
          }
        } else if (!(a.dbg & 8)) {
          // rows .. round_up(rows, 8) lie in the tile's zero tail (dz_slots % 8 == 0); warp w takes rows 4w.., 4w+16..
          switch (nb) {
            case 1: db_rows<1>(dbs, sz, dz_bytes, rows, lane_grp, lrow, lchunk, row_lo); break;
            case 2: db_rows<2>(dbs, sz, dz_bytes, rows, lane_grp, lrow, lchunk, row_lo); break;
            case 3: db_rows<3>(dbs, sz, dz_bytes, rows, lane_grp, lrow, lchunk, row_lo); break;
            case 4: db_rows<4>(dbs, sz, dz_bytes, rows, lane_grp, lrow, lchunk, row_lo); break;
            case 5: db_rows<5>(dbs, sz, dz_bytes, rows, lane_grp, lrow, lchunk, row_lo); break;
            case 6: db_rows<6>(dbs, sz, dz_bytes, rows, lane_grp, lrow, lchunk, row_lo); break;
            case 7: db_rows<7>(dbs, sz, dz_bytes, rows, lane_grp, lrow, lchunk, row_lo); break;
            default: db_rows<8>(dbs, sz, dz_bytes, rows, lane_grp, lrow, lchunk, row_lo); break;
          }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[st])) : "memory");
      }
      // fold the four row-lanes (lanes l, l^8, l^16 hold the same channels), then lanes 0..7 store 4 channels each
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < nb) {
          float4 v = dbs[j];
#pragma unroll
          for (int o = 8; o <= 16; o <<= 1) {
            v.x += __shfl_xor_sync(0xffffffffu, v.x, o); v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
            v.z += __shfl_xor_sync(0xffffffffu, v.z, o); v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
          }
          dbs[j] = v;
        }
      }
      float *dp = a.db_part + ((size_t)blockIdx.x * 4 + lane_grp) * (size_t)(a.n_cot * a.NT) + (size_t)cot * a.NT;
      if (lane < 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < nb) *(float4 *)(dp + j * 32 + lane * 4) = dbs[j];
      }
    }
    mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int m = lane_grp * 32 + lane;
    float *dst = a.partial + ((size_t)blockIdx.x * gridDim.y + blockIdx.y) * ACC * 128 * a.acc_cols;
    // The partial tile goes TMEM -> registers -> shared memory (the operand stages are idle: every MMA has completed) and
    // leaves as a few asynchronous BULK stores (cp.async.bulk shared -> global), instead of 4-byte-granular per-thread
    // stores that the kernel's exit then waits for.  Falls back to direct stores when the tile does not fit the stages.
    const size_t dump_bytes = (size_t)ACC * 128 * a.acc_cols * sizeof(float);
    const bool via_smem = dump_bytes <= (size_t)a.stages * stage_bytes && !(a.dbg & 128);
    float *sdump = (float *)smem;
    // the other epilogue warps may still be summing the last band's dz tile out of these very stages (db_rows)
    if (via_smem) asm volatile("bar.sync 1, %0;" ::"r"(kWgThreads - 64) : "memory");
    for (int acc = 0; acc < ((a.dbg & 32) ? 0 : ACC); ++acc) {
      const int rl = a.rn ? 0 : acc / (a.SG * a.CIB);
      const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * a.acc_cols);
      float *row = (via_smem ? sdump : dst) + ((size_t)acc * 128 + m) * a.acc_cols;
      for (int j0 = 0; j0 < a.acc_cols; j0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)j0, v);
        if (rl >= rg_valid || band0 >= band1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *(uint4 *)(row + j0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
    if (via_smem && !(a.dbg & 32)) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the bulk-copy engine
      asm volatile("bar.sync 1, %0;" ::"r"(kWgThreads - 64) : "memory");  // the four epilogue warps
      if (warp == 2 && elect_one()) {
        const uint32_t sbase = smem_u32(sdump);
        for (size_t off = 0; off < dump_bytes; off += 32768) {
          const uint32_t nbytes = (uint32_t)(dump_bytes - off < 32768 ? dump_bytes - off : 32768);
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"((const char *)dst + off), "r"(sbase + (uint32_t)off),
                       "r"(nbytes)
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the shared memory must outlive the reads (and the writes this kernel's exit publishes)
      }
      __syncwarp();
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

// Where GEMM column co, filter tap (ci, r, s) lives in a CTA's partial dump: grid.y index, accumulator, TMEM lane m, column n.
__device__ __forceinline__ void wg_locate(const WgArgs &a, int co, int ci, int r, int s, int &by, int &acc, int &m, int &n) {
  const int cot = co / a.NT;
  n = co - cot * a.NT;
  if (a.c4) {
    by = cot;
    acc = (r >> 2) * a.SG + (s >> 3);
    m = (r & 3) * 32 + (s & 7) * 4 + ci;
  } else if (a.bf16) {
    const int cblk = ci >> 6, rgi = r / a.RG, rl = r - rgi * a.RG;
    int cig, cb, sg, half;
    if (a.c2) { const int cp = cblk >> 1; cig = cp / a.CIB; cb = cp - cig * a.CIB; sg = s; half = cblk & 1; }
    else      { cig = cblk / a.CIB; cb = cblk - cig * a.CIB; sg = s >> 1; half = s & 1; }
    by = (cig * a.n_rg + rgi) * a.n_cot + cot;
    acc = (rl * a.SG + sg) * a.CIB + cb;
    m = half * 64 + (ci & 63);
  } else if (a.rn == 2) {
    const int cblk = ci >> 5, cig = cblk / a.CIB, cb = cblk - cig * a.CIB;
    const int sg = s >> 2, sl = s & 3, cob = n >> 5;
    by = cig * a.n_cot + cot;
    acc = sg * a.CIB + cb;
    m = sl * 32 + (ci & 31);
    n = ((a.kh - 1 - r) * (a.NT >> 5) + cob) * 32 + (n & 31);  // N-block (kh-1-r) * NT/32 + cob
  } else if (a.rn) {
    const int cblk = ci >> 5, cig = cblk / a.CIB, cb = cblk - cig * a.CIB;
    const int sg = s >> 2, sl = s & 3, cob = n >> 5;
    by = cig * a.n_cot + cot;
    acc = (cob * a.SG + sg) * a.CIB + cb;
    m = sl * 32 + (ci & 31);
    n = (a.kh - 1 - r) * 32 + (n & 31);  // N-block kh-1-r of the accumulator holds filter row r
  } else {
    const int cblk = ci >> 5, cig = cblk / a.CIB, cb = cblk - cig * a.CIB;
    const int rgi = r / a.RG, rl = r - rgi * a.RG;
    const int sg = s >> 2, sl = s & 3;
    by = (cig * a.n_rg + rgi) * a.n_cot + cot;
    acc = (rl * a.SG + sg) * a.CIB + cb;
    m = sl * 32 + (ci & 31);
  }
}

// GEMM column co' -> filter index (identity unless dz was un-shuffled on the fly: co' = ij * C + c  ->  c * r^2 + ij)
__device__ __forceinline__ int wg_filter_of(const WgArgs &a, int co) {
  if (a.z_ps <= 1) return co;
  const int rr = a.z_ps * a.z_ps, C = a.Co / rr, ij = co / C;
  return (co - ij * C) * rr + ij;
}

// dW[co][ci][r][s] (=|+=) scale * sum_over_splits partial[...]      (fixed summation order => deterministic)
// A work item is one (ci, tap) pair of a tile of 32 GEMM columns: a 128-byte partial row per split.  One block owns IPB
// consecutive items j = ci * taps + tap, i.e. a contiguous run of IPB floats in each of its 32 filters.
//   lanes8 == 0 (few splits):  the 8 warps take different items, each thread walks the splits with 8 independent accumulators
//   lanes8 == 1 (many splits): the 8 warps share every item, warp y sums splits y, y+8, ...; folded in a fixed order in smem
// The sums are staged in shared memory so that the OIHW stores are contiguous runs.  (The previous one-thread-per-output
// kernel spent 47 us on EDSR-256's 28 MB of partials: 4-byte stores 9 KB apart and 3 loads in flight per thread.)
__global__ void __launch_bounds__(256) k_wgrad_finish(WgArgs a, int splits, int gy, int IPB, int lanes8, float *dw, float *db,
                                                      float scale, int accumulate) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float tile[];  // [32][IPB + 1] (+ [8][33] fold buffer when lanes8)
  const int taps = a.kh * a.kw, total_items = a.Ci * taps, pitch = IPB + 1;
  const int co_tiles = (a.Co + 31) / 32, groups = (total_items + IPB - 1) / IPB;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ACC = a.rn == 2 ? a.SG * a.CIB : (a.rn ? (a.NT / 32) * a.SG * a.CIB : a.RG * a.SG * a.CIB);
  const size_t ss = (size_t)gy * ACC * 128 * a.acc_cols;  // floats between consecutive splits
  if ((int)blockIdx.x < co_tiles * groups) {
    const int cot32 = blockIdx.x % co_tiles, j0 = (blockIdx.x / co_tiles) * IPB;
    const int co = cot32 * 32 + tx;
    const int nitems = min(IPB, total_items - j0);
    float *fold = tile + 32 * pitch;
    for (int jj = lanes8 ? 0 : ty; jj < nitems; jj += lanes8 ? 1 : 8) {
      const int j = j0 + jj, ci = j / taps, tap = j - ci * taps, r = tap / a.kw, s = tap - r * a.kw;
      float sum = 0.f;
      if (co < a.Co) {
        int by, acc, m, n;
        wg_locate(a, co, ci, r, s, by, acc, m, n);
        const float *p = a.partial + (((size_t)by * ACC + acc) * 128 + m) * a.acc_cols + n;
        if (lanes8) {
          // 8 independent loads in flight per thread: a 148-way split is 2-3 round trips to L2 instead of 5
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f, s5 = 0.f, s6 = 0.f, s7 = 0.f;
          int z = ty;
          for (; z + 56 < splits; z += 64) {
            s0 += p[(size_t)z * ss]; s1 += p[(size_t)(z + 8) * ss]; s2 += p[(size_t)(z + 16) * ss]; s3 += p[(size_t)(z + 24) * ss];
            s4 += p[(size_t)(z + 32) * ss]; s5 += p[(size_t)(z + 40) * ss]; s6 += p[(size_t)(z + 48) * ss]; s7 += p[(size_t)(z + 56) * ss];
          }
          if (z < splits) s0 += p[(size_t)z * ss];
          if (z + 8 < splits) s1 += p[(size_t)(z + 8) * ss];
          if (z + 16 < splits) s2 += p[(size_t)(z + 16) * ss];
          if (z + 24 < splits) s3 += p[(size_t)(z + 24) * ss];
          if (z + 32 < splits) s4 += p[(size_t)(z + 32) * ss];
          if (z + 40 < splits) s5 += p[(size_t)(z + 40) * ss];
          if (z + 48 < splits) s6 += p[(size_t)(z + 48) * ss];
          sum = ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
        } else {
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f, s5 = 0.f, s6 = 0.f, s7 = 0.f;
          int z = 0;
          for (; z + 7 < splits; z += 8) {
            s0 += p[(size_t)z * ss]; s1 += p[(size_t)(z + 1) * ss]; s2 += p[(size_t)(z + 2) * ss]; s3 += p[(size_t)(z + 3) * ss];
            s4 += p[(size_t)(z + 4) * ss]; s5 += p[(size_t)(z + 5) * ss]; s6 += p[(size_t)(z + 6) * ss]; s7 += p[(size_t)(z + 7) * ss];
          }
          for (; z < splits; ++z) s0 += p[(size_t)z * ss];
          sum = ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
        }
      }
      if (lanes8) {
        fold[ty * 33 + tx] = sum;
        __syncthreads();
        if (ty == 0) {
          float t = fold[tx];
#pragma unroll
          for (int y = 1; y < 8; ++y) t += fold[y * 33 + tx];
          tile[tx * pitch + jj] = t;
        }
        __syncthreads();
      } else {
        tile[tx * pitch + jj] = sum;
      }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * nitems; e += 256) {
      const int col = e / nitems, jj = e - col * nitems;
      const int co2 = cot32 * 32 + col;
      if (co2 < a.Co) {
        float *d;
        if (a.ph_st > 0) {  // phase launch of a strided conv: launch tap (tr, ts) -> filter tap (r, s), or nothing
          const int j = j0 + jj, ci = j / taps, tap = j - ci * taps, tr = tap / a.kw, ts = tap - tr * a.kw;
          const int r = a.ph_st * (tr - a.pad) + a.ph_a + a.ph_pad0, s2 = a.ph_st * (ts - a.pad) + a.ph_b + a.ph_pad0;
          if (r < 0 || r >= a.ph_kh0 || s2 < 0 || s2 >= a.ph_kw0) continue;
          d = dw + (((size_t)co2 * a.Ci + ci) * a.ph_kh0 + r) * a.ph_kw0 + s2;
        } else {
          d = dw + (size_t)wg_filter_of(a, co2) * total_items + j0 + jj;
        }
        const float t = tile[col * pitch + jj] * scale;
        *d = accumulate ? *d + t : t;
      }
    }
  } else if (db != nullptr) {
    // bias gradient: splits x 4 warp partials per channel; 32 channels per block, 8 z-lanes folded through shared memory
    const int co = ((int)blockIdx.x - co_tiles * groups) * 32 + tx;
    const size_t pitch_db = (size_t)(a.n_cot * a.NT);
    float sum = 0.f;
    if (co < a.Co)
      for (int z = ty; z < splits * 4; z += 8) sum += a.db_part[(size_t)z * pitch_db + co];
    tile[ty * 33 + tx] = sum;
    __syncthreads();
    if (ty == 0 && co < a.Co) {
      float t = tile[tx];
#pragma unroll
      for (int y = 1; y < 8; ++y) t += tile[y * 33 + tx];
      t *= scale;
      float *d = db + wg_filter_of(a, co);
      *d = accumulate ? *d + t : t;
    }
  }
}

// x (N, C<=4, H, W; any strides) -> zero-padded NHWC4 image xp[N][H+2p][Wp][4] (Wp >= W+2p+8*SG), tf32-rounded.  The
// padding and the extra right-hand columns are real zeros so that every 8-pixel window of the overlapping view is finite.
__global__ void k_pack_nhwc4_padded(T4 x, float4 *__restrict__ xp, int N, int C, int H, int W, int pad, int Hp, int Wp,
                                    long long total) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % Wp);
    long long q = i / Wp;
    const int yy = (int)(q % Hp);
    const long long n = q / Hp;
    const int h = yy - pad, w_ = xx - pad;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (n < N && h >= 0 && h < H && w_ >= 0 && w_ < W) {
      const float *p = x.p + n * x.sn + (long long)h * x.sh + (long long)w_ * x.sw;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < C) v[c] = round_tf32(__ldg(p + c * x.sc));
    }
    xp[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

struct WgPlan {
  WgArgs a;
  size_t smem;
  dim3 grid;
  int db_blocks;
  size_t partial_floats, db_floats;
  int Hp, Wp;            // c4: padded NHWC4 image geometry
  size_t xpack_floats;   // c4: floats of the packed image (incl. the tail the last windows run into)
};

// Cin <= 4 (network inputs): see WgArgs::c4.
bool make_wg_plan_c4(const Geom &g, WgPlan *pl) {
  WgArgs &a = pl->a;
  a.z_ps = 1; a.z_cpb = 0;
  a.N = g.N; a.Ho = g.Ho; a.Wo = g.Wo; a.Ci = g.Ci; a.Co = g.Co;
  a.kh = g.kh; a.kw = g.kw; a.pad = g.pad;
  a.c4 = 1; a.bf16 = 0; a.c2 = 0; a.rn = 0; a.Hi = g.Hi;
  a.TW = g.Wo; a.bands_w = 1; a.dz_rowwise = 0;
  a.BW = g.Wo + g.kw - 1;  // == Wi + 2*pad at stride 1
  if (a.BW > 256 || g.kw > 16 || g.kh > 16) return false;
  a.SG = (g.kw + 7) / 8;
  a.RG = (g.kh + 3) / 4;
  a.CIB = 1;
  const int nacc = a.RG * a.SG;
  if (nacc > kMaxAcc) return false;
  const int co_pad = round_up_i(g.Co, 32);
  int NT = co_pad < 256 ? co_pad : 256;
  while (NT > 32 && (co_pad % NT || nacc * NT > 512)) NT -= 32;
  if (co_pad % NT || nacc * NT > 512) return false;
  int bestTH = 0, best_stages = 0;
  size_t best_stage = 0;
  int best_xs = 0, best_zs = 0;
  for (int TH = 8; TH >= 1; --TH) {
    if (TH > g.Ho && TH > 1) continue;
    const int x_slots = round_up_i((TH + 4 * a.RG - 1) * a.BW + 8 * a.SG + 8, 8);
    const int dz_slots = round_up_i(TH * a.BW, 8);
    const size_t stage = (size_t)x_slots * 128 + (size_t)(NT / 32) * dz_slots * 128;
    int stages = (int)((kMaxSmemBytes - 4096) / stage);
    if (stages < 2 && TH > 1) continue;  // a single-stage pipeline only as the last resort (k9 windows at TH = 1)
    if (stages < 1) continue;
    if (stages > 4) stages = 4;
    bestTH = TH; best_stages = stages; best_stage = stage; best_xs = x_slots; best_zs = dz_slots;
    break;  // largest band that still double-buffers: least halo re-read
  }
  if (!bestTH) return false;
  a.NT = NT; a.acc_cols = NT; a.TH = bestTH; a.BH = bestTH + g.kh - 1;
  a.x_slots = best_xs; a.dz_slots = best_zs; a.stages = best_stages;
  a.n_cig = 1; a.n_rg = 1; a.n_cot = co_pad / NT;
  a.bands_per_img = (g.Ho + a.TH - 1) / a.TH;
  a.num_bands = g.N * a.bands_per_img;
  a.ksteps = (a.TH * a.BW + 7) / 8;
  const int gy = a.n_cot;
  int target = 148 / gy;
  if (target < 1) target = 1;
  a.bands_per_cta = (a.num_bands + target - 1) / target;
  if (a.bands_per_cta < 1) a.bands_per_cta = 1;  // empty batch: plan for the workspace query only
  const int gx = (a.num_bands + a.bands_per_cta - 1) / a.bands_per_cta;
  int tc = 32;
  while (tc < nacc * NT) tc <<= 1;
  a.tmem_cols = tc;
  pl->grid = dim3(gx, gy);
  pl->smem = (size_t)best_stages * best_stage + 1024 + (2 * best_stages + 1) * 8 + 16;
  pl->partial_floats = (size_t)gx * gy * nacc * 128 * NT;
  pl->db_blocks = 0;
  pl->db_floats = (size_t)gx * 4 * a.n_cot * NT;
  pl->Hp = g.Hi + 2 * g.pad;
  pl->Wp = g.Wi + 2 * g.pad + 8 * a.SG;
  pl->xpack_floats = ((size_t)g.N * pl->Hp * pl->Wp + 64) * 4;
  return true;
}

// bf16 operands: 64-channel blocks, K = 16 pixels per MMA, M = 128 as two 64-lane blocks (WgArgs::bf16 / c2).
bool make_wg_plan_h(const Geom &g, WgPlan *pl, int z_ps) {
  WgArgs &a = pl->a;
  a.c4 = 0; a.bf16 = 1; a.rn = 0; a.Hi = g.Hi;
  a.z_ps = z_ps; a.z_cpb = z_ps > 1 ? g.Co / (z_ps * z_ps) / 64 : 0;
  pl->Hp = pl->Wp = 0;
  pl->xpack_floats = 0;
  a.N = g.N; a.Ho = g.Ho; a.Wo = g.Wo; a.Ci = g.Ci; a.Co = g.Co;
  a.kh = g.kh; a.kw = g.kw; a.pad = g.pad;
  const int cblocks = (g.Ci + 63) / 64;
  const int co_pad = round_up_i(g.Co, 16);
  double best_score = -1.0;
  WgArgs best = a;
  size_t best_smem = 0;
  for (int c2 = 0; c2 <= 1; ++c2) {
    if (c2 && (cblocks & 1)) continue;
    const int units = c2 ? cblocks / 2 : cblocks;      // ci units (blocks or pairs) to distribute
    const int SG = c2 ? g.kw : (g.kw + 1) / 2;
    const double tap_eff = c2 ? 1.0 : (double)g.kw / (2.0 * SG);
    for (int wsplit = 1; wsplit <= 16; ++wsplit) {
      const int TW = (g.Wo + wsplit - 1) / wsplit;
      if (wsplit > 1 && (TW < 16 || (g.Wo + TW - 1) / TW != wsplit)) continue;
      const int bands_w = (g.Wo + TW - 1) / TW;
      // row-wise dz loads land at t * BW * 128 B: keep every row on the 1024-B period of the 128B swizzle
      const int BW = bands_w > 1 ? round_up_i(TW + g.kw - 1, 8) : g.Wo + g.kw - 1;
      if (BW * z_ps > 256) continue;
      for (int NT = (co_pad < 256 ? co_pad : 256); NT >= 16; NT -= 16) {
        if (co_pad % NT) continue;
        if (z_ps > 1 && NT % 64) continue;  // whole 64-channel blocks per sub-pixel phase
        const int nb = (NT + 63) / 64;
        for (int CIB = units; CIB >= 1; --CIB) {
          if (units % CIB) continue;
          const int XT = CIB * (c2 ? 2 : 1);
          for (int RG = g.kh; RG >= 1; --RG) {
            const int acc = RG * SG * CIB;
            if (acc * NT > 512 || acc > kMaxAcc) continue;
            for (int TH = 16; TH >= 1; --TH) {
              if (TH > g.Ho && TH > 1) continue;
              const int BH = TH + RG - 1;
              if (BH > 256 || TH * z_ps > 256) continue;
              const int x_slots = round_up_i(BH * BW + g.kw + 16, 16);
              const int dz_slots = round_up_i(TH * BW, 16);
              const size_t stage = (size_t)XT * x_slots * 128 + (size_t)nb * dz_slots * 128;
              int stages = (int)((kMaxSmemBytes - 4096) / stage);
              if (stages < 2) continue;
              if (stages > 4) stages = 4;
              const int n_rg = (g.kh + RG - 1) / RG, n_cig = units / CIB, n_cot = co_pad / NT;
              // per image row, summed over grid.y: MMA cycles (K = 16 slots per MMA) vs L2->SM bytes
              const double costN = (32.0 + NT / 4.0) > NT / 2.0 ? (32.0 + NT / 4.0) : NT / 2.0;
              const double mma = (double)n_cig * n_rg * n_cot * acc * costN * (bands_w * BW / 16.0);
              const double bytes = ((double)n_cot * n_rg * cblocks * BH / TH * BW + (double)n_cig * n_rg * n_cot * nb * TW) * 128.0 * bands_w;
              // TMA writes share the SMEM port with the MMA operand reads (4 KB + NT*32 B per MMA)
              const double mma_c = mma * (1.0 + bytes / ((double)n_cig * n_rg * n_cot * acc * (bands_w * BW / 16.0) * (4096.0 + NT * 32.0)));
              double t = mma_c > bytes / 36.0 ? mma_c : bytes / 36.0;
              t += 500.0 * bands_w / TH * (n_cig * n_rg * n_cot);
              const double n_tma = XT + (bands_w > 1 ? (double)nb * TH : (double)nb);
              t += 40.0 * n_tma * bands_w / TH * (n_cig * n_rg * n_cot);
              if (stages == 2) t *= 1.08;
              (void)tap_eff;  // already in `mma`: the wasted half-accumulator of an odd tap count is issued like a useful one
              const double score = 1e9 / t + TH * 1e-3;
              if (score > best_score) {
                best_score = score;
                best = a;
                best.c2 = c2; best.SG = SG;
                best.TW = TW; best.bands_w = bands_w; best.BW = BW; best.dz_rowwise = bands_w > 1 ? 1 : 0;
                best.NT = NT; best.CIB = CIB; best.RG = RG; best.TH = TH; best.BH = BH;
                best.x_slots = x_slots; best.dz_slots = dz_slots; best.stages = stages;
                best.n_cig = n_cig; best.n_rg = n_rg; best.n_cot = n_cot;
                best_smem = (size_t)stages * stage + 1024 + (2 * stages + 1) * 8 + 16;
              }
            }
          }
        }
      }
    }
  }
  if (best_score < 0) return false;
  a = best;
  a.acc_cols = a.NT;
  a.bands_per_img = ((g.Ho + a.TH - 1) / a.TH) * a.bands_w;
  a.num_bands = g.N * a.bands_per_img;
  a.ksteps = (a.TH * a.BW + 15) / 16;
  const int gy = a.n_cig * a.n_rg * a.n_cot;
  int target = 148 / gy;
  if (target < 1) target = 1;
  a.bands_per_cta = (a.num_bands + target - 1) / target;
  if (a.bands_per_cta < 1) a.bands_per_cta = 1;
  const int gx = (a.num_bands + a.bands_per_cta - 1) / a.bands_per_cta;
  int cols = a.RG * a.SG * a.CIB * a.NT, tc = 32;
  while (tc < cols) tc <<= 1;
  a.tmem_cols = tc;
  pl->grid = dim3(gx, gy);
  pl->smem = best_smem;
  pl->partial_floats = (size_t)gx * gy * a.RG * a.SG * a.CIB * 128 * a.NT;
  pl->db_blocks = 0;
  pl->db_floats = (size_t)gx * 4 * a.n_cot * a.NT;
  return true;
}

bool make_wg_plan(const Geom &g, WgPlan *pl, bool bf16 = false, int z_ps = 1) {
  if (bf16) return make_wg_plan_h(g, pl, z_ps);
  if (g.Ci <= 4) return z_ps == 1 && make_wg_plan_c4(g, pl);
  WgArgs &a = pl->a;
  a.c4 = 0; a.bf16 = 0; a.c2 = 0; a.rn = 0; a.Hi = g.Hi;
  a.z_ps = z_ps; a.z_cpb = z_ps > 1 ? g.Co / (z_ps * z_ps) / 32 : 0;
  pl->Hp = pl->Wp = 0;
  pl->xpack_floats = 0;
  a.N = g.N; a.Ho = g.Ho; a.Wo = g.Wo; a.Ci = g.Ci; a.Co = g.Co;
  a.kh = g.kh; a.kw = g.kw; a.pad = g.pad;
  a.SG = (g.kw + 3) / 4;
  const int cblocks = g.Ci / 32;
  const int co_pad = round_up_i(g.Co, 32);
  double best_score = -1.0;
  WgArgs best = a;
  size_t best_smem = 0;
  // Enumerate (row stacking, column split, NT, CIB, RG, TH) and minimise a time model per image row, summed over the CTAs of grid.y:
  //   MMA cycles  = accumulators * cost(N) * (pitch / 8 K-steps),   cost(N) = max(32 + N/4, N/2) + 15  [measured, tools/bench_umma.cu:
  //                 an SS-mode tf32 MMA re-reads its 4 KB A tile and N*32 B of B from shared memory at 128 B/cycle; ~15 cycles
  //                 of every MMA's issue are not hidden behind the previous one (in-kernel traces, tools/trace_rs.py)]
  //   load cycles = TMA bytes (halo rows and halo columns included) / (L2->SM share of one SM ~ 23 B/cycle)
  //   + a fixed hand-off cost per band.
  // Row stacking (WgArgs::rn): N = kh * 32 per MMA and one accumulator per (co block, s-group, ci block); K runs over the
  // TH + kh - 1 rows of the x tile, so small bands pay (TH + kh - 1) / TH more K-steps.
  // Column split w: the band covers TW = ceil(Wo / w) output columns; x is fetched as a (TW + kw - 1)-wide halo box and dz
  // row by row at the same pitch.  Wide images (W >= 128) would otherwise be limited to TH = 1 (3x halo re-read).
  const int dbg_flags = tc_conv_get_dbg();
  // experiments only: SRB_WG_FORCE="TH,wsplit" pins the band shape of the generic tf32 planner (0 = free)
  static const int force_th = [] { const char *e = getenv("SRB_WG_FORCE"); return e ? atoi(e) : 0; }();
  static const int force_ws = [] { const char *e = getenv("SRB_WG_FORCE"); const char *c = e ? strchr(e, ',') : nullptr; return c ? atoi(c + 1) : 0; }();
  const bool force2 = (dbg_flags & 524288) != 0;  // tests: take the rows+co-stacked flavour wherever it has a plan
  for (int rni = 0; rni <= 2; ++rni) {
  const int rn = force2 ? 2 - rni : rni;
  if (force2 && rn != 2 && best_score >= 0) continue;
  // (a band owns the output rows with its own input-row indices: needs Ho <= Hi, i.e. 2 * pad <= kh - 1)
  if (rn && (z_ps != 1 || g.kh < 2 || g.kh * 32 > 256 || 2 * g.pad > g.kh - 1 || (dbg_flags & 256))) continue;
  if (!rn && (dbg_flags & 512)) continue;
  if (rn == 2 && (dbg_flags & 262144)) continue;
  for (int wsplit = 1; wsplit <= 16; ++wsplit) {
    const int TW = (g.Wo + wsplit - 1) / wsplit;
    if (wsplit > 1 && (TW < 16 || (g.Wo + TW - 1) / TW != wsplit)) continue;
    if (force_ws && wsplit != force_ws) continue;
    const int bands_w = (g.Wo + TW - 1) / TW;
    // row-wise dz loads land at t * BW * 128 B: keep every row on the 512-B period of the 32B-atom swizzle
    // rn == 2: K walks row by row, so a row is a whole number of 8-slot K-steps
    const int BW = rn == 2 ? round_up_i(TW + g.kw - 1, 8) : (bands_w > 1 ? round_up_i(TW + g.kw - 1, 4) : g.Wo + g.kw - 1);
    if (BW * z_ps > 256) continue;
    for (int NT = (co_pad < 256 ? co_pad : 256); NT >= 32; NT -= 32) {
      if (co_pad % NT) continue;
      if (rn == 2 && (NT < 64 || NT * g.kh > 256)) continue;  // NT == 32 is rn == 1; the MMA's N is at most 256
      for (int CIB = cblocks; CIB >= 1; --CIB) {
        if (cblocks % CIB) continue;
        for (int RG = g.kh; RG >= (rn ? g.kh : 1); --RG) {
          const int acc = rn == 2 ? a.SG * CIB : (rn ? (NT / 32) * a.SG * CIB : RG * a.SG * CIB);
          const int acc_cols = rn == 2 ? g.kh * NT : (rn ? g.kh * 32 : NT);
          if (acc * acc_cols > 512 || acc > kMaxAcc) continue;
          for (int TH = 16; TH >= 1; --TH) {
            if (TH > g.Ho && TH > 1) continue;
            if (rn && TH > g.Hi && TH > 1) continue;
            if (force_th && TH != force_th) continue;
            int BH = rn ? TH : TH + RG - 1;  // x box rows: rn bands are TH input rows, the halo is on the dz side
            if (BH > 256 || TH * z_ps > 256 || TH + g.kh - 1 > 256) continue;
            int x_slots = round_up_i(BH * BW + 4 * a.SG + 8, 8);
            // rn: TH + kh - 1 rows of dz, + the slots the last N-block's view reaches past them in the rounded-up last K-step
            int dz_slots = rn ? round_up_i((TH + g.kh - 1) * BW + 8, 8) : round_up_i(TH * BW, 8);
            size_t stage = (size_t)CIB * x_slots * 128 + (size_t)(NT / 32) * dz_slots * 128;
            int stages = (int)((kMaxSmemBytes - 4096) / stage);
            if (stages < 2) continue;
            if (stages > 4) stages = 4;
            int n_rg = (g.kh + RG - 1) / RG, n_cig = cblocks / CIB, n_cot = co_pad / NT;
            double costN = ((32.0 + acc_cols / 4.0) > acc_cols / 2.0 ? (32.0 + acc_cols / 4.0) : acc_cols / 2.0) + 15.0;
            double mma = (double)n_cig * n_rg * n_cot * acc * costN * (bands_w * BW / 8.0);
            const double zrows = rn ? (double)(TH + g.kh - 1) / TH : 1.0;  // dz rows loaded per band row
            double bytes = ((double)n_cot * n_rg * cblocks * BH / TH * BW + (double)n_cig * n_rg * (co_pad / 32) * TW * zrows) * 128.0 * bands_w;
            double t = mma > bytes / 23.0 ? mma : bytes / 23.0;
            t += 500.0 * bands_w / TH * (n_cig * n_rg * n_cot);  // per-band hand-off (barrier round trips, TMA issue)
            // every TMA instruction costs issue slots and a request round trip; row-wise dz loads issue NT/32 * TH small
            // boxes per band (edsr256: 32 x 2 KB), which measurably starves the pipeline
            const double n_tma = CIB + (rn == 2 ? (double)(NT / 32) * (TH + g.kh - 1)
                                                : (bands_w > 1 ? (double)(NT / 32) * TH : (double)(NT / 32)));
            t += 40.0 * n_tma * bands_w / TH * (n_cig * n_rg * n_cot);
            if (stages == 2) t *= 1.08;                            // less slack for the TMA latency
            double score = 1e9 / t + TH * 1e-3;
            if (score > best_score) {
              best_score = score;
              best = a;
              best.rn = rn; best.acc_cols = acc_cols;
              best.TW = TW; best.bands_w = bands_w; best.BW = BW; best.dz_rowwise = (bands_w > 1 || rn == 2) ? 1 : 0;
              best.NT = NT; best.CIB = CIB; best.RG = RG; best.TH = TH; best.BH = BH;
              best.x_slots = x_slots; best.dz_slots = dz_slots; best.stages = stages;
              best.n_cig = n_cig; best.n_rg = n_rg; best.n_cot = n_cot;
              best_smem = (size_t)stages * stage + 1024 + (2 * stages + 1) * 8 + 16;
            }
          }
        }
      }
    }
  }
  }  // rn
  if (best_score < 0) return false;
  a = best;
  a.bands_per_img = (((a.rn ? g.Hi : g.Ho) + a.TH - 1) / a.TH) * a.bands_w;
  a.num_bands = g.N * a.bands_per_img;
  a.ksteps = (a.TH * a.BW + 7) / 8;
  int gy = a.n_cig * a.n_rg * a.n_cot;
  int target = 148 / gy;  // at most one wave of CTAs (1 CTA/SM: big smem)
  if (target < 1) target = 1;
  a.bands_per_cta = (a.num_bands + target - 1) / target;
  if (a.bands_per_cta < 1) a.bands_per_cta = 1;  // empty batch: plan for the workspace query only
  int gx = (a.num_bands + a.bands_per_cta - 1) / a.bands_per_cta;
  const int nacc = a.rn == 2 ? a.SG * a.CIB : (a.rn ? (a.NT / 32) * a.SG * a.CIB : a.RG * a.SG * a.CIB);
  int cols = nacc * a.acc_cols, tc = 32;
  while (tc < cols) tc <<= 1;
  a.tmem_cols = tc;
  pl->grid = dim3(gx, gy);
  pl->smem = best_smem;
  pl->partial_floats = (size_t)gx * gy * nacc * 128 * a.acc_cols;
  pl->db_blocks = 0;
  pl->db_floats = (size_t)gx * 4 * a.n_cot * a.NT;
  return true;
}

}  // namespace

// small = dz (N,Co,Ho,Wo) NHWC, big = x (N,Ci,Hi,Wi) NHWC
bool tc_wgrad_supported(const Geom &g, const T4 &small, const T4 &big, int z_ps) {
  if (g.st != 1 || g.ps != 1 || g.N <= 0) return false;
  if (small.dt != big.dt) return false;  // mixed operand types: the caller converts one side first
  if (z_ps > 1) {  // `small` is PixelShuffle_r(dz): whole channel blocks per sub-pixel phase
    const int blk = big.dt == SRB_BF16 ? 64 : 32, rr = z_ps * z_ps;
    if (z_ps > 8 || g.Co % rr != 0 || (g.Co / rr) % blk != 0 || g.Ci <= 4) return false;
  }
  if (big.dt == SRB_BF16) {
    if (g.Ci % 8 != 0 || g.Ci < 8 || g.Co % 8 != 0 || g.Co > 1024) return false;   // 16-byte pixel rows on both tensors
    if (small.sc != 1 || big.sc != 1) return false;
    if ((small.sw % 8) || (small.sh % 8) || (small.sn % 8) || (big.sw % 8) || (big.sh % 8) || (big.sn % 8)) return false;
    if ((((uintptr_t)small.p) | ((uintptr_t)big.p)) & 15) return false;
    if (g.kh > 16 || g.kw > 16) return false;
    WgPlan ph;
    return make_wg_plan(g, &ph, true, z_ps);
  }
  const bool c4 = g.Ci <= 4;
  if ((!c4 && g.Ci % 32 != 0) || g.Co % 4 != 0 || g.Co > 1024) return false;
  if (small.sc != 1 || (!c4 && big.sc != 1)) return false;
  if ((small.sw % 4) || (small.sh % 4) || (small.sn % 4)) return false;
  if (!c4 && ((big.sw % 4) || (big.sh % 4) || (big.sn % 4))) return false;
  if (((uintptr_t)small.p) & 15) return false;
  if (!c4 && (((uintptr_t)big.p) & 15)) return false;
  if (g.kh > 16 || g.kw > 16) return false;
  WgPlan pl;
  return make_wg_plan(g, &pl, false, z_ps);
}

int tc_wgrad_describe(const Geom &g, char *buf, size_t n, bool bf16) {
  WgPlan pl;
  if (g.st != 1 || g.ps != 1 || (!bf16 && g.Ci > 4 && g.Ci % 32 != 0) || !make_wg_plan(g, &pl, bf16)) return snprintf(buf, n, "tc_wgrad: no plan");
  const WgArgs &a = pl.a;
  return snprintf(buf, n,
                  "tc_wgrad%s: band TH %d TW %d x%d BW %d BH %d, bands %d (%d per CTA), CIB %d RG %d SG %d NT %d, groups ci %d r %d co %d, "
                  "stages %d, smem %zu B, tmem %d cols, grid %d x %d, ksteps %d",
                  a.bf16 ? (a.c2 ? "-bf16(ci pairs)" : "-bf16(tap pairs)") : (a.rn == 2 ? "-rows+co-stacked" : (a.rn ? "-rows-stacked" : "")), a.TH, a.TW, a.bands_w, a.BW, a.BH, a.num_bands, a.bands_per_cta, a.CIB, a.RG, a.SG, a.NT, a.n_cig, a.n_rg, a.n_cot, a.stages,
                  pl.smem, a.tmem_cols, pl.grid.x, pl.grid.y, a.ksteps);
}

size_t tc_wgrad_ws_bytes(const Geom &g, bool bf16, int z_ps) {
  WgPlan pl;
  if (g.st != 1 || g.ps != 1 || (!bf16 && g.Ci > 4 && g.Ci % 32 != 0) || !make_wg_plan(g, &pl, bf16, z_ps)) return 0;
  return (pl.partial_floats + pl.db_floats + pl.xpack_floats) * sizeof(float) + 1024;
}

int tc_conv_wgrad(const Geom &g, const T4 &small, const T4 &big, float *dw, float *db_small, float scale,
                  int accumulate, void *ws, size_t ws_bytes, cudaStream_t st, int z_ps, const WgPhase *phase) {
  WgPlan pl;
  const bool bf = big.dt == SRB_BF16;
  const PdlScope pdl_scope(2.0 * g.N * g.Ho * g.Wo * (double)g.Co * g.Ci * g.kh * g.kw < 2.0e10);
  SRB_REQUIRE(small.dt == big.dt, SRB_EINVAL, "tc_wgrad: x and dz must have one dtype");
  SRB_REQUIRE(make_wg_plan(g, &pl, bf, z_ps), SRB_EUNSUPPORTED, "tc_wgrad: no plan");
  size_t need = (pl.partial_floats + pl.db_floats + pl.xpack_floats) * sizeof(float) + 256;
  uintptr_t wsp = ((uintptr_t)ws + 255) & ~(uintptr_t)255;
  SRB_REQUIRE(ws && wsp + need <= (uintptr_t)ws + ws_bytes, SRB_EWORKSPACE, "tc_wgrad workspace: need %zu bytes, have %zu",
              need, ws_bytes);
  WgArgs &a = pl.a;
  {  // A-descriptor offsets of the accumulators, in 16-byte units (one pixel slot = 128 B = 8 units)
    const int row_step = a.BW * 8, x_step = a.x_slots * 8;
    for (int j = 0; j < kMaxAcc; ++j) a.acc_off[j] = a.acc_boff[j] = 0;
    if (a.rn == 2) {
      // accumulator (sg, cb): A = ci block cb at tap s = 4*sg, B = the whole [row][co block] tile (all co blocks in one MMA)
      for (int sg = 0; sg < a.SG; ++sg)
        for (int cb = 0; cb < a.CIB; ++cb) a.acc_off[sg * a.CIB + cb] = sg * 32 + cb * x_step;
    } else if (a.rn) {
      // accumulator (co block, sg, cb): A = ci block cb at tap s = 4*sg (un-shifted rows), B = co block's dz view
      for (int cob = 0; cob < a.NT / 32; ++cob)
        for (int sg = 0; sg < a.SG; ++sg)
          for (int cb = 0; cb < a.CIB; ++cb) {
            const int j = (cob * a.SG + sg) * a.CIB + cb;
            a.acc_off[j] = sg * 32 + cb * x_step;
            a.acc_boff[j] = cob * a.dz_slots * 8;
          }
    } else if (a.c4) {
      for (int rg = 0; rg < a.RG; ++rg)
        for (int sg = 0; sg < a.SG; ++sg) a.acc_off[rg * a.SG + sg] = 4 * rg * row_step + sg * 64;
    } else if (a.bf16) {
      // tap pairs: accumulator (rl, sg, cb) starts at tap s = 2*sg of ci-block cb; ci pairs: at tap s = sg of ci-block 2*cb
      for (int rl = 0; rl < a.RG; ++rl)
        for (int sg = 0; sg < a.SG; ++sg)
          for (int cb = 0; cb < a.CIB; ++cb)
            a.acc_off[(rl * a.SG + sg) * a.CIB + cb] = rl * row_step + (a.c2 ? sg * 8 + 2 * cb * x_step : sg * 16 + cb * x_step);
    } else {
      for (int rl = 0; rl < a.RG; ++rl)
        for (int sg = 0; sg < a.SG; ++sg)
          for (int cb = 0; cb < a.CIB; ++cb) a.acc_off[(rl * a.SG + sg) * a.CIB + cb] = rl * row_step + sg * 32 + cb * x_step;
    }
  }
  a.dbg = tc_conv_get_dbg();
  a.ph_st = 0; a.ph_a = a.ph_b = a.ph_pad0 = a.ph_kh0 = a.ph_kw0 = 0;
  if (phase) {
    SRB_REQUIRE(!a.c4 && z_ps == 1, SRB_EUNSUPPORTED, "strided wgrad phases need the generic operand flavour");
    a.ph_st = phase->st; a.ph_a = phase->a; a.ph_b = phase->b; a.ph_pad0 = phase->pad0; a.ph_kh0 = phase->kh0; a.ph_kw0 = phase->kw0;
  }
  a.partial = (float *)wsp;
  float *db_part = a.partial + pl.partial_floats;
  a.db_part = db_small ? db_part : nullptr;

  CUtensorMap mapX, mapZ;
  if (a.c4) {
    float *xp = (float *)(((uintptr_t)(db_part + pl.db_floats) + 255) & ~(uintptr_t)255);
    const long long total = (long long)(pl.xpack_floats / 4);
    int pb = (int)((total + 255) / 256);
    if (pb > 148 * 16) pb = 148 * 16;
    launch_pdl(k_pack_nhwc4_padded, dim3(pb), dim3(256), 0, st, big, (float4 *)xp, g.N, g.Ci, g.Hi, g.Wi, g.pad, pl.Hp, pl.Wp, total);
    count_launch();
    SRB_CHECK_CUDA(cudaGetLastError());
    // overlapping view: "channel" c of slot x is float 4*x + c of the padded row -> 8 pixels x 4 channels per slot
    cuuint64_t dims[4] = {32, (cuuint64_t)a.BW, (cuuint64_t)pl.Hp, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {16, (cuuint64_t)pl.Wp * 16, (cuuint64_t)pl.Hp * pl.Wp * 16};
    cuuint32_t box[4] = {32, (cuuint32_t)a.BW, (cuuint32_t)a.BH, 1};
    int rc = encode_tiled(&mapX, xp, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
  } else {
    const cuuint64_t es = bf ? 2 : 4;
    cuuint64_t dims[4] = {(cuuint64_t)g.Ci, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)big.sw * es, (cuuint64_t)big.sh * es, (cuuint64_t)big.sn * es};
    cuuint32_t box[4] = {(cuuint32_t)(bf ? 64 : 32), (cuuint32_t)a.BW, (cuuint32_t)a.BH, 1};
    int rc = encode_tiled(&mapX, big.p, 4, dims, strides, box, bf ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, bf);
    if (rc) return rc;
  }
  {
    // z_ps > 1: the map describes the SHUFFLED gradient (C, Wo*r, Ho*r, N), traversed with element strides (1, r, r, 1)
    const cuuint64_t es = bf ? 2 : 4, r = (cuuint64_t)z_ps;
    cuuint64_t dims[4] = {(cuuint64_t)(g.Co / (z_ps * z_ps)), (cuuint64_t)g.Wo * r, (cuuint64_t)g.Ho * r, (cuuint64_t)g.N};
    cuuint64_t strides[3] = {(cuuint64_t)small.sw * es, (cuuint64_t)small.sh * es, (cuuint64_t)small.sn * es};
    cuuint32_t box[4] = {(cuuint32_t)(bf ? 64 : 32), (cuuint32_t)((a.dz_rowwise ? a.TW : a.BW) * z_ps),
                         (cuuint32_t)((a.dz_rowwise ? 1 : (a.rn ? a.TH + a.kh - 1 : a.TH)) * z_ps), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)z_ps, (cuuint32_t)z_ps, 1};
    int rc = encode_tiled(&mapZ, small.p, 4, dims, strides, box, bf ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, bf,
                          z_ps > 1 ? estr : nullptr);
    if (rc) return rc;
  }
  {
    static std::atomic<unsigned long long> attr_done{0};
    int rc = ensure_kernel_attrs(k_tc_wgrad, attr_done, kMaxSmemBytes, false);
    if (rc) return rc;
  }
  SRB_CHECK_CUDA(launch_pdl(k_tc_wgrad, pl.grid, dim3(kWgThreads), pl.smem, st, mapX, mapZ, a));
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  {
    const int taps = g.kh * g.kw, total_items = g.Ci * taps, splits = (int)pl.grid.x;
    const int co_tiles = (g.Co + 31) / 32;
    // many splits (small layers, one output tile): all 8 warps of a block share each item; else one item per warp
    const int lanes8 = splits > 24 ? 1 : 0;
    int IPB = lanes8 ? 4 : 36;
    // keep at least ~2 blocks per SM when the layer is small
    while (IPB > (lanes8 ? 1 : 8) && (long long)co_tiles * ((total_items + IPB - 1) / IPB) < 296) IPB = (IPB + 1) / 2;
    const int groups = (total_items + IPB - 1) / IPB;
    const unsigned blocks = (unsigned)(co_tiles * groups + (db_small ? co_tiles : 0));
    const size_t fsm = ((size_t)32 * (IPB + 1) + 8 * 33) * sizeof(float);
    launch_pdl(k_wgrad_finish, dim3(blocks), dim3(256), fsm, st, a, splits, (int)pl.grid.y, IPB, lanes8, dw, db_small, scale, accumulate);
    count_launch();
    SRB_CHECK_CUDA(cudaGetLastError());
  }
  return SRB_OK;
}

}  // namespace srb
