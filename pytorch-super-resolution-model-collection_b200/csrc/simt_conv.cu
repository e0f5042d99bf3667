// fp32 CUDA-core implicit-GEMM convolution kernels (generic shapes / strides / layouts).
//
// These are the exact-fp32 kernels of the engine: they run every layer the tcgen05 path does not
// take (Cin = 3 network inputs, Cout = 3 gradients, strided / transposed convolutions, 1x1 with tiny
// channel counts) and serve as the on-device cross-check for the tensor-core kernels.
// Three GEMM shapes, all smem-tiled with register blocking, im2col only as index arithmetic:
//   gather  : small[n,co,oy,ox] = sum_{ci,r,s} big[n,ci,oy*st-pad+r,ox*st-pad+s] * w[co,ci,r,s]
//             (Conv2d fprop, base_networks.py:42,66;  ConvTranspose2d dgrad)
//   scatter : big[n,ci,iy,ix]   = sum_{co,r,s} small[n,co,(iy+pad-r)/st,(ix+pad-s)/st] * w[co,ci,r,s]
//             (Conv2d dgrad;  ConvTranspose2d fprop, base_networks.py:77, fsrcnn.py:33)
//   wgrad   : dw[co,ci,r,s]     = sum_{n,oy,ox} small[n,co,oy,ox] * big[n,ci,oy*st-pad+r,ox*st-pad+s]
//             (+ db[co] = sum small, as one extra GEMM column), split-K over pixels, deterministic reduce.
#include "srb_common.cuh"

namespace srb {

namespace {

constexpr int BK = 16;

template <int BM, int BN, int TM, int TN>
struct Tile {
  static constexpr int NT = (BM / TM) * (BN / TN);
  static constexpr int LDA = BM + 4;
  static constexpr int LDB = BN + 4;
};

// ---------------------------------------------------------------------------------------------
// gather: M = (n,oy,ox), N = co, K = (r,s,ci)
// ---------------------------------------------------------------------------------------------
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
k_gather(Geom g, T4 in, const float *__restrict__ w, T4 out, Epi epi) {
  using TL = Tile<BM, BN, TM, TN>;
  constexpr int NT = TL::NT;
  __shared__ float As[BK][TL::LDA];
  __shared__ float Bs[BK][TL::LDB];

  const int tid = threadIdx.x;
  const long long M = (long long)g.N * g.Ho * g.Wo;
  const int K = g.kh * g.kw * g.Ci;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // A-load assignment: one pixel per thread (mi fixed), BK/(NT/BM) k's.
  static_assert(NT % BM == 0 || BM % NT == 0, "tile/threads mismatch");
  constexpr int A_KSTEP = (NT >= BM) ? NT / BM : 1;
  constexpr int A_MREP = (NT >= BM) ? 1 : BM / NT;
  int a_n[A_MREP], a_iy0[A_MREP], a_ix0[A_MREP];
  bool a_ok[A_MREP];
#pragma unroll
  for (int q = 0; q < A_MREP; ++q) {
    int mi = (tid % BM) + q * NT;
    long long m = m0 + mi;
    a_ok[q] = m < M;
    long long mm = a_ok[q] ? m : 0;
    int ox = (int)(mm % g.Wo);
    long long t = mm / g.Wo;
    int oy = (int)(t % g.Ho);
    a_n[q] = (int)(t / g.Ho);
    a_iy0[q] = oy * g.st - g.pad;
    a_ix0[q] = ox * g.st - g.pad;
  }
  const int a_k0 = (NT >= BM) ? tid / BM : 0;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int tx = tid % (BN / TN), ty = tid / (BN / TN);

  for (int k0 = 0; k0 < K; k0 += BK) {
    // ---- A tile
#pragma unroll
    for (int kk = a_k0; kk < BK; kk += A_KSTEP) {
      int k = k0 + kk;
      int tap = k / g.Ci, ci = k - tap * g.Ci;
      int r = tap / g.kw, s = tap - r * g.kw;
#pragma unroll
      for (int q = 0; q < A_MREP; ++q) {
        float v = 0.f;
        int iy = a_iy0[q] + r, ix = a_ix0[q] + s;
        if (k < K && a_ok[q] && iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi)
          v = __ldg(in.p + a_n[q] * in.sn + ci * in.sc + iy * in.sh + ix * in.sw);
        As[kk][(tid % BM) + q * NT] = v;
      }
    }
    // ---- B tile: element (kk, ni) = w[co=n0+ni][ci][r][s]
    for (int e = tid; e < BK * BN; e += NT) {
      int ni = e % BN, kk = e / BN;
      int k = k0 + kk, co = n0 + ni;
      float v = 0.f;
      if (k < K && co < g.Co) {
        int tap = k / g.Ci, ci = k - tap * g.Ci;
        v = __ldg(w + ((long long)co * g.Ci + ci) * (g.kh * g.kw) + tap);
      }
      Bs[kk][ni] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue: bias -> act -> (+residual) -> store at the pixel-shuffled address
  const float slope = (epi.act == SRB_ACT_PRELU) ? __ldg(epi.alpha) : epi.slope;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    long long m = m0 + ty * TM + i;
    if (m >= M) continue;
    int ox = (int)(m % g.Wo);
    long long t = m / g.Wo;
    int oy = (int)(t % g.Ho);
    int n = (int)(t / g.Ho);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int co = n0 + tx * TN + j;
      if (co >= g.Co) continue;
      float z = acc[i][j] + (epi.bias ? __ldg(epi.bias + co) : 0.f);
      long long off = ps_offset(out, g.ps, n, co, oy, ox);
      if (epi.preact.p) epi.preact.p[ps_offset(epi.preact, g.ps, n, co, oy, ox)] = z;
      float y = apply_act(z, epi.act, slope);
      if (epi.residual.p) y += __ldg(epi.residual.p + ps_offset(epi.residual, g.ps, n, co, oy, ox));
      if (epi.mask.p && !(__ldg(epi.mask.p + ps_offset(epi.mask, g.ps, n, co, oy, ox)) > 0.f)) y = 0.f;
      if (epi.round_tf32) y = round_tf32(y);
      out.p[off] = y;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// scatter: M = (n,iy,ix) of the big tensor, N = ci, K = (r,s,co)
// ---------------------------------------------------------------------------------------------
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
k_scatter(Geom g, T4 in, const float *__restrict__ w, T4 out, Epi epi) {
  using TL = Tile<BM, BN, TM, TN>;
  constexpr int NT = TL::NT;
  __shared__ float As[BK][TL::LDA];
  __shared__ float Bs[BK][TL::LDB];

  const int tid = threadIdx.x;
  const long long M = (long long)g.N * g.Hi * g.Wi;
  const int K = g.kh * g.kw * g.Co;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  constexpr int A_KSTEP = (NT >= BM) ? NT / BM : 1;
  constexpr int A_MREP = (NT >= BM) ? 1 : BM / NT;
  int a_n[A_MREP], a_y[A_MREP], a_x[A_MREP];
  bool a_ok[A_MREP];
#pragma unroll
  for (int q = 0; q < A_MREP; ++q) {
    int mi = (tid % BM) + q * NT;
    long long m = m0 + mi;
    a_ok[q] = m < M;
    long long mm = a_ok[q] ? m : 0;
    int ix = (int)(mm % g.Wi);
    long long t = mm / g.Wi;
    int iy = (int)(t % g.Hi);
    a_n[q] = (int)(t / g.Hi);
    a_y[q] = iy + g.pad;
    a_x[q] = ix + g.pad;
  }
  const int a_k0 = (NT >= BM) ? tid / BM : 0;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);

  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int kk = a_k0; kk < BK; kk += A_KSTEP) {
      int k = k0 + kk;
      int tap = k / g.Co, co = k - tap * g.Co;
      int r = tap / g.kw, s = tap - r * g.kw;
#pragma unroll
      for (int q = 0; q < A_MREP; ++q) {
        float v = 0.f;
        int ty_ = a_y[q] - r, tx_ = a_x[q] - s;
        if (k < K && a_ok[q] && ty_ >= 0 && tx_ >= 0) {
          int oy = ty_ / g.st, ox = tx_ / g.st;
          if (oy * g.st == ty_ && ox * g.st == tx_ && oy < g.Ho && ox < g.Wo)
            v = __ldg(in.p + ps_offset(in, g.ps, a_n[q], co, oy, ox));
        }
        As[kk][(tid % BM) + q * NT] = v;
      }
    }
    for (int e = tid; e < BK * BN; e += NT) {
      int ni = e % BN, kk = e / BN;
      int k = k0 + kk, ci = n0 + ni;
      float v = 0.f;
      if (k < K && ci < g.Ci) {
        int tap = k / g.Co, co = k - tap * g.Co;
        v = __ldg(w + ((long long)co * g.Ci + ci) * (g.kh * g.kw) + tap);
      }
      Bs[kk][ni] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const float slope = (epi.act == SRB_ACT_PRELU) ? __ldg(epi.alpha) : epi.slope;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    long long m = m0 + ty * TM + i;
    if (m >= M) continue;
    int ix = (int)(m % g.Wi);
    long long t = m / g.Wi;
    int iy = (int)(t % g.Hi);
    int n = (int)(t / g.Hi);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int ci = n0 + tx * TN + j;
      if (ci >= g.Ci) continue;
      float z = acc[i][j] + (epi.bias ? __ldg(epi.bias + ci) : 0.f);
      long long off = n * out.sn + ci * out.sc + iy * out.sh + ix * out.sw;
      if (epi.preact.p) epi.preact.p[n * epi.preact.sn + ci * epi.preact.sc + iy * epi.preact.sh + ix * epi.preact.sw] = z;
      float y = apply_act(z, epi.act, slope);
      if (epi.residual.p)
        y += __ldg(epi.residual.p + n * epi.residual.sn + ci * epi.residual.sc + iy * epi.residual.sh + ix * epi.residual.sw);
      if (epi.mask.p && !(__ldg(epi.mask.p + n * epi.mask.sn + ci * epi.mask.sc + iy * epi.mask.sh + ix * epi.mask.sw) > 0.f)) y = 0.f;
      if (epi.round_tf32) y = round_tf32(y);
      out.p[off] = y;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad: M = co, N = (ci,r,s) [+1 column of ones -> db], K = (n,oy,ox) split over blockIdx.z
// partial[z][co][NN]  with NN = Ci*kh*kw + 1
// ---------------------------------------------------------------------------------------------
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
k_wgrad(Geom g, T4 sm, T4 big, float *__restrict__ partial, long long kchunk) {
  using TL = Tile<BM, BN, TM, TN>;
  constexpr int NT = TL::NT;
  __shared__ float As[BK][TL::LDA];
  __shared__ float Bs[BK][TL::LDB];

  const int tid = threadIdx.x;
  const long long Kp = (long long)g.N * g.Ho * g.Wo;
  const int CRS = g.Ci * g.kh * g.kw;
  const int NN = CRS + 1;
  const int m0 = blockIdx.x * BM;  // co
  const int n0 = blockIdx.y * BN;  // column
  const long long kbeg = (long long)blockIdx.z * kchunk;
  long long kend = kbeg + kchunk;
  if (kend > Kp) kend = Kp;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);

  // B-load: this thread always loads the same column(s); pre-decode (ci,r,s).
  constexpr int B_KSTEP = (NT >= BN) ? NT / BN : 1;
  constexpr int B_NREP = (NT >= BN) ? 1 : BN / NT;
  int b_ci[B_NREP], b_r[B_NREP], b_s[B_NREP], b_kind[B_NREP];  // kind: 0 invalid, 1 data, 2 ones
#pragma unroll
  for (int q = 0; q < B_NREP; ++q) {
    int col = n0 + (tid % BN) + q * NT;
    if (col < CRS) {
      b_kind[q] = 1;
      b_ci[q] = col / (g.kh * g.kw);
      int rs = col - b_ci[q] * (g.kh * g.kw);
      b_r[q] = rs / g.kw;
      b_s[q] = rs - b_r[q] * g.kw;
    } else {
      b_kind[q] = (col == CRS) ? 2 : 0;
      b_ci[q] = b_r[q] = b_s[q] = 0;
    }
  }
  const int b_k0 = (NT >= BN) ? tid / BN : 0;

  for (long long k0 = kbeg; k0 < kend; k0 += BK) {
    // A tile: (kk, mi) = small[n, co=m0+mi, oy, ox]; kk fastest over threads (pixels contiguous in NCHW)
    for (int e = tid; e < BK * BM; e += NT) {
      int kk = e % BK, mi = e / BK;
      long long k = k0 + kk;
      int co = m0 + mi;
      float v = 0.f;
      if (k < kend && co < g.Co) {
        int ox = (int)(k % g.Wo);
        long long t = k / g.Wo;
        int oy = (int)(t % g.Ho);
        int n = (int)(t / g.Ho);
        v = __ldg(sm.p + ps_offset(sm, g.ps, n, co, oy, ox));
      }
      As[kk][mi] = v;
    }
#pragma unroll
    for (int kk = b_k0; kk < BK; kk += B_KSTEP) {
      long long k = k0 + kk;
      bool kok = k < kend;
      long long kc = kok ? k : 0;
      int ox = (int)(kc % g.Wo);
      long long t = kc / g.Wo;
      int oy = (int)(t % g.Ho);
      int n = (int)(t / g.Ho);
#pragma unroll
      for (int q = 0; q < B_NREP; ++q) {
        float v = 0.f;
        if (kok && b_kind[q] == 1) {
          int iy = oy * g.st - g.pad + b_r[q], ix = ox * g.st - g.pad + b_s[q];
          if (iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi)
            v = __ldg(big.p + n * big.sn + b_ci[q] * big.sc + iy * big.sh + ix * big.sw);
        } else if (kok && b_kind[q] == 2) {
          v = 1.f;
        }
        Bs[kk][(tid % BN) + q * NT] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  float *dst = partial + (long long)blockIdx.z * g.Co * NN;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int co = m0 + ty * TM + i;
    if (co >= g.Co) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int col = n0 + tx * TN + j;
      if (col < NN) dst[(long long)co * NN + col] = acc[i][j];
    }
  }
}

// Sum the split-K partials in a fixed order (deterministic), scale, store / accumulate.
__global__ void k_wgrad_reduce(const float *__restrict__ partial, int splits, int Co, int CRS, float *dw, float *db,
                               float scale, int accumulate) {
  const int NN = CRS + 1;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)Co * NN;
  if (idx >= total) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += partial[(long long)z * total + idx];
  s *= scale;
  int co = (int)(idx / NN), col = (int)(idx - (long long)co * NN);
  if (col < CRS) {
    float *d = dw + (long long)co * CRS + col;
    *d = accumulate ? (*d + s) : s;
  } else if (db) {
    db[co] = accumulate ? (db[co] + s) : s;
  }
}

// out[c] (=|+=) scale * sum_{n,h,w} t[n,c,h,w]   (bias gradient of a transposed conv)
__global__ void k_channel_sum(T4 t, int N, int C, int H, int W, float *out, float scale, int accumulate) {
  int c = blockIdx.x;
  long long total = (long long)N * H * W;
  float s = 0.f;
  for (long long i = threadIdx.x; i < total; i += blockDim.x) {
    int x = (int)(i % W);
    long long q = i / W;
    int y = (int)(q % H);
    int n = (int)(q / H);
    s += __ldg(t.p + n * t.sn + c * t.sc + y * t.sh + x * t.sw);
  }
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) out[c] = accumulate ? (out[c] + s * scale) : s * scale;
  }
}


// ---------------------------------------------------------------------------------------------
// wgrad for skinny outputs (Co <= 4: the 64->3 / 32->3 tail layers of SRCNN / VDSR / EDSR / SRGAN-G), stride 1,
// channels_last x.  The GEMM tiling above would waste 61 of 64 rows; this kernel is a direct reduction instead:
//   thread (ci, r) owns dw[0..CO)[ci][r][0..KW)  (KW*CO accumulators in registers)
// and walks the INPUT rows of its band: input row iy pairs with output row oy = iy - r + pad, so every x element is
// fetched from global once per CTA (the kh filter-row warps hit the same lines in L1) with ci contiguous across the
// warp (128-byte coalesced), while the CO dz values of a pixel are warp-uniform broadcast loads kept in a KW-deep
// register window.  One partial dW per band goes to the workspace in k_wgrad's [co][NN] layout (db = column CRS);
// k_wgrad_reduce_par sums the bands in a fixed order.
// ---------------------------------------------------------------------------------------------
template <int KW, int CO>
__global__ void __launch_bounds__(KW == 9 ? 288 : 64 * KW, KW == 3 ? 3 : (KW == 1 ? 6 : 1))
k_wgrad_smallco(Geom g, T4 sm, T4 big, float *__restrict__ partial, int TH, int bands_per_img, int ci_tile, int r_db) {
  constexpr int U = KW == 9 ? 9 : (KW == 5 ? 10 : 12);  // pixels per chunk: all loads of a chunk are issued before its FMAs
  constexpr int KC = KW > 1 ? KW - 1 : 1;
  const int band = blockIdx.x;
  const int n = band / bands_per_img;
  const int iy0 = (band - n * bands_per_img) * TH;
  const int iy1 = min(iy0 + TH, g.Hi);
  const int ci = blockIdx.y * ci_tile + (int)(threadIdx.x % ci_tile);
  const int r = threadIdx.x / ci_tile;
  const bool ci_ok = ci < g.Ci;
  const bool do_db = (ci == 0) && (r == r_db);
  float acc[KW][CO];
  float dbs[CO];
#pragma unroll
  for (int s = 0; s < KW; ++s)
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[s][c] = 0.f;
#pragma unroll
  for (int c = 0; c < CO; ++c) dbs[c] = 0.f;

  for (int iy = iy0; iy < iy1; ++iy) {
    const int oy = iy - r + g.pad;
    if (oy < 0 || oy >= g.Ho) continue;  // uniform per warp (ci_tile is a multiple of 32)
    // row-local offsets fit 32 bits (one image row of one tensor); the row bases carry the 64-bit part
    const float *xr = big.p + n * big.sn + (long long)iy * big.sh + (ci_ok ? ci : 0) * big.sc;
    const float *zrow[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) zrow[c] = sm.p + n * sm.sn + (long long)oy * sm.sh + c * sm.sc;
    const int xsw = (int)big.sw, zsw = (int)sm.sw;
    float zc[CO][KC];  // carried window: zc[c][t] = dz[c][oy][ix0 + pad - (KW-1) + t]
#pragma unroll
    for (int t = 0; t < KW - 1; ++t) {
      const int ox = g.pad - (KW - 1) + t;
#pragma unroll
      for (int c = 0; c < CO; ++c) {
        const float v = (ox >= 0 && ox < g.Wo) ? __ldg(zrow[c] + ox * zsw) : 0.f;
        zc[c][t] = v;
        dbs[c] += v;
      }
    }
    for (int ix0 = 0; ix0 < g.Wi; ix0 += U) {
      float xv[U], zv[CO][U + KW - 1];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int ix = ix0 + j, ox = ix + g.pad;
        xv[j] = (ix < g.Wi && ci_ok) ? __ldg(xr + ix * xsw) : 0.f;
#pragma unroll
        for (int c = 0; c < CO; ++c) zv[c][KW - 1 + j] = (ix < g.Wi && ox < g.Wo) ? __ldg(zrow[c] + ox * zsw) : 0.f;
      }
#pragma unroll
      for (int c = 0; c < CO; ++c) {
#pragma unroll
        for (int t = 0; t < KW - 1; ++t) zv[c][t] = zc[c][t];
#pragma unroll
        for (int j = 0; j < U; ++j) dbs[c] += zv[c][KW - 1 + j];
      }
#pragma unroll
      for (int j = 0; j < U; ++j)
#pragma unroll
        for (int s = 0; s < KW; ++s)
#pragma unroll
          for (int c = 0; c < CO; ++c) acc[s][c] = fmaf(xv[j], zv[c][KW - 1 + j - s], acc[s][c]);
#pragma unroll
      for (int c = 0; c < CO; ++c)
#pragma unroll
        for (int t = 0; t < KW - 1; ++t) zc[c][t] = zv[c][U + t];
    }
  }
  const int CRS = g.Ci * g.kh * KW, NN = CRS + 1;
  float *dst = partial + (long long)band * CO * NN;
  if (ci_ok) {
#pragma unroll
    for (int c = 0; c < CO; ++c)
#pragma unroll
      for (int s = 0; s < KW; ++s) dst[(long long)c * NN + (ci * g.kh + r) * KW + s] = acc[s][c];
  }
  if (do_db) {
#pragma unroll
    for (int c = 0; c < CO; ++c) dst[(long long)c * NN + CRS] = dbs[c];
  }
}

// Parallel, deterministic version of k_wgrad_reduce for many splits: block = (32 outputs, 8 split lanes), lane y sums
// splits y, y+8, ...; the eight lane sums are folded in a fixed order.
__global__ void __launch_bounds__(256)
k_wgrad_reduce_par(const float *__restrict__ partial, int splits, int Co, int CRS, float *dw, float *db, float scale,
                   int accumulate) {
  __shared__ float red[8][33];
  const int NN = CRS + 1;
  const long long total = (long long)Co * NN;
  const long long idx = (long long)blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (idx < total)
    for (int z = threadIdx.y; z < splits; z += 8) s += partial[(long long)z * total + idx];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && idx < total) {
    float t = red[0][threadIdx.x];
#pragma unroll
    for (int y = 1; y < 8; ++y) t += red[y][threadIdx.x];
    t *= scale;
    const int co = (int)(idx / NN), col = (int)(idx - (long long)co * NN);
    if (col < CRS) {
      float *d = dw + (long long)co * CRS + col;
      *d = accumulate ? (*d + t) : t;
    } else if (db) {
      db[co] = accumulate ? (db[co] + t) : t;
    }
  }
}

struct SmallCoPlan {
  int TH, bands_per_img, bands, ci_tile, ci_tiles, r_db;
};

// Applies to: Conv2d-orientation wgrad, stride 1, no PixelShuffle, Co <= 4, kw in {1,3,5,9}, channels_last x,
// and padding no larger than "same" (so that one filter row sees every output row: the db column).
bool smallco_plan(const Geom &g, const T4 &big, SmallCoPlan *p) {
  if (g.st != 1 || g.ps != 1 || g.Co > 4 || g.Co < 1 || g.N <= 0) return false;
  if (g.kh != g.kw || (g.kw != 1 && g.kw != 3 && g.kw != 5 && g.kw != 9)) return false;
  if (big.p && big.sc != 1) return false;
  if (2 * g.pad > g.kh - 1 || g.pad > g.kw - 1) return false;
  const int ci_tile = (g.Ci >= 64 && g.kw < 9) ? 64 : 32;  // threads = ci_tile * kh <= the kernel's launch bound
  p->ci_tile = ci_tile;
  p->ci_tiles = (g.Ci + ci_tile - 1) / ci_tile;
  p->r_db = g.pad;
  const long long rows = (long long)g.N * g.Hi;
  long long target = 148LL * 8 / p->ci_tiles;  // ~8 CTAs per SM in flight
  if (target < 1) target = 1;
  int TH = (int)((rows + target - 1) / target);
  if (TH < 1) TH = 1;
  if (TH > g.Hi) TH = g.Hi;
  p->TH = TH;
  p->bands_per_img = (g.Hi + TH - 1) / TH;
  p->bands = g.N * p->bands_per_img;
  return true;
}

template <int KW>
void launch_smallco(const Geom &g, const T4 &small, const T4 &big, float *partial, const SmallCoPlan &p, cudaStream_t st) {
  dim3 grid((unsigned)p.bands, (unsigned)p.ci_tiles);
  const int threads = p.ci_tile * g.kh;
  switch (g.Co) {
    case 1: k_wgrad_smallco<KW, 1><<<grid, threads, 0, st>>>(g, small, big, partial, p.TH, p.bands_per_img, p.ci_tile, p.r_db); break;
    case 2: k_wgrad_smallco<KW, 2><<<grid, threads, 0, st>>>(g, small, big, partial, p.TH, p.bands_per_img, p.ci_tile, p.r_db); break;
    case 3: k_wgrad_smallco<KW, 3><<<grid, threads, 0, st>>>(g, small, big, partial, p.TH, p.bands_per_img, p.ci_tile, p.r_db); break;
    default: k_wgrad_smallco<KW, 4><<<grid, threads, 0, st>>>(g, small, big, partial, p.TH, p.bands_per_img, p.ci_tile, p.r_db); break;
  }
}

}  // namespace

int simt_conv_gather(const Geom &g, const T4 &in, const float *w, const T4 &out, const Epi &epi, cudaStream_t st) {
  long long M = (long long)g.N * g.Ho * g.Wo;
  if (M == 0 || g.Co == 0) return SRB_OK;
  if (g.Co <= 8) {
    constexpr int BM = 128, BN = 8, TM = 2, TN = 4;
    dim3 grid((unsigned)((M + BM - 1) / BM), (g.Co + BN - 1) / BN);
    k_gather<BM, BN, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, st>>>(g, in, w, out, epi);
  } else {
    constexpr int BM = 64, BN = 64, TM = 4, TN = 4;
    dim3 grid((unsigned)((M + BM - 1) / BM), (g.Co + BN - 1) / BN);
    k_gather<BM, BN, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, st>>>(g, in, w, out, epi);
  }
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

int simt_conv_scatter(const Geom &g, const T4 &in_small, const float *w, const T4 &out_big, const Epi &epi,
                      cudaStream_t st) {
  long long M = (long long)g.N * g.Hi * g.Wi;
  if (M == 0 || g.Ci == 0) return SRB_OK;
  if (g.Ci <= 8) {
    constexpr int BM = 128, BN = 8, TM = 2, TN = 4;
    dim3 grid((unsigned)((M + BM - 1) / BM), (g.Ci + BN - 1) / BN);
    k_scatter<BM, BN, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, st>>>(g, in_small, w, out_big, epi);
  } else {
    constexpr int BM = 64, BN = 64, TM = 4, TN = 4;
    dim3 grid((unsigned)((M + BM - 1) / BM), (g.Ci + BN - 1) / BN);
    k_scatter<BM, BN, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, st>>>(g, in_small, w, out_big, epi);
  }
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

static int wgrad_splits(const Geom &g) {
  long long Kp = (long long)g.N * g.Ho * g.Wo;
  int NN = g.Ci * g.kh * g.kw + 1;
  long long tiles = (long long)((g.Co + 63) / 64) * ((NN + 63) / 64);
  long long want = (148LL * 4 + tiles - 1) / tiles;  // ~4 waves worth of CTAs
  long long maxs = (Kp + 255) / 256;                  // at least 256 pixels per split
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 1024) want = 1024;
  return (int)want;
}

size_t simt_wgrad_ws_bytes(const Geom &g) {
  int NN = g.Ci * g.kh * g.kw + 1;
  size_t a = (size_t)wgrad_splits(g) * g.Co * NN * sizeof(float);
  SmallCoPlan sp;
  T4 none{nullptr, 0, 0, 0, 0};
  if (smallco_plan(g, none, &sp)) {
    size_t b = (size_t)sp.bands * g.Co * NN * sizeof(float);
    if (b > a) a = b;
  }
  return a;
}

int simt_conv_wgrad(const Geom &g, const T4 &small, const T4 &big, float *dw, float *db_small, float scale,
                    int accumulate, void *ws, size_t ws_bytes, cudaStream_t st) {
  long long Kp = (long long)g.N * g.Ho * g.Wo;
  int CRS = g.Ci * g.kh * g.kw, NN = CRS + 1;
  SmallCoPlan sp;
  if (smallco_plan(g, big, &sp)) {
    size_t need_s = (size_t)sp.bands * g.Co * NN * sizeof(float);
    SRB_REQUIRE(ws && ws_bytes >= need_s, SRB_EWORKSPACE, "wgrad workspace: need %zu bytes, got %zu", need_s, ws_bytes);
    switch (g.kw) {
      case 1: launch_smallco<1>(g, small, big, (float *)ws, sp, st); break;
      case 3: launch_smallco<3>(g, small, big, (float *)ws, sp, st); break;
      case 5: launch_smallco<5>(g, small, big, (float *)ws, sp, st); break;
      default: launch_smallco<9>(g, small, big, (float *)ws, sp, st); break;
    }
    count_launch();
    SRB_CHECK_CUDA(cudaGetLastError());
    const long long total_s = (long long)g.Co * NN;
    k_wgrad_reduce_par<<<(unsigned)((total_s + 31) / 32), dim3(32, 8), 0, st>>>((const float *)ws, sp.bands, g.Co, CRS, dw,
                                                                                 db_small, scale, accumulate);
    count_launch();
    SRB_CHECK_CUDA(cudaGetLastError());
    return SRB_OK;
  }
  int splits = wgrad_splits(g);
  size_t need = (size_t)splits * g.Co * NN * sizeof(float);
  SRB_REQUIRE(ws && ws_bytes >= need, SRB_EWORKSPACE, "wgrad workspace: need %zu bytes, got %zu", need, ws_bytes);
  long long kchunk = (Kp + splits - 1) / splits;
  kchunk = (kchunk + BK - 1) / BK * BK;
  constexpr int BM = 64, BN = 64, TM = 4, TN = 4;
  dim3 grid((g.Co + BM - 1) / BM, (NN + BN - 1) / BN, splits);
  k_wgrad<BM, BN, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, st>>>(g, small, big, (float *)ws, kchunk);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  long long total = (long long)g.Co * NN;
  k_wgrad_reduce<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const float *)ws, splits, g.Co, CRS, dw, db_small,
                                                                   scale, accumulate);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

int channel_sum(const T4 &t, int N, int C, int H, int W, float *out, float scale, int accumulate, cudaStream_t st) {
  if (C == 0) return SRB_OK;
  k_channel_sum<<<C, 256, 0, st>>>(t, N, C, H, W, out, scale, accumulate);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

}  // namespace srb
