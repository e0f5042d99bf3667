// math = SRB_MATH_EXACT: fp32-accurate convolutions on the tf32 tensor cores by operand splitting ("3xTF32").
//
//   x = x_hi + x_lo,  w = w_hi + w_lo   with  hi = RN_tf32(v),  lo = RN_tf32(v - hi)   (v - hi is exact in fp32)
//   x*w  ~=  x_hi*w_hi + x_lo*w_hi + x_hi*w_lo          (dropped: x_lo*w_lo ~ 2^-22 |x w|; fp32 accumulate in TMEM)
//
// The three partial products are not three kernels: the split is expressed along the CHANNEL (reduction) axis, so the
// unchanged slot-linear kernel (tc_conv_sl.cu) runs ONE convolution with 3*Cin input channels
//       X' = [x_hi | x_lo | x_hi]   (NHWC, channel count rounded up to a multiple of 4, zero filled)
//       W' = [w_hi | w_hi | w_lo]   (same channel order)
// and every fused epilogue feature (bias, activation, residual, PixelShuffle store, packed ReLU bits) keeps working.
// dgrad does the same with dz and the output-channel axis of w.  wgrad's reduction axis is the pixel axis, so it runs
// the tensor-core wgrad three times on (x_hi, dz_hi), (x_lo, dz_hi), (x_hi, dz_lo), accumulating into dw.
// Activations are stored as full fp32 in this mode (no tf32 rounding in any epilogue).
// Why it exists: TF32 is 3-6e-4 per layer but 1.4-1.6e-3 through 20..69 layers; the north_star
// contract is 1e-3 against the fp32 reference at the BASELINE depths.
#include "srb_common.cuh"

namespace srb {

namespace {

__device__ __forceinline__ void split_tf32(float v, float &hi, float &lo) {
  hi = round_tf32(v);
  lo = round_tf32(v - hi);
}

// x (N,C,H,W; any strides) -> out NHWC with C3 >= 3C channels per pixel: [hi(0..C) | lo(0..C) | hi(0..C) | 0...]
__global__ void k_split3_nhwc(T4 x, float *__restrict__ out, int N, int C, int H, int W, int C3) {
  const long long total = (long long)N * H * W * C3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c3 = (int)(i % C3);
    long long q = i / C3;
    const int w = (int)(q % W); q /= W;
    const int h = (int)(q % H);
    const int n = (int)(q / H);
    float v = 0.f;
    if (c3 < 3 * C) {
      const int part = c3 / C, c = c3 - part * C;
      float hi, lo;
      split_tf32(__ldg(x.p + n * x.sn + c * x.sc + (long long)h * x.sh + (long long)w * x.sw), hi, lo);
      v = part == 1 ? lo : hi;
    }
    out[i] = v;
  }
}

// x (N,C,H,W; any strides) -> hi, lo (both dense NHWC)
__global__ void k_split2_nhwc(T4 x, float *__restrict__ hi_out, float *__restrict__ lo_out, int N, int C, int H, int W) {
  const long long total = (long long)N * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long q = i / C;
    const int w = (int)(q % W); q /= W;
    const int h = (int)(q % H);
    const int n = (int)(q / H);
    float hi, lo;
    split_tf32(__ldg(x.p + n * x.sn + c * x.sc + (long long)h * x.sh + (long long)w * x.sw), hi, lo);
    hi_out[i] = hi;
    lo_out[i] = lo;
  }
}

// w (O, I, taps) contiguous -> W' with the split laid along I (along_o == 0: out (O, I3, taps)) or along O
// (along_o == 1: out (O3, I, taps)), parts [hi | hi | lo], zero fill beyond 3x.
__global__ void k_split_w(const float *__restrict__ w, float *__restrict__ out, int O, int I, int taps, int X3, int along_o) {
  const long long total = along_o ? (long long)X3 * I * taps : (long long)O * X3 * taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    long long q = i / taps;
    int o, ii, part;
    bool valid;
    if (along_o) {
      ii = (int)(q % I);
      const int o3 = (int)(q / I);
      part = o3 / O; o = o3 - part * O;
      valid = o3 < 3 * O;
    } else {
      const int i3 = (int)(q % X3);
      o = (int)(q / X3);
      part = i3 / I; ii = i3 - part * I;
      valid = i3 < 3 * I;
    }
    float v = 0.f;
    if (valid) {
      float hi, lo;
      split_tf32(__ldg(w + ((long long)o * I + ii) * taps + t), hi, lo);
      v = part == 2 ? lo : hi;
    }
    out[i] = v;
  }
}

inline unsigned blocks_for(long long n) {
  long long b = (n + 255) / 256;
  if (b > 148LL * 16) b = 148LL * 16;
  if (b < 1) b = 1;
  return (unsigned)b;
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

inline int tripled(int c) { return (3 * c + 3) / 4 * 4; }

// Geometry of the channel-tripled gather conv and a dense NHWC view of its (workspace) input.
inline Geom tripled_geom(const Geom &g) {
  Geom g3 = g;
  g3.Ci = tripled(g.Ci);
  return g3;
}
inline T4 dense_nhwc(float *p, int C, int H, int W) {
  return T4{p, (long long)H * W * C, 1, (long long)W * C, C};
}

}  // namespace

bool exact_conv_supported(const Geom &g, const T4 &out) {
  if (g.st != 1) return false;
  const Geom g3 = tripled_geom(g);
  return tc_conv_supported(g3, dense_nhwc((float *)256, g3.Ci, g.Hi, g.Wi), out, false);
}

size_t exact_conv_ws_bytes(const Geom &g) {
  const Geom g3 = tripled_geom(g);
  return align256((size_t)g.N * g.Hi * g.Wi * g3.Ci * sizeof(float)) +
         align256((size_t)g.Co * g3.Ci * g.kh * g.kw * sizeof(float)) + tc_conv_ws_bytes(g3) + 1024;
}

// `g` is the gather geometry of THIS launch (for dgrad: the flipped geometry, g.Ci = the layer's Cout); w is the layer's
// filter in its original (Cout, Cin, kh, kw) layout.
int exact_conv_gather(const Geom &g, const T4 &in, const float *w, bool flip_transpose, const T4 &out, const Epi &epi,
                      void *ws, size_t ws_bytes, cudaStream_t st) {
  const Geom g3 = tripled_geom(g);
  const size_t xb = align256((size_t)g.N * g.Hi * g.Wi * g3.Ci * sizeof(float));
  const size_t wb = align256((size_t)g.Co * g3.Ci * g.kh * g.kw * sizeof(float));
  uintptr_t wsp = ((uintptr_t)ws + 255) & ~(uintptr_t)255;
  SRB_REQUIRE(ws && wsp + xb + wb <= (uintptr_t)ws + ws_bytes, SRB_EWORKSPACE, "exact conv workspace: need %zu bytes, have %zu",
              xb + wb + 256, ws_bytes);
  float *x3 = (float *)wsp, *w3 = (float *)(wsp + xb);
  void *ws_tc = (void *)(wsp + xb + wb);
  k_split3_nhwc<<<blocks_for((long long)g.N * g.Hi * g.Wi * g3.Ci), 256, 0, st>>>(in, x3, g.N, g.Ci, g.Hi, g.Wi, g3.Ci);
  // fprop: filter (Co, Ci, taps), split along Ci.  dgrad: filter (g.Ci, g.Co, taps) = (layer Cout, layer Cin), split along
  // its first axis (the launch's reduction axis)
  if (!flip_transpose)
    k_split_w<<<blocks_for((long long)g.Co * g3.Ci * g.kh * g.kw), 256, 0, st>>>(w, w3, g.Co, g.Ci, g.kh * g.kw, g3.Ci, 0);
  else
    k_split_w<<<blocks_for((long long)g3.Ci * g.Co * g.kh * g.kw), 256, 0, st>>>(w, w3, g.Ci, g.Co, g.kh * g.kw, g3.Ci, 1);
  count_launch(2);
  SRB_CHECK_CUDA(cudaGetLastError());
  return tc_conv_gather(g3, dense_nhwc(x3, g3.Ci, g.Hi, g.Wi), w3, flip_transpose, out, epi, ws_tc,
                        (size_t)((uintptr_t)ws + ws_bytes - (uintptr_t)ws_tc), st);
}

bool exact_wgrad_supported(const Geom &g) {
  if (g.st != 1 || g.ps != 1 || g.Ci <= 4) return false;  // Cin <= 4 layers: the CUDA-core fp32 wgrad is exact and tiny
  return tc_wgrad_supported(g, dense_nhwc((float *)256, g.Co, g.Ho, g.Wo), dense_nhwc((float *)256, g.Ci, g.Hi, g.Wi));
}

size_t exact_wgrad_ws_bytes(const Geom &g) {
  return 2 * align256((size_t)g.N * g.Hi * g.Wi * g.Ci * sizeof(float)) +
         2 * align256((size_t)g.N * g.Ho * g.Wo * g.Co * sizeof(float)) + tc_wgrad_ws_bytes(g) + 1024;
}

int exact_conv_wgrad(const Geom &g, const T4 &small, const T4 &big, float *dw, float *db_small, float scale, int accumulate,
                     void *ws, size_t ws_bytes, cudaStream_t st) {
  const size_t xb = align256((size_t)g.N * g.Hi * g.Wi * g.Ci * sizeof(float));
  const size_t zb = align256((size_t)g.N * g.Ho * g.Wo * g.Co * sizeof(float));
  uintptr_t wsp = ((uintptr_t)ws + 255) & ~(uintptr_t)255;
  SRB_REQUIRE(ws && wsp + 2 * xb + 2 * zb <= (uintptr_t)ws + ws_bytes, SRB_EWORKSPACE,
              "exact wgrad workspace: need %zu bytes, have %zu", 2 * xb + 2 * zb + 256, ws_bytes);
  float *xh = (float *)wsp, *xl = (float *)(wsp + xb), *zh = (float *)(wsp + 2 * xb), *zl = (float *)(wsp + 2 * xb + zb);
  void *ws_tc = (void *)(wsp + 2 * xb + 2 * zb);
  const size_t ws_tc_bytes = (size_t)((uintptr_t)ws + ws_bytes - (uintptr_t)ws_tc);
  k_split2_nhwc<<<blocks_for((long long)g.N * g.Hi * g.Wi * g.Ci), 256, 0, st>>>(big, xh, xl, g.N, g.Ci, g.Hi, g.Wi);
  k_split2_nhwc<<<blocks_for((long long)g.N * g.Ho * g.Wo * g.Co), 256, 0, st>>>(small, zh, zl, g.N, g.Co, g.Ho, g.Wo);
  count_launch(2);
  SRB_CHECK_CUDA(cudaGetLastError());
  const T4 txh = dense_nhwc(xh, g.Ci, g.Hi, g.Wi), txl = dense_nhwc(xl, g.Ci, g.Hi, g.Wi);
  const T4 tzh = dense_nhwc(zh, g.Co, g.Ho, g.Wo), tzl = dense_nhwc(zl, g.Co, g.Ho, g.Wo);
  int rc = tc_conv_wgrad(g, tzh, txh, dw, db_small, scale, accumulate, ws_tc, ws_tc_bytes, st);
  if (rc) return rc;
  rc = tc_conv_wgrad(g, tzh, txl, dw, nullptr, scale, 1, ws_tc, ws_tc_bytes, st);
  if (rc) return rc;
  return tc_conv_wgrad(g, tzl, txh, dw, db_small, scale, 1, ws_tc, ws_tc_bytes, st);
}

}  // namespace srb
