// Shared declarations for libsrb200 (sm_100a).  Internal; the public surface is include/srb200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include "srb200.h"

namespace srb {

// Thread-local error string (srb_last_error) and the process-wide launch counter.
void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define SRB_CHECK_CUDA(expr)                                                                   \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      srb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__));    \
      return SRB_ECUDA;                                                                        \
    }                                                                                          \
  } while (0)

#define SRB_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      srb::set_error(__VA_ARGS__);    \
      return (code);                  \
    }                                 \
  } while (0)

// Device-side strided NCHW view.  Strides are in ELEMENTS; `dt` is the element type (SRB_F32: p is what it says;
// SRB_BF16: p really points at __nv_bfloat16 data -- only the kernels that declare bf16 support look at dt).
struct T4 {
  float *p;
  long long sn, sc, sh, sw;
  int dt;
};
static inline T4 to_t4(const srb_tensor4 *t) {
  T4 r;
  if (t) { r.p = (float *)t->data; r.sn = t->sn; r.sc = t->sc; r.sh = t->sh; r.sw = t->sw; r.dt = t->dtype; }
  else   { r.p = nullptr; r.sn = r.sc = r.sh = r.sw = 0; r.dt = SRB_F32; }
  return r;
}
static inline bool is_bf16(const T4 &t) { return t.p != nullptr && t.dt == SRB_BF16; }

// Geometry of one "gather" convolution:  small[n,co,oy,ox] <-> big[n,ci,oy*st-pad+r,ox*st-pad+s].
//   Conv2d:           big = x (Ci=Cin, Hi=H, Wi=W),  small = conv output (Co=Cout*ps*ps, Ho, Wo)
//   ConvTranspose2d:  big = y (Ci=Cout, Hi=Ho_t, Wi=Wo_t), small = x (Co=Cin, Ho=H, Wo=W)
// Weight element for (co,ci,r,s) sits at w[((co*Ci + ci)*kh + r)*kw + s] in both cases
// (Conv2d OIHW; ConvTranspose2d (Cin,Cout,kh,kw) with co=Cin index, ci=Cout index).
struct Geom {
  int N, Ci, Hi, Wi;
  int Co, Ho, Wo;
  int kh, kw, st, pad;
  int ps;  // PixelShuffle factor applied to the *small* side tensor addressing (Conv2d only)
};

struct Epi {
  const float *bias;   // indexed by channel of the written tensor (pre-shuffle channel), or null
  const float *alpha;  // PReLU slope (device scalar) or null
  int act;
  float slope;
  T4 residual;  // p == null -> none
  T4 preact;    // p == null -> none
  T4 mask;      // p == null -> none; else y = mask > 0 ? y : 0 applied last (ReLU backward of the producer of the written tensor)
  // Packed ReLU sign pattern of the WRITTEN tensor, 16 channels per uint16: word ((n*Ho+oy)*Wo+ox)*(Co/16) + c/16, bit c%16
  // (tensor path only, Co % 16 == 0, no PixelShuffle).  bits_out: set where the pre-activation is > 0 (fprop of a ReLU
  // layer); bits_in: the written value is zeroed where the bit is clear (dgrad of the layer that consumes that ReLU output).
  unsigned short *bits_out;
  const unsigned short *bits_in;
  int round_tf32;  // store y RN-rounded to tf32 (feeds a tensor-core consumer)
  // Regression loss fused into the epilogue of the network's last conv (espcn.py:129, srcnn.py:129, edsr.py:153):
  //   loss_kind 1 = MSE, 2 = L1 (mean over all elements); target has y's shape; every epilogue warp leaves the sum of its
  //   (y-t)^2 / |y-t| terms in loss_part[cta * 8 + warp] (folded in a fixed order afterwards: deterministic);
  //   the gradient g = loss_coef * 2 (y - t) or loss_coef * sign(y - t) goes either to dz_unshuf -- the conv's own
  //   (N, Ho, Wo, Co) NHWC layout, i.e. PixelShuffle already undone (PS(4) -> NCHW layers) -- or to dz (y's layout, ps == 1).
  //   out.p may be null: the loss is then the only product of the forward pass.
  int loss_kind;
  float loss_coef;
  T4 target;
  T4 dz;
  float *dz_unshuf;
  float *loss_part;
};

#ifdef __CUDACC__
// byte / 255 as torchvision's ToTensor computes it (`img.float().div(255)`, dataset.py:90): the correctly rounded quotient from one
// multiply by RN(1/255) and two FMAs (remainder, correction) -- bit-identical to the IEEE division for all 256 byte values
// (checked exhaustively, tests/test_oracle.py), while b * (1/255) alone is 1 ulp off for 126 of them.
__device__ __forceinline__ float byte_over_255(unsigned int b) {
  const float r = 1.0f / 255.0f, bf = (float)b;
  const float q0 = bf * r;
  const float rem = __fmaf_rn(-q0, 255.0f, bf);
  return __fmaf_rn(rem, r, q0);
}
#endif

// Programmatic dependent launch (PDL).  Kernels on the hot path call pdl_trigger() first thing -- the NEXT kernel of the stream
// may then be scheduled as soon as every CTA of this one has started, so its launch latency and prologue (barrier init, TMEM
// allocation, shared-memory zeroing, tensor-map prefetch) overlap this kernel's execution -- and pdl_wait() before their first
// access to global memory that an earlier kernel may still be writing or reading (it returns when all prerequisite grids have
// completed and flushed; a no-op for kernels launched without the attribute).  Captured into CUDA graphs as programmatic edges.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// SRB_NO_PDL=1 in the environment launches everything with plain stream serialisation (A/B measurements, debugging)
inline bool pdl_enabled() {
  static const bool on = [] { const char *e = getenv("SRB_NO_PDL"); return !(e && e[0] == '1'); }();
  return on;
}

// Per-call switch: the conv / wgrad entry points turn the attribute off for big layers (measured on one box: ESPCN cfg2, 4-16 GFLOP
// per launch, +1.5 %; EDSR-256 bf16, 39-620 GFLOP per launch, -2.7 %: nothing to hide there, and the early-resident dependents cost).
inline bool &pdl_scope_allowed() {
  static thread_local bool allowed = true;
  return allowed;
}
struct PdlScope {
  bool prev;
  explicit PdlScope(bool allow) : prev(pdl_scope_allowed()) { pdl_scope_allowed() = allow; }
  ~PdlScope() { pdl_scope_allowed() = prev; }
};

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_enabled() && pdl_scope_allowed()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ float round_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__device__ __forceinline__ float apply_act(float z, int act, float slope) {
  // ATen semantics: relu(z)=max(z,0); prelu/lrelu take the slope branch at z<=0 (value identical at 0).
  if (act == SRB_ACT_NONE) return z;
  if (act == SRB_ACT_RELU) return z > 0.f ? z : 0.f;
  return z > 0.f ? z : z * slope;
}

// Address of logical (n, k, oy, ox) of a conv output whose memory holds PixelShuffle_r of it:
//   out[n, c, oy*r+i, ox*r+j] = conv[n, c*r*r + i*r + j, oy, ox]      (base_networks.py:157; ATen pixel_shuffle)
__device__ __forceinline__ long long ps_offset(const T4 &t, int r, int n, int k, int oy, int ox) {
  if (r == 1) return n * t.sn + k * t.sc + oy * t.sh + ox * t.sw;
  int rr = r * r;
  int c = k / rr, ij = k - c * rr;
  int i = ij / r, j = ij - i * r;
  return n * t.sn + c * t.sc + (long long)(oy * r + i) * t.sh + (long long)(ox * r + j) * t.sw;
}

// Entry points implemented in the .cu files (host side).
int simt_conv_gather(const Geom &g, const T4 &in, const float *w, const T4 &out, const Epi &epi, cudaStream_t st);
int simt_conv_scatter(const Geom &g, const T4 &in_small, const float *w, const T4 &out_big, const Epi &epi,
                      cudaStream_t st);
size_t simt_wgrad_ws_bytes(const Geom &g);
int simt_conv_wgrad(const Geom &g, const T4 &small, const T4 &big, float *dw, float *db_small, float scale,
                    int accumulate, void *ws, size_t ws_bytes, cudaStream_t st);
int channel_sum(const T4 &t, int N, int C, int H, int W, float *out, float scale, int accumulate, cudaStream_t st);

// Tensor-core (tcgen05) path, tc_conv.cu.
bool tc_conv_supported(const Geom &g, const T4 &in, const T4 &out, bool scatter_as_gather, int in_ps = 1, int pad_w = -1);
size_t tc_conv_ws_bytes(const Geom &g);
// Options of the slot-linear conv beyond the plain stride-1 gather.
struct ConvOpt {
  // in_ps > 1: `in` holds the logical input in r x r INTERLEAVED form -- logical channel (ij, c), pixel (y, x) is
  // in[c, y*r + ij/r, x*r + ij%r] (g.Ci = C * r * r logical channels, `in` has C) -- and is fetched through a stride-r TMA
  // traversal.  Two uses: dz of a PixelShuffle layer (wmode 0: logical channel c*r*r + ij of the filter), and the r x r
  // phase images of a STRIDED convolution's input (wmode 1).
  int in_ps = 1;
  int in_h = 0, in_w = 0;  // true spatial size of `in` when it is not g.Hi*in_ps x g.Wi*in_ps (strided convs with odd sizes)
  int pad_w = -1;          // horizontal padding when it differs from g.pad (-1: same)
  // How launch filter element (n, k, tap (tr, ts)) maps to the layer's filter w[(o * I + i) * kh0 * kw0 + r * kw0 + s]:
  //   wmode 0: (r, s) = (tr, ts)                       [flip_transpose: (kh-1-tr, kw-1-ts) and o/i swapped]
  //   wmode 1: strided-input phases: k = ph * C + c, ph = a * st + b: r = st * (tr + dmin_r) + a + pad0 (valid if 0 <= r < kh0),
  //            same for s; invalid (phase, tap) pairs are skipped by the kernel (phase tap masks) -- o = n, i = c
  //   wmode 2: one OUTPUT phase of a transposed / strided-backward conv: r = ra + st * (tmax_a - tr), s = rb + st * (tmax_b - ts);
  //            o = k, i = n (the launch's output channel is the filter's second index)
  int wmode = 0;
  int st = 1, pad0 = 0, kh0 = 0, kw0 = 0;
  int dmin_r = 0, dmin_s = 0;
  int ra = 0, rb = 0, tmax_a = 0, tmax_b = 0;
  float *loss_out = nullptr;  // fused loss (Epi::loss_kind): receives the mean loss
};
int tc_conv_gather(const Geom &g, const T4 &in, const float *w, bool flip_transpose, const T4 &out, const Epi &epi,
                   void *ws, size_t ws_bytes, cudaStream_t st, const ConvOpt &opt = ConvOpt());
// Strided Conv2d forward / ConvTranspose2d backward-data (input phases) and their adjoints (output phases) on the tensor path
bool tc_strided_gather_supported(const Geom &g, const T4 &big, const T4 &small);
int tc_strided_gather(const Geom &g, const T4 &big, const float *w, const T4 &small, const Epi &epi, void *ws, size_t ws_bytes,
                      cudaStream_t st);
bool tc_strided_scatter_supported(const Geom &g, const T4 &small, const T4 &big);
int tc_strided_scatter(const Geom &g, const T4 &small, const float *w, const T4 &big, const Epi &epi, void *ws, size_t ws_bytes,
                       cudaStream_t st);
size_t tc_strided_ws_bytes(const Geom &g);
// Packed-weight cache (tc_conv_sl.cu): only filters handed in by the user (not the temporaries of the 3xTF32 mode) may be cached
inline bool &weight_cache_scope_allowed() {
  static thread_local bool allowed = false;
  return allowed;
}
struct WeightCacheScope {
  bool prev;
  explicit WeightCacheScope(bool allow) : prev(weight_cache_scope_allowed()) { weight_cache_scope_allowed() = allow; }
  ~WeightCacheScope() { weight_cache_scope_allowed() = prev; }
};
int tc_weight_cache_enable(int on);
int tc_weight_cache_repack(cudaStream_t st);
int tc_weight_cache_entries();
void tc_conv_set_trace(long long *buf, long long max_ctas);
void tc_conv_set_dbg(int flags);
int tc_conv_get_dbg();
int tc_conv_describe(const Geom &g, char *buf, size_t n, bool bf16 = false);
int tc_wgrad_describe(const Geom &g, char *buf, size_t n, bool bf16 = false);
bool tc_wgrad_supported(const Geom &g, const T4 &small, const T4 &big, int z_ps = 1);
size_t tc_wgrad_ws_bytes(const Geom &g, bool bf16 = false, int z_ps = 1);
// z_ps > 1: `small` holds PixelShuffle_r of dz (in y's layout); un-shuffled by the TMA traversal, dw/db come out in filter order
// One input phase of a STRIDED convolution's weight gradient (strided.cu): the launch is a stride-1 wgrad between dz and the phase
// image x[st*i + a, st*j + b]; its tap (tr, ts) is filter tap r = st * (tr - pad') + a + pad0, s = st * (ts - pad') + b + pad0 of the
// kh0 x kw0 filter (taps outside the filter are computed and dropped).
struct WgPhase {
  int st, a, b, pad0, kh0, kw0;
};
int tc_conv_wgrad(const Geom &g, const T4 &small, const T4 &big, float *dw, float *db_small, float scale,
                  int accumulate, void *ws, size_t ws_bytes, cudaStream_t st, int z_ps = 1, const WgPhase *phase = nullptr);
bool tc_strided_wgrad_supported(const Geom &g, const T4 &small, const T4 &big);
size_t tc_strided_wgrad_ws_bytes(const Geom &g);
int tc_strided_wgrad(const Geom &g, const T4 &small, const T4 &big, float *dw, float *db_small, float scale, int accumulate,
                     void *ws, size_t ws_bytes, cudaStream_t st);


// 3xTF32 split-operand mode (exact.cu): fp32-accurate results from the same tensor-core kernels.
bool exact_conv_supported(const Geom &g, const T4 &out);
size_t exact_conv_ws_bytes(const Geom &g);
int exact_conv_gather(const Geom &g, const T4 &in, const float *w, bool flip_transpose, const T4 &out, const Epi &epi,
                      void *ws, size_t ws_bytes, cudaStream_t st);
bool exact_wgrad_supported(const Geom &g);
size_t exact_wgrad_ws_bytes(const Geom &g);
int exact_conv_wgrad(const Geom &g, const T4 &small, const T4 &big, float *dw, float *db_small, float scale,
                     int accumulate, void *ws, size_t ws_bytes, cudaStream_t st);

}  // namespace srb
