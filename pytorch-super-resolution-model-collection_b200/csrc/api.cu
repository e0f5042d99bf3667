// extern "C" surface of libsrb200.so (see include/srb200.h) + the small elementwise kernels.
#include "srb_common.cuh"
#include <atomic>
#include <string.h>

namespace srb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

namespace {

// dz = dy * act'(ref);  PReLU: dalpha += sum(dy * z * [z<=0]) (ATen _prelu_kernel_backward semantics)
// Iterates in dz's memory order: channels_last -> (n,h,w,c), else (n,c,h,w).
__global__ void k_act_bwd(T4 dy, T4 ref, T4 dz, int N, int C, int H, int W, int act, float slope_in,
                          const float *__restrict__ alpha, float *dalpha, int cl, int rnd) {
  pdl_trigger();
  pdl_wait();
  const float slope = (act == SRB_ACT_PRELU) ? __ldg(alpha) : slope_in;
  long long total = (long long)N * C * H * W;
  float da = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int n, c, h, w;
    long long t = i;
    if (cl) { c = (int)(t % C); t /= C; w = (int)(t % W); t /= W; h = (int)(t % H); n = (int)(t / H); }
    else    { w = (int)(t % W); t /= W; h = (int)(t % H); t /= H; c = (int)(t % C); n = (int)(t / C); }
    float g = __ldg(dy.p + n * dy.sn + c * dy.sc + h * dy.sh + w * dy.sw);
    float r = __ldg(ref.p + n * ref.sn + c * ref.sc + h * ref.sh + w * ref.sw);
    float o;
    if (act == SRB_ACT_RELU) o = r > 0.f ? g : 0.f;
    else {
      o = r > 0.f ? g : g * slope;
      if (act == SRB_ACT_PRELU && !(r > 0.f)) da += g * r;
    }
    if (rnd) o = round_tf32(o);
    dz.p[n * dz.sn + c * dz.sc + h * dz.sh + w * dz.sw] = o;
  }
  if (act == SRB_ACT_PRELU && dalpha) {
    for (int o = 16; o > 0; o >>= 1) da += __shfl_xor_sync(0xffffffffu, da, o);
    __shared__ float red[32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = da;
    __syncthreads();
    if (threadIdx.x < 32) {
      da = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
      for (int o = 16; o > 0; o >>= 1) da += __shfl_xor_sync(0xffffffffu, da, o);
      if (threadIdx.x == 0) atomicAdd(dalpha, da);
    }
  }
}

// out[n, k, h, w] (NHWC) = dz[n, c, h*r+i, w*r+j] with k = c*r*r + i*r + j, rounded to tf32
__global__ void k_pixel_unshuffle(T4 dz, T4 out, int N, int K, int H, int W, int r, int rnd) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)N * H * W * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int k = (int)(i % K);
    long long t = i / K;
    int w = (int)(t % W); t /= W;
    int h = (int)(t % H);
    int n = (int)(t / H);
    float v = __ldg(dz.p + ps_offset(dz, r, n, k, h, w));
    out.p[n * out.sn + (long long)h * out.sh + (long long)w * out.sw + k] = rnd ? round_tf32(v) : v;
  }
}

// Tiled pixel_unshuffle for row-contiguous dz (sw == 1): one block per output row (n, h).  The C*r source rows
// (W*r floats each, fully coalesced reads) are staged in shared memory, then written as the W*K contiguous floats of the
// NHWC output row (fully coalesced stores).  Row pitch W*r + 4 keeps the (i, j) gather off a single bank.
__global__ void __launch_bounds__(256)
k_pixel_unshuffle_rows(T4 dz, T4 out, int N, int C, int H, int W, int r, int rnd) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float srow[];
  const int n = blockIdx.x / H, h = blockIdx.x - n * H;
  const int Wr = W * r, pitch = Wr + 4, rows = C * r, K = C * r * r;
  // load: rows x Wr floats
  if ((Wr & 3) == 0 && (dz.sh & 3) == 0 && (dz.sc & 3) == 0 && (dz.sn & 3) == 0 && ((uintptr_t)dz.p & 15) == 0) {
    const int w4 = Wr >> 2;
    for (int e = threadIdx.x; e < rows * w4; e += blockDim.x) {
      const int row = e / w4, x4 = e - row * w4;
      const int c = row / r, i = row - c * r;
      const float4 v = __ldg((const float4 *)(dz.p + n * dz.sn + c * dz.sc + (long long)(h * r + i) * dz.sh) + x4);
      *(float4 *)(srow + row * pitch + 4 * x4) = v;
    }
  } else {
    for (int e = threadIdx.x; e < rows * Wr; e += blockDim.x) {
      const int row = e / Wr, x = e - row * Wr;
      const int c = row / r, i = row - c * r;
      srow[row * pitch + x] = __ldg(dz.p + n * dz.sn + c * dz.sc + (long long)(h * r + i) * dz.sh + x);
    }
  }
  __syncthreads();
  float *orow = out.p + n * out.sn + (long long)h * out.sh;
  if (r == 4 && out.sw == K) {
    // one float4 = the 4 j's of (w, c, i): smem s[c*4+i][4w .. 4w+3] -> out[w*K + c*16 + i*4 ..]
    const int q = K >> 2;  // float4 per pixel
    for (int e = threadIdx.x; e < W * q; e += blockDim.x) {
      const int w = e / q, ci = e - w * q;  // ci = c*4 + i
      float4 v = *(const float4 *)(srow + ci * pitch + 4 * w);
      if (rnd) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
      *(float4 *)(orow + (long long)w * K + 4 * ci) = v;
    }
  } else {
    const int rr = r * r;
    for (int e = threadIdx.x; e < W * K; e += blockDim.x) {
      const int w = e / K, k = e - w * K;
      const int c = k / rr, ij = k - c * rr, i = ij / r, j = ij - i * r;
      const float v = srow[(c * r + i) * pitch + w * r + j];
      orow[(long long)w * out.sw + k] = rnd ? round_tf32(v) : v;
    }
  }
}

// Same-layout dense tensors: flat float4 walk (the common case: dy, ref, dz all NHWC- or all NCHW-contiguous)
__global__ void k_act_bwd_flat(const float4 *__restrict__ dy, const float4 *__restrict__ ref, float4 *dz, long long n4,
                               int act, float slope_in, const float *__restrict__ alpha, float *dalpha, int rnd) {
  pdl_trigger();
  pdl_wait();
  const float slope = (act == SRB_ACT_PRELU) ? __ldg(alpha) : (act == SRB_ACT_RELU ? 0.f : slope_in);
  float da = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g = __ldg(dy + i), r = __ldg(ref + i);
    float4 o;
    o.x = r.x > 0.f ? g.x : g.x * slope; o.y = r.y > 0.f ? g.y : g.y * slope;
    o.z = r.z > 0.f ? g.z : g.z * slope; o.w = r.w > 0.f ? g.w : g.w * slope;
    if (act == SRB_ACT_PRELU) {
      da += (r.x > 0.f ? 0.f : g.x * r.x) + (r.y > 0.f ? 0.f : g.y * r.y) + (r.z > 0.f ? 0.f : g.z * r.z) +
            (r.w > 0.f ? 0.f : g.w * r.w);
    }
    if (rnd) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
    dz[i] = o;
  }
  if (act == SRB_ACT_PRELU && dalpha) {
    for (int o = 16; o > 0; o >>= 1) da += __shfl_xor_sync(0xffffffffu, da, o);
    __shared__ float red[32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = da;
    __syncthreads();
    if (threadIdx.x < 32) {
      da = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
      for (int o = 16; o > 0; o >>= 1) da += __shfl_xor_sync(0xffffffffu, da, o);
      if (threadIdx.x == 0) atomicAdd(dalpha, da);
    }
  }
}

__global__ void k_prelu_fwd(const float *__restrict__ x, const float *__restrict__ alpha, float *y, long long n) {
  pdl_trigger();
  pdl_wait();
  const float a = __ldg(alpha);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = x[i];
    y[i] = v > 0.f ? v : v * a;
  }
}

__global__ void k_prelu_bwd(const float *__restrict__ x, const float *__restrict__ dy, const float *__restrict__ alpha,
                            float *dx, float *dalpha, long long n) {
  pdl_trigger();
  pdl_wait();
  const float a = __ldg(alpha);
  float da = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = x[i], g = dy[i];
    dx[i] = v > 0.f ? g : g * a;
    if (!(v > 0.f)) da += g * v;
  }
  for (int o = 16; o > 0; o >>= 1) da += __shfl_xor_sync(0xffffffffu, da, o);
  __shared__ float red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = da;
  __syncthreads();
  if (threadIdx.x < 32) {
    da = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) da += __shfl_xor_sync(0xffffffffu, da, o);
    if (threadIdx.x == 0 && dalpha) atomicAdd(dalpha, da);
  }
}

__global__ void k_round_tf32(const float *__restrict__ x, float *y, long long n) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = round_tf32(x[i]);
}

// ---- fused regression losses (srcnn.py:84 / edsr.py:98: nn.MSELoss / nn.L1Loss with mean reduction) -------------------
// forward: per-block partial sums of (y-t)^2 or |y-t| in a fixed order, a second tiny kernel folds them (deterministic);
// backward: dy = g * 2 (y - t) / n   or   g * sign(y - t) / n   in one pass (g = upstream gradient, a device scalar).
__global__ void __launch_bounds__(256) k_loss_partial(const float4 *__restrict__ y, const float4 *__restrict__ t, long long n4,
                                                      const float *__restrict__ ytail, const float *__restrict__ ttail, int ntail,
                                                      int l1, float *__restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = __ldg(y + i), b = __ldg(t + i);
    const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
    s += l1 ? (fabsf(d0) + fabsf(d1)) + (fabsf(d2) + fabsf(d3)) : (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < ntail) {
    const float d = __ldg(ytail + threadIdx.x) - __ldg(ttail + threadIdx.x);
    s += l1 ? fabsf(d) : d * d;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = red[0];
    for (int w = 1; w < 8; ++w) v += red[w];
    partial[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256) k_loss_finish(const float *__restrict__ partial, int nblocks, float inv_n, float *loss) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < nblocks; i += 256) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = red[0] * inv_n;
}

__global__ void __launch_bounds__(256) k_loss_bwd(const float *__restrict__ y, const float *__restrict__ t, long long n, int l1,
                                                  float inv_n, const float *__restrict__ g, float *__restrict__ dy) {
  pdl_trigger();
  pdl_wait();
  const float c = __ldg(g) * inv_n * (l1 ? 1.f : 2.f);
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = __ldg((const float4 *)y + i), b = __ldg((const float4 *)t + i);
    float4 o;
    if (l1) {  // ATen l1_loss backward: sign(y - t), 0 at equality
      o.x = a.x > b.x ? c : (a.x < b.x ? -c : 0.f); o.y = a.y > b.y ? c : (a.y < b.y ? -c : 0.f);
      o.z = a.z > b.z ? c : (a.z < b.z ? -c : 0.f); o.w = a.w > b.w ? c : (a.w < b.w ? -c : 0.f);
    } else {
      o.x = c * (a.x - b.x); o.y = c * (a.y - b.y); o.z = c * (a.z - b.z); o.w = c * (a.w - b.w);
    }
    ((float4 *)dy)[i] = o;
  }
  if (blockIdx.x == 0) {
    const long long i = (n4 << 2) + threadIdx.x;
    if (i < n) {
      const float a = __ldg(y + i), b = __ldg(t + i);
      dy[i] = l1 ? (a > b ? c : (a < b ? -c : 0.f)) : c * (a - b);
    }
  }
}

// dz (N, C<=3, H, W; any strides) -> NHWC4 (16 B per pixel, missing channels zero), tf32-rounded: lets the skinny output layers
// (64->3, 32->3) use the tensor-core wgrad with Co = 4
__global__ void k_pack_dz4(T4 dz, float4 *__restrict__ out, int N, int C, int H, int W) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)N * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    long long q = i / W;
    const int h = (int)(q % H);
    const int n = (int)(q / H);
    const float *p = dz.p + n * dz.sn + (long long)h * dz.sh + (long long)w * dz.sw;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < C) v[c] = round_tf32(__ldg(p + c * dz.sc));
    out[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// torchvision ToTensor (dataset.py:90,94,98 of the reference): uint8 HWC image in [0,255] -> float CHW in [0,1]
__global__ void k_image_to_tensor(const unsigned char *__restrict__ src, float *__restrict__ dst, int N, int H, int W, int C,
                                  float scale) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)N * C * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    long long q = i / W;
    const int h = (int)(q % H); q /= H;
    const int c = (int)(q % C);
    const long long n = q / C;
    const unsigned int b = __ldg(src + ((n * H + h) * W + w) * C + c);
    dst[i] = scale == (1.0f / 255.0f) ? byte_over_255(b) : (float)b * scale;  // 1/255: ToTensor's exact division
  }
}

// The same for C <= 4 channels, W % 4 == 0 and 4-byte aligned buffers: a thread takes FOUR consecutive pixels of one row -- 4*C
// bytes = C aligned 32-bit loads -- and writes one float4 into each channel plane (the scalar kernel above does a strided byte load,
// three 64-bit divisions and a 4-byte store per element: 75 us of every ESPCN e2e step when it ran beside the training kernels).
template <int C>
__global__ void __launch_bounds__(256) k_image_to_tensor_v4(const uint32_t *__restrict__ src, float4 *__restrict__ dst, int N, int H, int W4,
                                                            float scale) {
  pdl_trigger();
  pdl_wait();
  const long long groups = (long long)N * H * W4;  // groups of 4 pixels
  const long long plane4 = (long long)H * W4;      // float4s per channel plane
  for (long long gidx = (long long)blockIdx.x * blockDim.x + threadIdx.x; gidx < groups; gidx += (long long)gridDim.x * blockDim.x) {
    const long long n = gidx / plane4, hw4 = gidx - n * plane4;
    uint32_t wd[C];
#pragma unroll
    for (int j = 0; j < C; ++j) wd[j] = __ldg(src + gidx * C + j);  // 4 pixels x C bytes, pixel-major
    const bool exact = scale == (1.0f / 255.0f);  // ToTensor: the correctly rounded byte / 255
    float v[C][4];
#pragma unroll
    for (int px = 0; px < 4; ++px)
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int b = px * C + c;
        const unsigned int byte = (wd[b >> 2] >> ((b & 3) * 8)) & 0xffu;
        v[c][px] = exact ? byte_over_255(byte) : (float)byte * scale;
      }
#pragma unroll
    for (int c = 0; c < C; ++c) dst[(n * C + c) * plane4 + hw4] = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
  }
}

// ---- bf16 storage mode helpers ------------------------------------------------------------------------------------------
__device__ __forceinline__ float h2f(unsigned short h) { return __uint_as_float((uint32_t)h << 16); }
__device__ __forceinline__ unsigned short f2h(float v) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(0.f), "f"(v));
  return (unsigned short)(r & 0xffffu);
}

// dz = dy * act'(ref) over n contiguous bf16 elements (same dense layout on all three); PReLU: dalpha += sum(dy*z*[z<=0])
__global__ void k_act_bwd_flat_h(const unsigned short *__restrict__ dy, const unsigned short *__restrict__ ref,
                                 unsigned short *dz, long long n, int act, float slope_in, const float *__restrict__ alpha,
                                 float *dalpha) {
  pdl_trigger();
  pdl_wait();
  const float slope = (act == SRB_ACT_PRELU) ? __ldg(alpha) : (act == SRB_ACT_RELU ? 0.f : slope_in);
  float da = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = h2f(__ldg(dy + i)), r = h2f(__ldg(ref + i));
    dz[i] = f2h(r > 0.f ? g : g * slope);
    if (act == SRB_ACT_PRELU && !(r > 0.f)) da += g * r;
  }
  if (act == SRB_ACT_PRELU && dalpha) {
    for (int o = 16; o > 0; o >>= 1) da += __shfl_xor_sync(0xffffffffu, da, o);
    __shared__ float red[32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = da;
    __syncthreads();
    if (threadIdx.x < 32) {
      da = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
      for (int o = 16; o > 0; o >>= 1) da += __shfl_xor_sync(0xffffffffu, da, o);
      if (threadIdx.x == 0) atomicAdd(dalpha, da);
    }
  }
}

// pixel_unshuffle, bf16 NHWC -> bf16 NHWC: out[n,h,w,k = c*r*r + i*r + j] = dz[n, h*r+i, w*r+j, c]; one thread per
// (pixel, sub-pixel, 8-channel group): one 16-byte load, eight 2-byte stores
__global__ void k_pixel_unshuffle_h(T4 dz, T4 out, int N, int C, int H, int W, int r) {
  pdl_trigger();
  pdl_wait();
  const int rr = r * r, cg = C >> 3;
  const long long total = (long long)N * H * W * rr * cg;
  const unsigned short *src = (const unsigned short *)dz.p;
  unsigned short *dst = (unsigned short *)out.p;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g8 = (int)(i % cg);
    long long q = i / cg;
    const int ij = (int)(q % rr); q /= rr;
    const int w = (int)(q % W); q /= W;
    const int h = (int)(q % H);
    const int n = (int)(q / H);
    const int ii = ij / r, jj = ij - ii * r;
    const uint4 v = __ldg((const uint4 *)(src + n * dz.sn + (long long)(h * r + ii) * dz.sh + (long long)(w * r + jj) * dz.sw + g8 * 8));
    unsigned short *o = dst + n * out.sn + (long long)h * out.sh + (long long)w * out.sw + (long long)(g8 * 8) * rr + ij;
    o[0 * rr] = (unsigned short)(v.x & 0xffffu); o[1 * rr] = (unsigned short)(v.x >> 16);
    o[2 * rr] = (unsigned short)(v.y & 0xffffu); o[3 * rr] = (unsigned short)(v.y >> 16);
    o[4 * rr] = (unsigned short)(v.z & 0xffffu); o[5 * rr] = (unsigned short)(v.z >> 16);
    o[6 * rr] = (unsigned short)(v.w & 0xffffu); o[7 * rr] = (unsigned short)(v.w >> 16);
  }
}

// bf16 NHWC (N,C,H,W logical) -> dense fp32 NHWC (tf32-representable by construction: bf16 has 8 mantissa bits)
__global__ void k_h2f_nhwc(T4 x, float *__restrict__ out, int N, int C, int H, int W) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)N * H * W * C;
  const unsigned short *src = (const unsigned short *)x.p;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long q = i / C;
    const int w = (int)(q % W); q /= W;
    const int h = (int)(q % H);
    const int n = (int)(q / H);
    out[i] = h2f(__ldg(src + n * x.sn + c * x.sc + (long long)h * x.sh + (long long)w * x.sw));
  }
}

// dz (N, C<=8, H, W; fp32, any strides) -> bf16 NHWC8 (16 B per pixel, missing channels zero)
__global__ void k_pack_dz8_h(T4 dz, uint4 *__restrict__ out, int N, int C, int H, int W) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)N * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    long long q = i / W;
    const int h = (int)(q % H);
    const int n = (int)(q / H);
    const float *p = dz.p + n * dz.sn + (long long)h * dz.sh + (long long)w * dz.sw;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = c < C ? __ldg(p + c * dz.sc) : 0.f;
    uint4 o;
    o.x = (uint32_t)f2h(v[0]) | ((uint32_t)f2h(v[1]) << 16); o.y = (uint32_t)f2h(v[2]) | ((uint32_t)f2h(v[3]) << 16);
    o.z = (uint32_t)f2h(v[4]) | ((uint32_t)f2h(v[5]) << 16); o.w = (uint32_t)f2h(v[6]) | ((uint32_t)f2h(v[7]) << 16);
    out[i] = o;
  }
}

// x *= *g unless *g == 1 (the usual upstream gradient of a scalar loss): every block reads g and leaves early
__global__ void k_scale_by_scalar(float *x, long long n, const float *__restrict__ g, int rnd) {
  pdl_trigger();
  pdl_wait();
  const float s = __ldg(g);
  if (s == 1.0f) return;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i] * s;
    x[i] = rnd ? round_tf32(v) : v;
  }
}

// utils.img_interp(imgs, scale, 'bicubic') (utils.py:242-269): per image ToPILImage -> PIL resize(BICUBIC) -> ToTensor, i.e.
// Pillow's 8-bit two-pass resample (horizontal first, 22-bit fixed-point coefficients, 8-bit intermediate), optionally
// followed by utils.shave (utils.py:197-205).  One thread per output pixel; bit-exact integer arithmetic.
__global__ void k_pil_bicubic(const float *__restrict__ x, float *__restrict__ y, int N, int C, int H, int W, int TH, int TW,
                              const int *__restrict__ bw, const int *__restrict__ kw, const int *__restrict__ bh,
                              const int *__restrict__ kh, int ksize, int shave) {
  pdl_trigger();
  pdl_wait();
  const int OH = TH - 2 * shave, OW = TW - 2 * shave;
  const long long total = (long long)N * C * OH * OW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % OW) + shave;
    long long q = i / OW;
    const int oy = (int)(q % OH) + shave;
    const long long nc = q / OH;
    const float *img = x + nc * H * W;
    const int x0 = __ldg(bw + 2 * ox), nx = __ldg(bw + 2 * ox + 1), y0 = __ldg(bh + 2 * oy), ny = __ldg(bh + 2 * oy + 1);
    long long acc = 1LL << 21;
    for (int yy = 0; yy < ny; ++yy) {
      const float *row = img + (long long)(y0 + yy) * W + x0;
      long long h = 1LL << 21;
      for (int xx = 0; xx < nx; ++xx) {
        const int u8 = (int)(unsigned char)(int)(__ldg(row + xx) * 255.0f);  // ToPILImage: mul(255).byte() (truncation)
        h += (long long)u8 * __ldg(kw + ox * ksize + xx);
      }
      long long t = h >> 22;
      t = t < 0 ? 0 : (t > 255 ? 255 : t);
      acc += t * __ldg(kh + oy * ksize + yy);
    }
    long long r = acc >> 22;
    r = r < 0 ? 0 : (r > 255 ? 255 : r);
    y[i] = (float)r / 255.0f;  // ToTensor
  }
}

// torch.optim.Adam (espcn.py:79, edsr.py:93; amsgrad off) over FLAT buffers: every parameter, gradient and moment of the model is a
// slice of one fp32 array (srb200.GradBucket / srb200.FlatAdam), so the whole optimizer is this one elementwise launch.
// Same arithmetic and order as torch's fused implementation (fused_adam_utils.cuh adam_math, ADAM mode):
//   g += wd * p;  m = m + (g - m) * (1 - b1);  v = b2 * v + (1 - b2) * g * g;  p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// state[0] = step count t (float, on the device: CUDA-graph replays advance it), state[1] = blocks finished (last block commits t+1).
__global__ void __launch_bounds__(256) k_adam_flat(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                                   float *__restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd,
                                                   float *state) {
  pdl_trigger();
  pdl_wait();
  const float t = state[0] + 1.f;
  const float bc1 = 1.f - powf(b1, t), bc2_sqrt = sqrtf(1.f - powf(b2, t));
  const float step_size = lr / bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i];
    const float pi = p[i];
    if (wd != 0.f) gi += pi * wd;
    const float mi = m[i] + (gi - m[i]) * (1.f - b1);
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const float done = atomicAdd(state + 1, 1.f);
    if (done == (float)(gridDim.x - 1)) {  // every block has read state[0] (it read it before arriving here)
      state[0] = t;
      state[1] = 0.f;
    }
  }
}

inline unsigned ew_blocks(long long n) {
  long long b = (n + 255) / 256;
  if (b > 148LL * 16) b = 148LL * 16;
  if (b < 1) b = 1;
  return (unsigned)b;
}

int check_params(const srb_conv_params *p) {
  SRB_REQUIRE(p != nullptr, SRB_EINVAL, "null params");
  SRB_REQUIRE(p->N >= 0 && p->Cin > 0 && p->Cout > 0 && p->H > 0 && p->W > 0, SRB_EINVAL, "bad tensor dims");
  SRB_REQUIRE(p->kh > 0 && p->kw > 0 && p->stride > 0 && p->pad >= 0 && p->out_pad >= 0, SRB_EINVAL,
              "bad kernel/stride/pad");
  SRB_REQUIRE(p->ps >= 1, SRB_EINVAL, "ps must be >= 1");
  SRB_REQUIRE(p->act >= SRB_ACT_NONE && p->act <= SRB_ACT_LRELU, SRB_EINVAL, "bad activation %d", p->act);
  SRB_REQUIRE(p->math >= SRB_MATH_FP32 && p->math <= SRB_MATH_BF16, SRB_EINVAL, "bad math mode %d", p->math);
  SRB_REQUIRE(!(p->transposed && p->ps != 1), SRB_EUNSUPPORTED, "PixelShuffle fused with ConvTranspose2d");
  SRB_REQUIRE(p->transposed || p->out_pad == 0, SRB_EINVAL, "out_pad only for transposed");
  SRB_REQUIRE(!p->transposed || p->out_pad < p->stride, SRB_EINVAL, "out_pad must be < stride");
  return SRB_OK;
}

// Geometry in "gather" orientation (see srb_common.cuh).
int make_geom(const srb_conv_params *p, Geom *g) {
  int rc = check_params(p);
  if (rc) return rc;
  int Ho, Wo;
  if (!p->transposed) {
    Ho = (p->H + 2 * p->pad - p->kh) / p->stride + 1;
    Wo = (p->W + 2 * p->pad - p->kw) / p->stride + 1;
    SRB_REQUIRE(p->H + 2 * p->pad >= p->kh && p->W + 2 * p->pad >= p->kw, SRB_EINVAL, "kernel larger than padded input");
    *g = Geom{p->N, p->Cin, p->H, p->W, p->Cout * p->ps * p->ps, Ho, Wo, p->kh, p->kw, p->stride, p->pad, p->ps};
  } else {
    Ho = (p->H - 1) * p->stride - 2 * p->pad + p->kh + p->out_pad;
    Wo = (p->W - 1) * p->stride - 2 * p->pad + p->kw + p->out_pad;
    SRB_REQUIRE(Ho > 0 && Wo > 0, SRB_EINVAL, "transposed conv output empty");
    // big = y (Ci := Cout), small = x (Co := Cin)
    *g = Geom{p->N, p->Cout, Ho, Wo, p->Cin, p->H, p->W, p->kh, p->kw, p->stride, p->pad, 1};
  }
  return SRB_OK;
}

Epi make_epi(const srb_conv_params *p, const float *bias, const float *alpha, const srb_tensor4 *residual,
             const srb_tensor4 *preact, int round_out) {
  Epi e;
  e.bias = bias;
  e.alpha = alpha;
  e.act = p->act;
  e.slope = p->slope;
  e.residual = to_t4(residual);
  e.preact = to_t4(preact);
  e.mask = to_t4(nullptr);
  e.bits_out = nullptr;
  e.bits_in = nullptr;
  e.round_tf32 = round_out;
  e.loss_kind = 0;
  e.loss_coef = 0.f;
  e.target = to_t4(nullptr);
  e.dz = to_t4(nullptr);
  e.dz_unshuf = nullptr;
  e.loss_part = nullptr;
  return e;
}

inline bool is_cl(const T4 &t, int C) { return t.sc == 1 && t.sw == C; }

// single-pass tf32 operands (activations are stored tf32-rounded between layers)?  EXACT keeps full fp32 activations.
inline bool is_tf32_math(int math) { return math == SRB_MATH_TF32 || math == SRB_MATH_AUTO; }

// bf16 storage mode: an activation with C channels is a bf16 tensor iff its pixel rows are 16-byte multiples (C % 8 == 0);
// the 3-channel network edges stay fp32
inline int bf16_act_dtype(int C) { return (C % 8 == 0 && C >= 8) ? SRB_BF16 : SRB_F32; }
inline size_t a256(size_t v) { return (v + 255) & ~(size_t)255; }

// Does the *written* tensor feed tensor-core consumers?  (channels_last, C % 4 == 0, C >= 8)
inline int want_round(const srb_conv_params *p, const T4 &t, int C) {
  return (is_tf32_math(p->math) && is_cl(t, C) && (C % 4) == 0 && C >= 8) ? 1 : 0;
}

// Skinny-output wgrad on the tensor path: geometry with Co padded to 4 and the NHWC4 view of the packed dz.
// Workspace layout: [packed dz][dw for 4 output channels][db for 4][tensor-core wgrad workspace].
struct SkinnyWg {
  Geom g4;
  T4 dz4;
  size_t pack_bytes, dw_bytes, total;
};
inline bool skinny_wgrad_plan(const srb_conv_params *p, const Geom &g, const T4 &big, SkinnyWg *sk) {
  if (p->transposed || !is_tf32_math(p->math) || g.ps != 1 || g.st != 1 || g.Co >= 4) return false;
  sk->g4 = g;
  sk->g4.Co = 4;
  sk->dz4 = T4{(float *)256, (long long)g.Ho * g.Wo * 4, 1, (long long)g.Wo * 4, 4};  // pointer patched by the caller
  if (!tc_wgrad_supported(sk->g4, sk->dz4, big)) return false;
  sk->pack_bytes = (((size_t)g.N * g.Ho * g.Wo * 4 * sizeof(float)) + 255) & ~(size_t)255;
  sk->dw_bytes = (((size_t)4 * g.Ci * g.kh * g.kw + 4) * sizeof(float) + 255) & ~(size_t)255;
  sk->total = sk->pack_bytes + sk->dw_bytes + tc_wgrad_ws_bytes(sk->g4) + 512;
  return true;
}

}  // namespace
}  // namespace srb

using namespace srb;

extern "C" {

int srb_version(void) { return SRB200_VERSION; }
const char *srb_last_error(void) { return g_err; }
int64_t srb_launch_count(void) { return (int64_t)g_launches.load(); }

int srb_conv_out_hw(const srb_conv_params *p, int32_t *Ho, int32_t *Wo) {
  Geom g;
  int rc = make_geom(p, &g);
  if (rc) return rc;
  if (!p->transposed) { *Ho = g.Ho; *Wo = g.Wo; }
  else                { *Ho = g.Hi; *Wo = g.Wi; }
  return SRB_OK;
}

int srb_conv_uses_tensor_path(const srb_conv_params *p, int pass, int x_cl, int y_cl) {
  Geom g;
  if (make_geom(p, &g)) return 0;
  if (p->math == SRB_MATH_FP32 || p->transposed) return 0;
  // fake dense views with the requested layouts
  T4 x{nullptr, 0, 0, 0, 0}, y{nullptr, 0, 0, 0, 0};
  if (x_cl) { x.sc = 1; x.sw = g.Ci; x.sh = (long long)g.Wi * g.Ci; x.sn = x.sh * g.Hi; }
  else      { x.sw = 1; x.sh = g.Wi; x.sc = (long long)g.Hi * g.Wi; x.sn = x.sc * g.Ci; }
  int Cy = p->Cout, Hy = g.Ho * g.ps, Wy = g.Wo * g.ps;
  if (y_cl) { y.sc = 1; y.sw = Cy; y.sh = (long long)Wy * Cy; y.sn = y.sh * Hy; }
  else      { y.sw = 1; y.sh = Wy; y.sc = (long long)Hy * Wy; y.sn = y.sc * Cy; }
  x.p = y.p = (float *)32;
  if (p->math == SRB_MATH_BF16) {
    x.dt = bf16_act_dtype(g.Ci);
    y.dt = bf16_act_dtype(Cy);
    if (pass == 0) return tc_conv_supported(g, x, y, false) ? 1 : 0;
    T4 ysmall = y;  // dz in conv-output geometry (after the un-shuffle when ps > 1)
    ysmall.dt = bf16_act_dtype(g.Co);
    if (pass == 1) {
      if (g.ps != 1 || g.st != 1 || g.kh != g.kw) return 0;
      Geom gd{g.N, g.Co, g.Ho, g.Wo, g.Ci, g.Hi, g.Wi, g.kh, g.kw, 1, g.kh - 1 - g.pad, 1};
      return (gd.pad >= 0 && tc_conv_supported(gd, ysmall, x, true)) ? 1 : 0;
    }
    return 1;  // wgrad: always a tensor-core plan in this mode (mixed edges are converted first)
  }
  const bool exact = p->math == SRB_MATH_EXACT;
  if (pass == 0) return (exact ? exact_conv_supported(g, y) : tc_conv_supported(g, x, y, false)) ? 1 : 0;
  if (pass == 1) {
    if (g.ps != 1 || g.st != 1 || g.kh != g.kw) return 0;
    Geom gd{g.N, g.Co, g.Ho, g.Wo, g.Ci, g.Hi, g.Wi, g.kh, g.kw, 1, g.kh - 1 - g.pad, 1};
    if (gd.pad < 0) return 0;  // pad > k-1: srb_conv_dgrad takes the CUDA-core scatter kernel
    return (exact ? exact_conv_supported(gd, x) : tc_conv_supported(gd, y, x, true)) ? 1 : 0;
  }
  if (!exact && g.st > 1) return tc_strided_wgrad_supported(g, y, x) ? 1 : 0;  // phase launches (strided.cu)
  return (exact ? exact_wgrad_supported(g) : tc_wgrad_supported(g, y, x)) ? 1 : 0;
}

int srb_conv_backward_folds_ps(const srb_conv_params *p, int x_cl, int dz_cl) {
  Geom g;
  if (make_geom(p, &g) || p->transposed || g.ps <= 1 || g.st != 1 || g.kh != g.kw) return 0;
  if (!(is_tf32_math(p->math) || p->math == SRB_MATH_BF16) || !x_cl || !dz_cl) return 0;
  const int dt = p->math == SRB_MATH_BF16 ? SRB_BF16 : SRB_F32;
  if (p->math == SRB_MATH_BF16 && (bf16_act_dtype(g.Ci) != SRB_BF16 || bf16_act_dtype(p->Cout) != SRB_BF16)) return 0;
  // dense channels_last views: x (N,Ci,Hi,Wi); dz in y's layout (N,Cout,Ho*r,Wo*r)
  T4 x{(float *)256, (long long)g.Hi * g.Wi * g.Ci, 1, (long long)g.Wi * g.Ci, g.Ci, dt};
  T4 z{(float *)256, (long long)g.Ho * g.ps * g.Wo * g.ps * p->Cout, 1, (long long)g.Wo * g.ps * p->Cout, p->Cout, dt};
  Geom gd{g.N, g.Co, g.Ho, g.Wo, g.Ci, g.Hi, g.Wi, g.kh, g.kw, 1, g.kh - 1 - g.pad, 1};
  Geom g1 = g;
  g1.ps = 1;
  return (gd.pad >= 0 && tc_conv_supported(gd, z, x, true, g.ps) && tc_wgrad_supported(g1, z, x, g.ps)) ? 1 : 0;
}

/* Debug only (not in the public header): per-CTA phase timestamps of the next k_conv_sl launches go to buf (8 x int64 per CTA). */
void srb_debug_set_trace(void *buf, long long max_ctas) { tc_conv_set_trace((long long *)buf, max_ctas); }
/* Debug / A-B knobs (not in the public header; tests and tools/ only).  Bits:
 *      1  empty epilogue            2  operands loaded only once       4  no MMAs issued            8  no db column sums
 *     16  skip a fence (sl) / the smem zeroing (wgrad)                 32  plain arrive instead of commit (rs) / no dump (wgrad)
 *     64  cluster multicast ring   128  force k_conv_sl (no row-stacked kernel) / wgrad direct dump
 *    256  never stack filter rows in wgrad        512  always stack them        1024  row-stacked fprop/dgrad regardless of size
 *   8192 / 16384  cap MTB          32768  narrow (320-thread) k_conv_sl        65536  one row stream per CTA in k_conv_rs
 * 262144  no rows+co stacking (rn = 2) in wgrad   524288  prefer it wherever it has a plan */
void srb_debug_set_flags(int flags) { tc_conv_set_dbg(flags); }

int srb_conv_describe_plan(const srb_conv_params *p, int pass, char *buf, size_t n) {
  Geom g;
  if (!buf || n == 0) return SRB_EINVAL;
  buf[0] = 0;
  int rc = make_geom(p, &g);
  if (rc) return rc;
  if (p->transposed || p->math == SRB_MATH_FP32) { snprintf(buf, n, "fp32 CUDA-core kernels"); return SRB_OK; }
  if (p->math == SRB_MATH_EXACT && pass != 2) g.Ci = (3 * g.Ci + 3) / 4 * 4;  // the channel-tripled launch (exact.cu)
  if (p->math == SRB_MATH_BF16) {
    if (pass == 0) tc_conv_describe(g, buf, n, g.Ci > 4);
    else if (pass == 1) {
      Geom gd{g.N, g.Co, g.Ho, g.Wo, g.Ci, g.Hi, g.Wi, g.kh, g.kw, 1, g.kh - 1 - g.pad, 1};
      tc_conv_describe(gd, buf, n, gd.Ci > 4);
    } else tc_wgrad_describe(g, buf, n, g.Ci > 4 && g.Co >= 8);
    return SRB_OK;
  }
  if (pass == 0) tc_conv_describe(g, buf, n);
  else if (pass == 1) {
    Geom gd{g.N, g.Co, g.Ho, g.Wo, g.Ci, g.Hi, g.Wi, g.kh, g.kw, 1, g.kh - 1 - g.pad, 1};
    if (g.st != 1 || gd.pad < 0) snprintf(buf, n, "fp32 CUDA-core kernels"); else tc_conv_describe(gd, buf, n);
  } else tc_wgrad_describe(g, buf, n);
  return SRB_OK;
}

size_t srb_conv_workspace_bytes(const srb_conv_params *p, int pass) {
  Geom g;
  if (make_geom(p, &g)) return 0;
  size_t a = 0, b = 0;
  if (pass == 2) {
    a = simt_wgrad_ws_bytes(g);
    b = tc_wgrad_ws_bytes(g);
    if (g.ps > 1 && !p->transposed) {  // PixelShuffle layer: dz may be un-shuffled by the kernel's own TMA traversal
      Geom g1 = g;
      g1.ps = 1;
      const size_t c1 = tc_wgrad_ws_bytes(g1, false, g.ps), c2 = tc_wgrad_ws_bytes(g1, true, g.ps);
      if (c1 > b) b = c1;
      if (c2 > b) b = c2;
    }
    if (g.st > 1 && !p->transposed) { const size_t c = tc_strided_wgrad_ws_bytes(g); if (c > b) b = c; }
    SkinnyWg sk;
    T4 cl{(float *)256, (long long)g.Hi * g.Wi * g.Ci, 1, (long long)g.Wi * g.Ci, g.Ci};  // would-be channels_last x
    if (skinny_wgrad_plan(p, g, cl, &sk) && sk.total > b) b = sk.total;
  } else {
    b = tc_conv_ws_bytes(g);
    if (pass == 0) b += 40 * 1024;  // per-warp loss partials of srb_conv_fprop_loss
    if (g.st > 1) { const size_t c = tc_strided_ws_bytes(g); if (c > b) b = c; }
  }
  if (p->math == SRB_MATH_BF16 && !p->transposed && g.st == 1 && pass == 2) {
    // bf16 wgrad + the two mixed-edge conversions (dz -> fp32 NHWC for Cin <= 4; dz -> bf16 NHWC8 for Cout < 8)
    Geom g8 = g;
    g8.Co = g.Co < 8 ? 8 : g.Co;
    size_t c = tc_wgrad_ws_bytes(g8, true) + a256((size_t)g.N * g.Ho * g.Wo * 8 * 2) + a256((size_t)8 * g.Ci * g.kh * g.kw * 4 + 64) +
               (g.Ci <= 4 ? a256((size_t)g.N * g.Ho * g.Wo * g.Co * 4) + tc_wgrad_ws_bytes(g, false) : 0) + 2048;
    if (c > b) b = c;
  }
  if (p->math == SRB_MATH_EXACT && !p->transposed && g.st == 1) {
    size_t c = 0;
    if (pass == 2) c = exact_wgrad_ws_bytes(g);
    else if (pass == 0) c = exact_conv_ws_bytes(g);
    else {
      Geom gd{g.N, g.Co, g.Ho, g.Wo, g.Ci, g.Hi, g.Wi, g.kh, g.kw, 1, g.kh - 1 - g.pad, 1};
      if (gd.pad >= 0) c = exact_conv_ws_bytes(gd);
    }
    if (c > b) b = c;
  }
  return (a > b ? a : b) + 256;
}

int srb_conv_fprop(const srb_conv_params *p, const srb_tensor4 *x, const float *w, const float *bias,
                   const float *alpha, const srb_tensor4 *residual, const srb_tensor4 *y, const srb_tensor4 *preact,
                   uint16_t *relu_bits, void *ws, size_t ws_bytes, void *stream) {
  Geom g;
  int rc = make_geom(p, &g);
  if (rc) return rc;
  if (p->N == 0) return SRB_OK;  // empty batch: nothing to compute
  const WeightCacheScope wc_scope(p->math != SRB_MATH_EXACT);  // `w` is the caller's filter: its packed copies may be cached
  SRB_REQUIRE(x && x->data && w && y && y->data, SRB_EINVAL, "null tensor");
  SRB_REQUIRE(p->act != SRB_ACT_PRELU || alpha, SRB_EINVAL, "PReLU needs alpha");
  cudaStream_t st = (cudaStream_t)stream;
  T4 tx = to_t4(x), ty = to_t4(y);
  if (p->math == SRB_MATH_BF16) {
    SRB_REQUIRE(!p->transposed && g.st == 1, SRB_EUNSUPPORTED, "bf16 storage mode: strided / transposed convolutions are not built");
    Epi e = make_epi(p, bias, alpha, residual, preact, 0);
    e.bits_out = relu_bits;
    SRB_REQUIRE(tc_conv_supported(g, tx, ty, false), SRB_EUNSUPPORTED,
                "bf16 storage mode: x must be a bf16 channels_last tensor with Cin %% 8 == 0 (or fp32 with Cin <= 4)");
    SRB_REQUIRE(!relu_bits || (p->ps == 1 && (p->Cout & 15) == 0), SRB_EUNSUPPORTED, "relu_bits needs no PixelShuffle and Cout %% 16 == 0");
    return tc_conv_gather(g, tx, w, false, ty, e, ws, ws_bytes, st);
  }
  SRB_REQUIRE(tx.dt == SRB_F32 && ty.dt == SRB_F32, SRB_EUNSUPPORTED, "bf16 tensors need math = SRB_MATH_BF16");
  if (!p->transposed) {
    Epi e = make_epi(p, bias, alpha, residual, preact, want_round(p, ty, p->Cout));
    e.bits_out = relu_bits;
    const bool ex = p->math == SRB_MATH_EXACT && exact_conv_supported(g, ty);
    const bool tc = ex || (is_tf32_math(p->math) && tc_conv_supported(g, tx, ty, false));
    SRB_REQUIRE(!relu_bits || (tc && p->ps == 1 && (p->Cout & 15) == 0), SRB_EUNSUPPORTED,
                "relu_bits needs the tensor path, no PixelShuffle and Cout %% 16 == 0");
    if (ex) return exact_conv_gather(g, tx, w, false, ty, e, ws, ws_bytes, st);
    if (tc) return tc_conv_gather(g, tx, w, false, ty, e, ws, ws_bytes, st);
    // strided Conv2d (SRGAN D, srgan.py:54-63): the input's st x st phase images as extra channels of a stride-1 tensor-core conv
    if (is_tf32_math(p->math) && g.st > 1 && tc_strided_gather_supported(g, tx, ty)) return tc_strided_gather(g, tx, w, ty, e, ws, ws_bytes, st);
    return simt_conv_gather(g, tx, w, ty, e, st);
  }
  SRB_REQUIRE(!relu_bits, SRB_EUNSUPPORTED, "relu_bits with a transposed convolution");
  Epi e = make_epi(p, bias, alpha, residual, preact, want_round(p, ty, p->Cout));
  // ConvTranspose2d (fsrcnn.py:33, DeconvBlock base_networks.py:77): st x st output phases, each a stride-1 tensor-core conv
  if (is_tf32_math(p->math) && tc_strided_scatter_supported(g, tx, ty)) return tc_strided_scatter(g, tx, w, ty, e, ws, ws_bytes, st);
  return simt_conv_scatter(g, tx, w, ty, e, st);
}

int srb_conv_fprop_loss(const srb_conv_params *p, const srb_tensor4 *x, const float *w, const float *bias,
                        const srb_tensor4 *target, int loss_kind, const srb_tensor4 *y, const srb_tensor4 *dz, int dz_unshuffled,
                        float *loss, void *ws, size_t ws_bytes, void *stream) {
  Geom g;
  int rc = make_geom(p, &g);
  if (rc) return rc;
  const WeightCacheScope wc_scope(p->math != SRB_MATH_EXACT);  // `w` is the caller's filter: its packed copies may be cached
  SRB_REQUIRE(p->N > 0 && x && x->data && w && target && target->data && dz && dz->data && loss, SRB_EINVAL, "null tensor");
  SRB_REQUIRE(loss_kind == 0 || loss_kind == 1, SRB_EINVAL, "loss kind: 0 = MSE, 1 = L1");
  SRB_REQUIRE(!p->transposed && p->act == SRB_ACT_NONE && (is_tf32_math(p->math) || p->math == SRB_MATH_BF16), SRB_EUNSUPPORTED,
              "fused loss: Conv2d without activation on the tensor path (math auto / tf32 / bf16)");
  T4 tx = to_t4(x), ty = to_t4(y), tt = to_t4(target), tdz = to_t4(dz);
  SRB_REQUIRE(tt.dt == SRB_F32 || tt.dt == SRB_U8, SRB_EUNSUPPORTED, "fused loss: the target is fp32, or uint8 image bytes (t = byte / 255)");
  if (!ty.p) {  // geometry of y without storing it: the target's if that is fp32, a dense NCHW one beside a uint8 target
    ty = tt; ty.p = nullptr;
    if (tt.dt == SRB_U8) {
      const long long Wy = (long long)g.Wo * g.ps, Hy = (long long)g.Ho * g.ps, Cy = g.Co / (g.ps * g.ps);
      ty.dt = SRB_F32; ty.sw = 1; ty.sh = Wy; ty.sc = Hy * Wy; ty.sn = Cy * Hy * Wy;
    }
  }
  T4 probe = ty;
  probe.p = (float *)256;
  SRB_REQUIRE(tc_conv_supported(g, tx, probe, false), SRB_EUNSUPPORTED, "fused loss: this layer does not run on the tensor path");
  Epi e = make_epi(p, bias, nullptr, nullptr, nullptr, 0);
  e.loss_kind = loss_kind + 1;
  e.loss_coef = (float)(1.0 / ((double)g.N * g.Ho * g.Wo * g.Co));
  e.target = tt;
  e.round_tf32 = is_tf32_math(p->math) ? 1 : 0;  // rounds the emitted gradient (it feeds tf32 dgrad / wgrad)
  if (dz_unshuffled) {
    SRB_REQUIRE(tdz.dt == SRB_F32 && tdz.sc == 1 && tdz.sw == g.Co && tdz.sh == (long long)g.Wo * g.Co &&
                    tdz.sn == (long long)g.Ho * g.Wo * g.Co, SRB_EINVAL, "un-shuffled dz must be a dense NHWC (N,Cout*r*r,Ho,Wo) tensor");
    e.dz_unshuf = tdz.p;
  } else {
    e.dz = tdz;
  }
  ConvOpt lo;
  lo.loss_out = loss;
  return tc_conv_gather(g, tx, w, false, ty, e, ws, ws_bytes, (cudaStream_t)stream, lo);
}

int srb_weight_cache_enable(int on) { return tc_weight_cache_enable(on); }
int srb_weight_cache_repack(void *stream) { return tc_weight_cache_repack((cudaStream_t)stream); }
int srb_weight_cache_entries(void) { return tc_weight_cache_entries(); }

int srb_adam_step_flat(float *p, const float *g, float *m, float *v, int64_t n, float lr, float beta1, float beta2, float eps,
                       float weight_decay, float *state, void *stream) {
  SRB_REQUIRE(p && g && m && v && state && n >= 0 && lr >= 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f,
              SRB_EINVAL, "bad Adam arguments");
  if (n == 0) return SRB_OK;
  launch_pdl(k_adam_flat, dim3(ew_blocks(n)), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, (long long)n, lr, beta1, beta2, eps, weight_decay,
             state);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

int srb_scale_by_scalar(float *x, int64_t n, const float *g, int round_to_tf32, void *stream) {
  SRB_REQUIRE(x && g && n >= 0, SRB_EINVAL, "bad scale args");
  if (n == 0) return SRB_OK;
  launch_pdl(k_scale_by_scalar, dim3(ew_blocks(n)), dim3(256), 0, (cudaStream_t)stream, x, n, g, round_to_tf32);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

int srb_act_bwd(const srb_conv_params *p, const srb_tensor4 *dy, const srb_tensor4 *ref, const float *alpha,
                const srb_tensor4 *dz, float *dalpha, void *stream) {
  Geom g;
  int rc = make_geom(p, &g);
  if (rc) return rc;
  if (p->N == 0) return SRB_OK;
  SRB_REQUIRE(dy && ref && dz && dy->data && ref->data && dz->data, SRB_EINVAL, "null tensor");
  SRB_REQUIRE(p->act != SRB_ACT_NONE, SRB_EINVAL, "act_bwd with act none");
  SRB_REQUIRE(p->act != SRB_ACT_PRELU || alpha, SRB_EINVAL, "PReLU needs alpha");
  int C = p->Cout, H, W;
  if (!p->transposed) { H = g.Ho * g.ps; W = g.Wo * g.ps; }
  else                { H = g.Hi; W = g.Wi; }
  T4 tdz = to_t4(dz);
  long long total = (long long)p->N * C * H * W;
  if (total == 0) return SRB_OK;
  if (tdz.dt == SRB_BF16) {
    T4 a = to_t4(dy), b = to_t4(ref);
    auto same = [&](const T4 &t) { return t.dt == SRB_BF16 && t.sn == tdz.sn && t.sc == tdz.sc && t.sh == tdz.sh && t.sw == tdz.sw; };
    const bool nhwc = tdz.sc == 1 && tdz.sw == C && tdz.sh == (long long)W * C && tdz.sn == (long long)H * W * C;
    const bool nchw = tdz.sw == 1 && tdz.sh == W && tdz.sc == (long long)H * W && tdz.sn == (long long)C * H * W;
    SRB_REQUIRE(same(a) && same(b) && (nhwc || nchw), SRB_EUNSUPPORTED, "bf16 act_bwd needs dy, ref, dz in one dense layout");
    launch_pdl(k_act_bwd_flat_h, dim3(ew_blocks(total)), dim3(256), 0, (cudaStream_t)stream, (const unsigned short *)a.p, (const unsigned short *)b.p,
                                                                         (unsigned short *)tdz.p, total, p->act, p->slope, alpha, dalpha);
    count_launch();
    SRB_CHECK_CUDA(cudaGetLastError());
    return SRB_OK;
  }
  {
    T4 a = to_t4(dy), b = to_t4(ref);
    auto dense = [&](const T4 &t) {
      bool nchw = t.sw == 1 && t.sh == W && t.sc == (long long)H * W && t.sn == (long long)C * H * W;
      bool nhwc = t.sc == 1 && t.sw == C && t.sh == (long long)W * C && t.sn == (long long)H * W * C;
      return nchw ? 1 : (nhwc ? 2 : 0);
    };
    int la = dense(a), lb = dense(b), lc = dense(tdz);
    if (la && la == lb && la == lc && (total & 3) == 0 &&
        ((((uintptr_t)a.p) | ((uintptr_t)b.p) | ((uintptr_t)tdz.p)) & 15) == 0) {
      launch_pdl(k_act_bwd_flat, dim3(ew_blocks(total / 4)), dim3(256), 0, (cudaStream_t)stream, 
          (const float4 *)a.p, (const float4 *)b.p, (float4 *)tdz.p, total / 4, p->act, p->slope, alpha, dalpha,
          want_round(p, tdz, C));
      count_launch();
      SRB_CHECK_CUDA(cudaGetLastError());
      return SRB_OK;
    }
  }
  int cl = (tdz.sc == 1) ? 1 : 0;
  launch_pdl(k_act_bwd, dim3(ew_blocks(total)), dim3(256), 0, (cudaStream_t)stream, to_t4(dy), to_t4(ref), tdz, p->N, C, H, W, p->act,
                                                                 p->slope, alpha, dalpha, cl,
                                                                 want_round(p, tdz, C));
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

int srb_conv_dgrad(const srb_conv_params *p, const srb_tensor4 *dz, const float *w, const srb_tensor4 *relu_mask,
                   const uint16_t *relu_bits, const srb_tensor4 *dx, void *ws, size_t ws_bytes, void *stream) {
  return srb_conv_dgrad_add(p, dz, w, relu_mask, relu_bits, nullptr, dx, ws, ws_bytes, stream);
}

int srb_conv_dgrad_add(const srb_conv_params *p, const srb_tensor4 *dz, const float *w, const srb_tensor4 *relu_mask,
                       const uint16_t *relu_bits, const srb_tensor4 *add, const srb_tensor4 *dx, void *ws, size_t ws_bytes,
                       void *stream) {
  Geom g;
  int rc = make_geom(p, &g);
  if (rc) return rc;
  if (p->N == 0) return SRB_OK;
  const WeightCacheScope wc_scope(p->math != SRB_MATH_EXACT);  // `w` is the caller's filter: its packed copies may be cached
  SRB_REQUIRE(dz && dz->data && w && dx && dx->data, SRB_EINVAL, "null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  T4 tdz = to_t4(dz), tdx = to_t4(dx);
  Epi e;
  memset(&e, 0, sizeof(e));
  e.act = SRB_ACT_NONE;
  e.mask = to_t4(relu_bits ? nullptr : relu_mask);  // the packed pattern wins when both are given
  e.bits_out = nullptr;
  e.bits_in = relu_bits;
  e.round_tf32 = want_round(p, tdx, p->Cin);
  if (add && add->data) {
    // dx = dgrad(dz) + add through the epilogue's residual operand (added before the ReLU mask): stride-1 Conv2d without
    // PixelShuffle on the tensor path only -- decided before anything is launched
    T4 tadd = to_t4(add);
    Geom gd{g.N, g.Co, g.Ho, g.Wo, g.Ci, g.Hi, g.Wi, g.kh, g.kw, 1, g.kh - 1 - g.pad, 1};
    const bool tensor_ok = !p->transposed && g.ps == 1 && g.st == 1 && g.kh == g.kw && gd.pad >= 0 &&
                           (is_tf32_math(p->math) || p->math == SRB_MATH_BF16) && tadd.dt == tdx.dt &&
                           (p->math == SRB_MATH_BF16 || (tdz.dt == SRB_F32 && tdx.dt == SRB_F32)) &&
                           tc_conv_supported(gd, tdz, tdx, true, 1);
    SRB_REQUIRE(tensor_ok, SRB_EUNSUPPORTED, "dgrad with an added gradient: stride-1 Conv2d on the tensor path (math auto / tf32 / bf16)");
    e.residual = tadd;
  }
  if (p->math == SRB_MATH_BF16) {
    SRB_REQUIRE(!p->transposed && g.st == 1 && g.kh == g.kw, SRB_EUNSUPPORTED,
                "bf16 storage mode: dgrad needs a stride-1 square-kernel Conv2d");
    Geom gd{g.N, g.Co, g.Ho, g.Wo, g.Ci, g.Hi, g.Wi, g.kh, g.kw, 1, g.kh - 1 - g.pad, 1};
    SRB_REQUIRE(gd.pad >= 0 && tc_conv_supported(gd, tdz, tdx, true, g.ps), SRB_EUNSUPPORTED,
                "bf16 storage mode: dz must be bf16 channels_last with Cout %% 8 == 0 (PixelShuffle layers: Cout %% 64 == 0, "
                "else un-shuffle first), or fp32 with Cout <= 4");
    SRB_REQUIRE(!relu_bits || (p->Cin & 15) == 0, SRB_EUNSUPPORTED, "relu_bits needs Cin %% 16 == 0");
    SRB_REQUIRE(!e.mask.p, SRB_EUNSUPPORTED, "bf16 storage mode: pass relu_bits, not relu_mask");
    e.round_tf32 = 0;
    ConvOpt po;
    po.in_ps = g.ps;
    return tc_conv_gather(gd, tdz, w, true, tdx, e, ws, ws_bytes, st, po);
  }
  SRB_REQUIRE(tdz.dt == SRB_F32 && tdx.dt == SRB_F32, SRB_EUNSUPPORTED, "bf16 tensors need math = SRB_MATH_BF16");
  if (!p->transposed) {
    if (is_tf32_math(p->math) && g.ps > 1 && g.st == 1 && g.kh == g.kw) {
      // PixelShuffle layer: dz arrives in y's (shuffled) layout; the un-shuffle is the TMA traversal of the A operand
      Geom gd{g.N, g.Co, g.Ho, g.Wo, g.Ci, g.Hi, g.Wi, g.kh, g.kw, 1, g.kh - 1 - g.pad, 1};
      if (gd.pad >= 0 && tc_conv_supported(gd, tdz, tdx, true, g.ps)) {
        SRB_REQUIRE(!relu_bits || (p->Cin & 15) == 0, SRB_EUNSUPPORTED, "relu_bits needs Cin %% 16 == 0");
        ConvOpt po;
        po.in_ps = g.ps;
        return tc_conv_gather(gd, tdz, w, true, tdx, e, ws, ws_bytes, st, po);
      }  // else: the CUDA-core scatter kernel below un-shuffles through its own addressing
    }
    if (p->math != SRB_MATH_FP32 && g.ps == 1 && g.st == 1 && g.kh == g.kw) {
      // stride-1 dgrad == gather conv of dz with the flipped, transposed filter and pad' = k-1-pad
      Geom gd{g.N, g.Co, g.Ho, g.Wo, g.Ci, g.Hi, g.Wi, g.kh, g.kw, 1, g.kh - 1 - g.pad, 1};
      if (p->math == SRB_MATH_EXACT) {
        if (gd.pad >= 0 && exact_conv_supported(gd, tdx)) {
          SRB_REQUIRE(!relu_bits || (p->Cin & 15) == 0, SRB_EUNSUPPORTED, "relu_bits needs Cin %% 16 == 0");
          return exact_conv_gather(gd, tdz, w, true, tdx, e, ws, ws_bytes, st);
        }
      } else if (gd.pad >= 0 && tc_conv_supported(gd, tdz, tdx, true)) {
        SRB_REQUIRE(!relu_bits || (p->Cin & 15) == 0, SRB_EUNSUPPORTED, "relu_bits needs Cin %% 16 == 0");
        return tc_conv_gather(gd, tdz, w, true, tdx, e, ws, ws_bytes, st);
      }
    }
    SRB_REQUIRE(!relu_bits, SRB_EUNSUPPORTED, "relu_bits needs the tensor-path dgrad (use relu_mask)");
    if (is_tf32_math(p->math) && g.st > 1 && tc_strided_scatter_supported(g, tdz, tdx)) return tc_strided_scatter(g, tdz, w, tdx, e, ws, ws_bytes, st);
    return simt_conv_scatter(g, tdz, w, tdx, e, st);
  }
  SRB_REQUIRE(!relu_bits, SRB_EUNSUPPORTED, "relu_bits with a transposed convolution (use relu_mask)");
  // ConvTranspose2d backward-data is a plain gather conv of dz (big side) producing dx (small side)
  if (is_tf32_math(p->math) && g.st > 1 && tc_strided_gather_supported(g, tdz, tdx)) return tc_strided_gather(g, tdz, w, tdx, e, ws, ws_bytes, st);
  return simt_conv_gather(g, tdz, w, tdx, e, st);
}

int srb_conv_wgrad(const srb_conv_params *p, const srb_tensor4 *x, const srb_tensor4 *dz, float *dw, float *db,
                   float scale, int accumulate, void *ws, size_t ws_bytes, void *stream) {
  Geom g;
  int rc = make_geom(p, &g);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->N == 0) {  // empty batch: the gradient is exactly zero
    SRB_REQUIRE(dw, SRB_EINVAL, "null tensor");
    if (!accumulate) {
      size_t wn = (size_t)p->Cin * p->Cout * p->ps * p->ps * p->kh * p->kw;
      SRB_CHECK_CUDA(cudaMemsetAsync(dw, 0, wn * sizeof(float), st));
      if (db) SRB_CHECK_CUDA(cudaMemsetAsync(db, 0, (size_t)p->Cout * p->ps * p->ps * sizeof(float), st));
    }
    return SRB_OK;
  }
  SRB_REQUIRE(x && x->data && dz && dz->data && dw, SRB_EINVAL, "null tensor");
  T4 tx = to_t4(x), tdz = to_t4(dz);
  if (p->math == SRB_MATH_BF16) {
    SRB_REQUIRE(!p->transposed && g.st == 1, SRB_EUNSUPPORTED, "bf16 storage mode: wgrad needs a stride-1 Conv2d");
    uintptr_t wsp = ((uintptr_t)ws + 255) & ~(uintptr_t)255;
    const uintptr_t ws_end = (uintptr_t)ws + ws_bytes;
    const int z_ps = g.ps;
    g.ps = 1;  // from here on g is the plain conv; dz's shuffled layout is handled by the dz TMA traversal (z_ps)
    if (tx.dt == SRB_BF16 && tdz.dt == SRB_BF16) {
      SRB_REQUIRE(tc_wgrad_supported(g, tdz, tx, z_ps), SRB_EUNSUPPORTED,
                  "bf16 wgrad: no plan for this layer (PixelShuffle layers need Cout %% 64 == 0, else un-shuffle dz first)");
      return tc_conv_wgrad(g, tdz, tx, dw, db, scale, accumulate, ws, ws_bytes, st, z_ps);
    }
    SRB_REQUIRE(z_ps == 1, SRB_EUNSUPPORTED, "bf16 storage mode: PixelShuffle layer with an fp32 side (un-shuffle dz first)");
    if (tx.dt == SRB_F32 && tdz.dt == SRB_BF16) {
      // network input layer (Cin <= 4): dz -> fp32 NHWC (exactly representable in tf32), then the tf32 c4 wgrad
      SRB_REQUIRE(g.Ci <= 4, SRB_EUNSUPPORTED, "bf16 storage mode: fp32 x with Cin > 4");
      const size_t zb = a256((size_t)g.N * g.Ho * g.Wo * g.Co * sizeof(float));
      SRB_REQUIRE(ws && wsp + zb <= ws_end, SRB_EWORKSPACE, "bf16 wgrad workspace too small");
      float *z32 = (float *)wsp;
      launch_pdl(k_h2f_nhwc, dim3(ew_blocks((long long)g.N * g.Ho * g.Wo * g.Co)), dim3(256), 0, st, tdz, z32, g.N, g.Co, g.Ho, g.Wo);
      count_launch();
      SRB_CHECK_CUDA(cudaGetLastError());
      T4 tz32{z32, (long long)g.Ho * g.Wo * g.Co, 1, (long long)g.Wo * g.Co, g.Co, SRB_F32};
      SRB_REQUIRE(tc_wgrad_supported(g, tz32, tx), SRB_EUNSUPPORTED, "bf16 storage mode: no tf32 plan for the input layer's wgrad");
      return tc_conv_wgrad(g, tz32, tx, dw, db, scale, accumulate, (void *)(wsp + zb), (size_t)(ws_end - (wsp + zb)), st);
    }
    if (tx.dt == SRB_BF16 && tdz.dt == SRB_F32) {
      // network output layer (Cout < 8): dz -> bf16 NHWC8, wgrad for 8 output channels, keep the first Cout filters
      SRB_REQUIRE(g.Co < 8 && !accumulate, SRB_EUNSUPPORTED, "bf16 storage mode: fp32 dz needs Cout < 8 and accumulate == 0");
      Geom g8 = g;
      g8.Co = 8;
      const size_t pb = a256((size_t)g.N * g.Ho * g.Wo * 8 * 2), db8 = a256((size_t)8 * g.Ci * g.kh * g.kw * 4 + 64);
      SRB_REQUIRE(ws && wsp + pb + db8 <= ws_end, SRB_EWORKSPACE, "bf16 wgrad workspace too small");
      unsigned short *pack = (unsigned short *)wsp;
      float *dw8 = (float *)(wsp + pb);
      float *dbias8 = dw8 + (size_t)8 * g.Ci * g.kh * g.kw;
      launch_pdl(k_pack_dz8_h, dim3(ew_blocks((long long)g.N * g.Ho * g.Wo)), dim3(256), 0, st, tdz, (uint4 *)pack, g.N, g.Co, g.Ho, g.Wo);
      count_launch();
      SRB_CHECK_CUDA(cudaGetLastError());
      T4 tz8{(float *)pack, (long long)g.Ho * g.Wo * 8, 1, (long long)g.Wo * 8, 8, SRB_BF16};
      SRB_REQUIRE(tc_wgrad_supported(g8, tz8, tx), SRB_EUNSUPPORTED, "bf16 storage mode: no plan for the output layer's wgrad");
      rc = tc_conv_wgrad(g8, tz8, tx, dw8, db ? dbias8 : nullptr, scale, 0, (void *)(wsp + pb + db8), (size_t)(ws_end - (wsp + pb + db8)), st);
      if (rc) return rc;
      SRB_CHECK_CUDA(cudaMemcpyAsync(dw, dw8, (size_t)g.Co * g.Ci * g.kh * g.kw * sizeof(float), cudaMemcpyDeviceToDevice, st));
      if (db) SRB_CHECK_CUDA(cudaMemcpyAsync(db, dbias8, (size_t)g.Co * sizeof(float), cudaMemcpyDeviceToDevice, st));
      return SRB_OK;
    }
    SRB_REQUIRE(false, SRB_EUNSUPPORTED, "bf16 storage mode: fp32 x and fp32 dz (use math = auto)");
  }
  SRB_REQUIRE(tx.dt == SRB_F32 && tdz.dt == SRB_F32, SRB_EUNSUPPORTED, "bf16 tensors need math = SRB_MATH_BF16");
  if (!p->transposed) {
    if (p->math == SRB_MATH_EXACT && exact_wgrad_supported(g))
      return exact_conv_wgrad(g, tdz, tx, dw, db, scale, accumulate, ws, ws_bytes, st);
    if (is_tf32_math(p->math) && g.ps > 1) {
      Geom g1 = g;
      g1.ps = 1;
      if (tc_wgrad_supported(g1, tdz, tx, g.ps))
        return tc_conv_wgrad(g1, tdz, tx, dw, db, scale, accumulate, ws, ws_bytes, st, g.ps);
      // else: the CUDA-core wgrad below un-shuffles through its own addressing
    }
    if (is_tf32_math(p->math) && tc_wgrad_supported(g, tdz, tx))
      return tc_conv_wgrad(g, tdz, tx, dw, db, scale, accumulate, ws, ws_bytes, st);
    if (is_tf32_math(p->math) && g.st > 1 && tc_strided_wgrad_supported(g, tdz, tx))  // SRGAN D (srgan.py:54-63): st*st phase launches
      return tc_strided_wgrad(g, tdz, tx, dw, db, scale, accumulate, ws, ws_bytes, st);
    SkinnyWg sk;
    uintptr_t wsp = ((uintptr_t)ws + 255) & ~(uintptr_t)255;
    if (!accumulate && ws && skinny_wgrad_plan(p, g, tx, &sk) && wsp + sk.total <= (uintptr_t)ws + ws_bytes) {
      // 64->3 / 32->3 tails: pad dz to NHWC4, run the tensor-core wgrad for 4 output channels, keep the first Co
      float *pack = (float *)wsp;
      float *dw4 = (float *)(wsp + sk.pack_bytes);
      float *db4 = dw4 + (size_t)4 * g.Ci * g.kh * g.kw;
      void *ws_tc = (void *)(wsp + sk.pack_bytes + sk.dw_bytes);
      const long long px = (long long)g.N * g.Ho * g.Wo;
      launch_pdl(k_pack_dz4, dim3(ew_blocks(px)), dim3(256), 0, st, tdz, (float4 *)pack, g.N, g.Co, g.Ho, g.Wo);
      count_launch();
      SRB_CHECK_CUDA(cudaGetLastError());
      sk.dz4.p = pack;
      rc = tc_conv_wgrad(sk.g4, sk.dz4, tx, dw4, db ? db4 : nullptr, scale, 0, ws_tc,
                         (size_t)((uintptr_t)ws + ws_bytes - (uintptr_t)ws_tc), st);
      if (rc) return rc;
      SRB_CHECK_CUDA(cudaMemcpyAsync(dw, dw4, (size_t)g.Co * g.Ci * g.kh * g.kw * sizeof(float), cudaMemcpyDeviceToDevice, st));
      if (db) SRB_CHECK_CUDA(cudaMemcpyAsync(db, db4, (size_t)g.Co * sizeof(float), cudaMemcpyDeviceToDevice, st));
      return SRB_OK;
    }
    return simt_conv_wgrad(g, tdz, tx, dw, db, scale, accumulate, ws, ws_bytes, st);
  }
  // transposed: small = x (Co := Cin), big = dz (Ci := Cout); db is a channel sum over the big side
  rc = simt_conv_wgrad(g, tx, tdz, dw, nullptr, scale, accumulate, ws, ws_bytes, st);
  if (rc) return rc;
  if (db) return channel_sum(tdz, g.N, g.Ci, g.Hi, g.Wi, db, scale, accumulate, st);
  return SRB_OK;
}

int srb_pixel_unshuffle(const srb_conv_params *p, const srb_tensor4 *dz, const srb_tensor4 *out, void *stream) {
  Geom g;
  int rc = make_geom(p, &g);
  if (rc) return rc;
  SRB_REQUIRE(!p->transposed && p->ps > 1, SRB_EINVAL, "pixel_unshuffle needs a PixelShuffle conv");
  if (p->N == 0) return SRB_OK;
  SRB_REQUIRE(dz && dz->data && out && out->data, SRB_EINVAL, "null tensor");
  SRB_REQUIRE(out->sc == 1, SRB_EINVAL, "pixel_unshuffle writes channels_last");
  if (dz->dtype == SRB_BF16 || out->dtype == SRB_BF16) {
    SRB_REQUIRE(dz->dtype == SRB_BF16 && out->dtype == SRB_BF16 && dz->sc == 1 && (p->Cout % 8) == 0 &&
                    (dz->sw % 8) == 0 && (dz->sh % 8) == 0 && (dz->sn % 8) == 0 && (((uintptr_t)dz->data) & 15) == 0,
                SRB_EUNSUPPORTED, "bf16 pixel_unshuffle: both tensors bf16 channels_last, C %% 8 == 0");
    const long long tot = (long long)g.N * g.Ho * g.Wo * g.ps * g.ps * (p->Cout / 8);
    launch_pdl(k_pixel_unshuffle_h, dim3(ew_blocks(tot)), dim3(256), 0, (cudaStream_t)stream, to_t4(dz), to_t4(out), g.N, p->Cout, g.Ho, g.Wo, g.ps);
    count_launch();
    SRB_CHECK_CUDA(cudaGetLastError());
    return SRB_OK;
  }
  long long total = (long long)g.N * g.Ho * g.Wo * g.Co;
  const int rnd = is_tf32_math(p->math) ? 1 : 0;  // EXACT / FP32 keep the full fp32 gradient
  const size_t row_smem = (size_t)p->Cout * g.ps * ((size_t)g.Wo * g.ps + 4) * sizeof(float);
  if (dz->sw == 1 && row_smem <= 48 * 1024 && (long long)g.N * g.Ho < (1LL << 31)) {
    launch_pdl(k_pixel_unshuffle_rows, dim3((unsigned)(g.N * g.Ho)), dim3(256), row_smem, (cudaStream_t)stream, to_t4(dz), to_t4(out), g.N, p->Cout,
                                                                                            g.Ho, g.Wo, g.ps, rnd);
    count_launch();
    SRB_CHECK_CUDA(cudaGetLastError());
    return SRB_OK;
  }
  launch_pdl(k_pixel_unshuffle, dim3(ew_blocks(total)), dim3(256), 0, (cudaStream_t)stream, to_t4(dz), to_t4(out), g.N, g.Co, g.Ho, g.Wo,
                                                                         g.ps, rnd);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

int srb_prelu_fwd(const float *x, const float *alpha, float *y, int64_t n, void *stream) {
  SRB_REQUIRE(x && alpha && y && n >= 0, SRB_EINVAL, "bad prelu args");
  if (n == 0) return SRB_OK;
  launch_pdl(k_prelu_fwd, dim3(ew_blocks(n)), dim3(256), 0, (cudaStream_t)stream, x, alpha, y, n);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

int srb_prelu_bwd(const float *x, const float *dy, const float *alpha, float *dx, float *dalpha, int64_t n,
                  void *stream) {
  SRB_REQUIRE(x && dy && alpha && dx && n >= 0, SRB_EINVAL, "bad prelu args");
  if (n == 0) return SRB_OK;
  launch_pdl(k_prelu_bwd, dim3(ew_blocks(n)), dim3(256), 0, (cudaStream_t)stream, x, dy, alpha, dx, dalpha, n);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

size_t srb_loss_workspace_bytes(void) { return 148 * 16 * sizeof(float); }

int srb_loss_fwd(int kind, const float *y, const float *t, int64_t n, float *loss, void *ws, size_t ws_bytes, void *stream) {
  SRB_REQUIRE(y && t && loss && n > 0 && (kind == 0 || kind == 1), SRB_EINVAL, "bad loss args");
  SRB_REQUIRE(ws && ws_bytes >= srb_loss_workspace_bytes(), SRB_EWORKSPACE, "loss workspace too small");
  SRB_REQUIRE(((((uintptr_t)y) | ((uintptr_t)t)) & 15) == 0, SRB_EUNSUPPORTED, "loss tensors must be 16-byte aligned");
  const long long n4 = n >> 2;
  const unsigned blocks = ew_blocks(n4 > 0 ? n4 : 1);
  launch_pdl(k_loss_partial, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (const float4 *)y, (const float4 *)t, n4, y + (n4 << 2), t + (n4 << 2),
                                                           (int)(n & 3), kind, (float *)ws);
  launch_pdl(k_loss_finish, dim3(1), dim3(256), 0, (cudaStream_t)stream, (const float *)ws, (int)blocks, 1.0f / (float)n, loss);
  count_launch(2);
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

int srb_loss_bwd(int kind, const float *y, const float *t, int64_t n, const float *grad_loss, float *dy, void *stream) {
  SRB_REQUIRE(y && t && grad_loss && dy && n > 0 && (kind == 0 || kind == 1), SRB_EINVAL, "bad loss args");
  SRB_REQUIRE(((((uintptr_t)y) | ((uintptr_t)t) | ((uintptr_t)dy)) & 15) == 0, SRB_EUNSUPPORTED,
              "loss tensors must be 16-byte aligned");
  launch_pdl(k_loss_bwd, dim3(ew_blocks((n >> 2) > 0 ? (n >> 2) : 1)), dim3(256), 0, (cudaStream_t)stream, y, t, n, kind, 1.0f / (float)n, grad_loss, dy);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

int srb_image_to_tensor(const uint8_t *src_nhwc, float *dst_nchw, int32_t N, int32_t H, int32_t W, int32_t C, float scale,
                        void *stream) {
  SRB_REQUIRE(N >= 0 && H > 0 && W > 0 && C > 0, SRB_EINVAL, "bad image_to_tensor args");
  if (N == 0) return SRB_OK;  // an empty batch has no storage behind it
  SRB_REQUIRE(src_nhwc && dst_nchw, SRB_EINVAL, "null image_to_tensor buffer");
  const bool v4 = C <= 4 && (W & 3) == 0 && ((uintptr_t)src_nhwc & 3) == 0 && ((uintptr_t)dst_nchw & 15) == 0;
  const long long groups = (long long)N * H * (W / 4);
  cudaStream_t st = (cudaStream_t)stream;
  if (v4 && C == 1) launch_pdl(k_image_to_tensor_v4<1>, dim3(ew_blocks(groups)), dim3(256), 0, st, (const uint32_t *)src_nhwc, (float4 *)dst_nchw, N, H, W / 4, scale);
  else if (v4 && C == 2) launch_pdl(k_image_to_tensor_v4<2>, dim3(ew_blocks(groups)), dim3(256), 0, st, (const uint32_t *)src_nhwc, (float4 *)dst_nchw, N, H, W / 4, scale);
  else if (v4 && C == 3) launch_pdl(k_image_to_tensor_v4<3>, dim3(ew_blocks(groups)), dim3(256), 0, st, (const uint32_t *)src_nhwc, (float4 *)dst_nchw, N, H, W / 4, scale);
  else if (v4 && C == 4) launch_pdl(k_image_to_tensor_v4<4>, dim3(ew_blocks(groups)), dim3(256), 0, st, (const uint32_t *)src_nhwc, (float4 *)dst_nchw, N, H, W / 4, scale);
  else launch_pdl(k_image_to_tensor, dim3(ew_blocks((long long)N * C * H * W)), dim3(256), 0, st, src_nhwc, dst_nchw, N, H, W, C, scale);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

int srb_img_interp_bicubic(const float *x, float *y, int32_t N, int32_t C, int32_t H, int32_t W, int32_t TH, int32_t TW,
                           const int32_t *bounds_w, const int32_t *coeffs_w, const int32_t *bounds_h, const int32_t *coeffs_h,
                           int32_t ksize, int32_t shave, void *stream) {
  SRB_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0 && TH > 2 * shave && TW > 2 * shave && shave >= 0 && ksize > 0, SRB_EINVAL,
              "bad img_interp sizes");
  if (N == 0) return SRB_OK;  // empty batch: nothing to read or write (the buffers may be null)
  SRB_REQUIRE(x && y && bounds_w && coeffs_w && bounds_h && coeffs_h, SRB_EINVAL, "null img_interp buffer");
  launch_pdl(k_pil_bicubic, dim3(ew_blocks((long long)N * C * (TH - 2 * shave) * (TW - 2 * shave))), dim3(256), 0, (cudaStream_t)stream, 
      x, y, N, C, H, W, TH, TW, bounds_w, coeffs_w, bounds_h, coeffs_h, ksize, shave);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

int srb_round_tf32(const float *x, float *y, int64_t n, void *stream) {
  SRB_REQUIRE(x && y && n >= 0, SRB_EINVAL, "bad round args");
  if (n == 0) return SRB_OK;
  launch_pdl(k_round_tf32, dim3(ew_blocks(n)), dim3(256), 0, (cudaStream_t)stream, x, y, n);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

}  // extern "C"
