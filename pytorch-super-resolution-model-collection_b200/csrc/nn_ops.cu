// The non-convolution layers around the conv stacks of SRGAN: train-mode BatchNorm2d with the
// activation and the residual add fused (base_networks.py:46,64,117,137,145), Linear (DenseBlock, base_networks.py:4-36 /
// srgan.py:66-70), MaxPool2d(2) (VGG19 features[4], srgan.py:84-90) and BCELoss (srgan.py:157,276-297).
// All of them are HBM- or L2-bound streaming work on fp32 data: CUDA-core kernels with coalesced, vectorised accesses and
// deterministic (fixed-order) reductions; no tensor-core reshaping (a 16-row GEMM has 8 flop/byte).
#include "srb_common.cuh"

namespace srb {

namespace {

inline unsigned nblocks(long long n, int per) {
  long long b = (n + per - 1) / per;
  if (b > 148LL * 8) b = 148LL * 8;
  if (b < 1) b = 1;
  return (unsigned)b;
}

__device__ __forceinline__ float act_fwd(float z, int act, float slope) {
  if (act == SRB_ACT_NONE) return z;
  if (act == SRB_ACT_RELU) return z > 0.f ? z : 0.f;
  return z > 0.f ? z : z * slope;
}
__device__ __forceinline__ float act_grad(float z, int act, float slope) {
  if (act == SRB_ACT_NONE) return 1.f;
  if (act == SRB_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  return z > 0.f ? 1.f : slope;
}

// ---------------------------------------------------------------------------------------------------------------------
// BatchNorm2d, training mode, dense NHWC input (P pixels x C channels).  Block = 64 channels x 4 pixel lanes.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bn_stats(const float *__restrict__ x, long long P, int C, double *__restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  __shared__ double sh[2][4][64];
  const int tc = threadIdx.x & 63, pl = threadIdx.x >> 6;
  const int c = blockIdx.y * 64 + tc;
  double s = 0.0, q = 0.0;
  if (c < C)
    for (long long p = (long long)blockIdx.x * 4 + pl; p < P; p += (long long)gridDim.x * 4) {
      const double v = (double)__ldg(x + p * C + c);
      s += v;
      q += v * v;
    }
  sh[0][pl][tc] = s;
  sh[1][pl][tc] = q;
  __syncthreads();
  if (pl == 0 && c < C) {
    s = (sh[0][0][tc] + sh[0][1][tc]) + (sh[0][2][tc] + sh[0][3][tc]);
    q = (sh[1][0][tc] + sh[1][1][tc]) + (sh[1][2][tc] + sh[1][3][tc]);
    partial[((size_t)blockIdx.x * 2 + 0) * C + c] = s;
    partial[((size_t)blockIdx.x * 2 + 1) * C + c] = q;
  }
}

// Column sums of the [gx][2][C] double partials: block = 32 channels (x) x 16 row lanes (y).  Lane y adds rows y, y+16, ... with four
// independent accumulators (the one-thread-per-channel loop this replaces walked up to 592 dependent L2 round trips: 52 us);
// the 16 lanes are folded through shared memory in a fixed order, so the result does not depend on scheduling.
constexpr int kBnFinLanes = 16;
__device__ __forceinline__ void bn_partial_sums(const double *__restrict__ partial, int gx, int C, int c, double &s_out, double &q_out) {
  __shared__ double fold[2][kBnFinLanes][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  double s0 = 0.0, s1 = 0.0, q0 = 0.0, q1 = 0.0;
  if (c < C) {
    int b = ty;
    for (; b + kBnFinLanes < gx; b += 2 * kBnFinLanes) {
      s0 += partial[((size_t)b * 2 + 0) * C + c];
      q0 += partial[((size_t)b * 2 + 1) * C + c];
      s1 += partial[((size_t)(b + kBnFinLanes) * 2 + 0) * C + c];
      q1 += partial[((size_t)(b + kBnFinLanes) * 2 + 1) * C + c];
    }
    if (b < gx) {
      s0 += partial[((size_t)b * 2 + 0) * C + c];
      q0 += partial[((size_t)b * 2 + 1) * C + c];
    }
  }
  fold[0][ty][tx] = s0 + s1;
  fold[1][ty][tx] = q0 + q1;
  __syncthreads();
  double s = 0.0, q = 0.0;
  if (ty == 0) {
#pragma unroll
    for (int y = 0; y < kBnFinLanes; ++y) { s += fold[0][y][tx]; q += fold[1][y][tx]; }
  }
  s_out = s;
  q_out = q;
}

// mean / biased variance from the partials (fixed order), running statistics updated like nn.BatchNorm2d (unbiased variance)
__global__ void k_bn_finalize(const double *__restrict__ partial, int gx, int C, double count, float momentum, float eps,
                              float *running_mean, float *running_var, float *save_mean, float *save_invstd) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s, q;
  bn_partial_sums(partial, gx, C, c, s, q);
  if (c >= C || threadIdx.y != 0) return;
  const double mean = s / count;
  double var = q / count - mean * mean;
  if (var < 0.0) var = 0.0;
  save_mean[c] = (float)mean;
  save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
  }
}

// eval mode: statistics are the running ones
__global__ void k_bn_eval_stats(const float *__restrict__ running_mean, const float *__restrict__ running_var, int C, float eps,
                                float *save_mean, float *save_invstd) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  save_mean[c] = running_mean[c];
  save_invstd[c] = 1.0f / sqrtf(running_var[c] + eps);
}

// y = act(gamma * (x - mean) * invstd + beta) + residual, float4 over channels (C % 4 == 0)
__global__ void __launch_bounds__(256) k_bn_apply(const float4 *__restrict__ x, float4 *__restrict__ y, long long n4, int C4,
                                                  const float4 *__restrict__ mean, const float4 *__restrict__ invstd,
                                                  const float4 *__restrict__ gamma, const float4 *__restrict__ beta, int act,
                                                  float slope_in, const float *__restrict__ alpha,
                                                  const float4 *__restrict__ residual, int rnd) {
  pdl_trigger();
  pdl_wait();
  const float slope = act == SRB_ACT_PRELU ? __ldg(alpha) : slope_in;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    const float4 v = __ldg(x + i), m = __ldg(mean + c4), is = __ldg(invstd + c4), g = __ldg(gamma + c4), b = __ldg(beta + c4);
    float4 o;
    o.x = act_fwd((v.x - m.x) * is.x * g.x + b.x, act, slope);
    o.y = act_fwd((v.y - m.y) * is.y * g.y + b.y, act, slope);
    o.z = act_fwd((v.z - m.z) * is.z * g.z + b.z, act, slope);
    o.w = act_fwd((v.w - m.w) * is.w * g.w + b.w, act, slope);
    if (residual) {
      const float4 r = __ldg(residual + i);
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (rnd) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
    y[i] = o;
  }
}

// backward pass 1: per channel sum(dz), sum(dz * xhat) with dz = dy * act'(z), z = gamma * xhat + beta; PReLU: sum(dy * z * [z<=0])
__global__ void __launch_bounds__(256) k_bn_bwd_stats(const float *__restrict__ x, const float *__restrict__ dy, long long P, int C,
                                                      const float *__restrict__ mean, const float *__restrict__ invstd,
                                                      const float *__restrict__ gamma, const float *__restrict__ beta, int act,
                                                      float slope_in, const float *__restrict__ alpha, double *__restrict__ partial,
                                                      double *__restrict__ alpha_partial) {
  pdl_trigger();
  pdl_wait();
  __shared__ double sh[3][4][64];
  const float slope = act == SRB_ACT_PRELU ? __ldg(alpha) : slope_in;
  const int tc = threadIdx.x & 63, pl = threadIdx.x >> 6;
  const int c = blockIdx.y * 64 + tc;
  double s = 0.0, q = 0.0, da = 0.0;
  if (c < C) {
    const float m = __ldg(mean + c), is = __ldg(invstd + c), g = __ldg(gamma + c), b = __ldg(beta + c);
    for (long long p = (long long)blockIdx.x * 4 + pl; p < P; p += (long long)gridDim.x * 4) {
      const float xh = (__ldg(x + p * C + c) - m) * is;
      const float z = g * xh + b;
      const float gy = __ldg(dy + p * C + c);
      const float dz = gy * act_grad(z, act, slope);
      s += (double)dz;
      q += (double)dz * (double)xh;
      if (act == SRB_ACT_PRELU && !(z > 0.f)) da += (double)gy * (double)z;
    }
  }
  sh[0][pl][tc] = s; sh[1][pl][tc] = q; sh[2][pl][tc] = da;
  __syncthreads();
  if (pl == 0) {
    s = (sh[0][0][tc] + sh[0][1][tc]) + (sh[0][2][tc] + sh[0][3][tc]);
    q = (sh[1][0][tc] + sh[1][1][tc]) + (sh[1][2][tc] + sh[1][3][tc]);
    da = (sh[2][0][tc] + sh[2][1][tc]) + (sh[2][2][tc] + sh[2][3][tc]);
    if (c < C) {
      partial[((size_t)blockIdx.x * 2 + 0) * C + c] = s;
      partial[((size_t)blockIdx.x * 2 + 1) * C + c] = q;
    }
    sh[2][0][tc] = c < C ? da : 0.0;
  }
  __syncthreads();
  if (threadIdx.x == 0 && alpha_partial) {
    double t = 0.0;
    for (int i = 0; i < 64; ++i) t += sh[2][0][i];
    alpha_partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}

__global__ void k_bn_bwd_finalize(const double *__restrict__ partial, int gx, int C, double count, float *dgamma, float *dbeta,
                                  float *mean_dz, float *mean_dzx, const double *__restrict__ alpha_partial, int n_alpha,
                                  float *dalpha, float scale, int accumulate) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s, q;
  bn_partial_sums(partial, gx, C, c, s, q);
  if (threadIdx.y != 0) return;
  if (c < C) {
    mean_dz[c] = (float)(s / count);
    mean_dzx[c] = (float)(q / count);
    const float g = (float)q * scale, b2 = (float)s * scale;
    dgamma[c] = accumulate ? dgamma[c] + g : g;
    dbeta[c] = accumulate ? dbeta[c] + b2 : b2;
  }
  if (c == 0 && dalpha && alpha_partial) {
    double t = 0.0;
    for (int i = 0; i < n_alpha; ++i) t += alpha_partial[i];
    *dalpha += (float)t;
  }
}

// backward pass 2: dx = gamma * invstd * (dz - mean(dz) - xhat * mean(dz * xhat))
__global__ void __launch_bounds__(256) k_bn_bwd_apply(const float4 *__restrict__ x, const float4 *__restrict__ dy, float4 *__restrict__ dx,
                                                      long long n4, int C4, const float4 *__restrict__ mean,
                                                      const float4 *__restrict__ invstd, const float4 *__restrict__ gamma,
                                                      const float4 *__restrict__ beta, const float4 *__restrict__ mean_dz,
                                                      const float4 *__restrict__ mean_dzx, int act, float slope_in,
                                                      const float *__restrict__ alpha, int rnd) {
  pdl_trigger();
  pdl_wait();
  const float slope = act == SRB_ACT_PRELU ? __ldg(alpha) : slope_in;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    const float4 v = __ldg(x + i), gy = __ldg(dy + i), m = __ldg(mean + c4), is = __ldg(invstd + c4), g = __ldg(gamma + c4),
                 b = __ldg(beta + c4), a1 = __ldg(mean_dz + c4), a2 = __ldg(mean_dzx + c4);
    float4 o;
#define SRB_BN_BWD1(f)                                                            \
    {                                                                             \
      const float xh = (v.f - m.f) * is.f;                                        \
      const float dz = gy.f * act_grad(g.f * xh + b.f, act, slope);               \
      o.f = g.f * is.f * (dz - a1.f - xh * a2.f);                                 \
    }
    SRB_BN_BWD1(x) SRB_BN_BWD1(y) SRB_BN_BWD1(z) SRB_BN_BWD1(w)
#undef SRB_BN_BWD1
    if (rnd) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
    dx[i] = o;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Linear: y[b, o] = sum_i x[b, i] * w[o, i] + bias[o]   (B <= 16 rows per launch; the weight matrix is streamed once)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kLinB = 16;    // batch rows per launch
constexpr int kLinKC = 512;  // input features staged per step (16 x 512 floats = 32 KB)

// forward partials: block = 8 warps x 4 outputs; grid (O / 32, K splits); partial[ks][b][o]
__global__ void __launch_bounds__(256) k_linear_fwd(const float *__restrict__ x, const float *__restrict__ w, int B, int I, int O,
                                                    int k_per_split, float *__restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  __shared__ float4 xs[kLinB][kLinKC / 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o0 = blockIdx.x * 32 + warp * 4;
  const int k0 = blockIdx.y * k_per_split, k1 = min(I, k0 + k_per_split);
  float acc[4][kLinB];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int b = 0; b < kLinB; ++b) acc[j][b] = 0.f;
  for (int kc = k0; kc < k1; kc += kLinKC) {
    const int kn = min(kLinKC, k1 - kc);  // multiple of 4 (host guarantees I % 4 == 0 and split sizes % 4 == 0)
    __syncthreads();
    for (int e = threadIdx.x; e < kLinB * (kLinKC / 4); e += 256) {
      const int b = e / (kLinKC / 4), q = e - b * (kLinKC / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b < B && q * 4 < kn) v = __ldg((const float4 *)(x + (size_t)b * I + kc) + q);
      xs[b][q] = v;
    }
    __syncthreads();
    for (int q = lane; q * 4 < kn; q += 32) {
      float4 wv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        wv[j] = (o0 + j < O) ? __ldg((const float4 *)(w + (size_t)(o0 + j) * I + kc) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int b = 0; b < kLinB; ++b) {
        const float4 xv = xs[b][q];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          acc[j][b] += (wv[j].x * xv.x + wv[j].y * xv.y) + (wv[j].z * xv.z + wv[j].w * xv.w);
      }
    }
  }
  // fold the 32 lanes in a fixed (butterfly) order; lane 0 writes
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int b = 0; b < kLinB; ++b) {
      float v = acc[j][b];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && o0 + j < O && b < B) partial[((size_t)blockIdx.y * kLinB + b) * O + o0 + j] = v;
    }
}

// y = act(sum_splits partial + bias): act 0 none, 3 lrelu(slope), 4 sigmoid
__global__ void k_linear_fwd_finish(const float *__restrict__ partial, int splits, int B, int O, const float *__restrict__ bias,
                                    float *__restrict__ y) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * O) return;
  const int b = i / O, o = i - b * O;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += partial[((size_t)k * kLinB + b) * O + o];
  if (bias) s += __ldg(bias + o);
  y[i] = s;
}

// dx[b, i] = sum_o dy[b, o] * w[o, i]: thread = one float4 of i, all batch rows; grid (I / 1024, O splits); partial[os][b][i]
__global__ void __launch_bounds__(256) k_linear_dx(const float *__restrict__ dy, const float *__restrict__ w, int B, int I, int O,
                                                   int o_per_split, float *__restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float dys[];  // [o_per_split][kLinB]
  const int o0 = blockIdx.y * o_per_split, o1 = min(O, o0 + o_per_split);
  for (int e = threadIdx.x; e < (o1 - o0) * kLinB; e += 256) {
    const int oo = e / kLinB, b = e - oo * kLinB;
    dys[e] = b < B ? __ldg(dy + (size_t)b * O + o0 + oo) : 0.f;
  }
  __syncthreads();
  const long long i4 = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i4 * 4 >= I) return;
  float4 acc[kLinB];
#pragma unroll
  for (int b = 0; b < kLinB; ++b) acc[b] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int o = o0; o < o1; ++o) {
    const float4 wv = __ldg((const float4 *)(w + (size_t)o * I) + i4);
    const float *d = dys + (o - o0) * kLinB;
#pragma unroll
    for (int b = 0; b < kLinB; ++b) {
      const float g = d[b];
      acc[b].x += g * wv.x; acc[b].y += g * wv.y; acc[b].z += g * wv.z; acc[b].w += g * wv.w;
    }
  }
  for (int b = 0; b < B; ++b) *((float4 *)(partial + ((size_t)blockIdx.y * kLinB + b) * I) + i4) = acc[b];
}

__global__ void k_linear_dx_finish(const float4 *__restrict__ partial, int splits, int B, long long I4, float4 *__restrict__ dx) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * I4) return;
  const long long b = i / I4, q = i - b * I4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < splits; ++k) {
    const float4 v = partial[((size_t)k * kLinB + b) * I4 + q];
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  dx[i] = s;
}

// dw[o, i] (=|+=) scale * sum_b dy[b, o] * x[b, i]; db[o] likewise.  Block = 32 outputs x 1024 inputs (thread = one float4 of i)
__global__ void __launch_bounds__(256) k_linear_dw(const float *__restrict__ x, const float *__restrict__ dy, int B, int I, int O,
                                                   float *__restrict__ dw, float *__restrict__ db, float scale, int accumulate) {
  pdl_trigger();
  pdl_wait();
  __shared__ float dys[32][kLinB];
  const int o0 = blockIdx.y * 32;
  for (int e = threadIdx.x; e < 32 * kLinB; e += 256) {
    const int oo = e / kLinB, b = e - oo * kLinB;
    dys[oo][b] = (b < B && o0 + oo < O) ? __ldg(dy + (size_t)b * O + o0 + oo) * scale : 0.f;
  }
  __syncthreads();
  if (blockIdx.x == 0 && db && threadIdx.x < 32 && o0 + (int)threadIdx.x < O) {
    float s = 0.f;
    for (int b = 0; b < kLinB; ++b) s += dys[threadIdx.x][b];
    float *d = db + o0 + threadIdx.x;
    *d = accumulate ? *d + s : s;
  }
  const long long i4 = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i4 * 4 >= I) return;
  float4 xv[kLinB];
#pragma unroll
  for (int b = 0; b < kLinB; ++b) xv[b] = b < B ? __ldg((const float4 *)(x + (size_t)b * I) + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int oo = 0; oo < 32 && o0 + oo < O; ++oo) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int b = 0; b < kLinB; ++b) {
      const float g = dys[oo][b];
      s.x += g * xv[b].x; s.y += g * xv[b].y; s.z += g * xv[b].z; s.w += g * xv[b].w;
    }
    float4 *d = (float4 *)(dw + (size_t)(o0 + oo) * I) + i4;
    if (accumulate) { const float4 p = *d; s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w; }
    *d = s;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// MaxPool2d(kernel 2, stride 2), dense NHWC; the first maximum in (dy, dx) scan order wins (ATen)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void k_maxpool2_fwd(const float *__restrict__ x, float *__restrict__ y, unsigned char *__restrict__ idx, int N, int C,
                               int H, int W) {
  pdl_trigger();
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Ho * Wo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long q = i / C;
    const int ox = (int)(q % Wo); q /= Wo;
    const int oy = (int)(q % Ho);
    const long long n = q / Ho;
    const float *p = x + ((n * H + 2 * oy) * W + 2 * ox) * C + c;
    float best = __ldg(p);
    int bi = 0;
    const float v1 = __ldg(p + C), v2 = __ldg(p + (long long)W * C), v3 = __ldg(p + (long long)W * C + C);
    if (v1 > best || v1 != v1) { best = v1; bi = 1; }
    if (v2 > best || v2 != v2) { best = v2; bi = 2; }
    if (v3 > best || v3 != v3) { best = v3; bi = 3; }
    y[i] = best;
    idx[i] = (unsigned char)bi;
  }
}

__global__ void k_maxpool2_bwd(const float *__restrict__ dy, const unsigned char *__restrict__ idx, float *__restrict__ dx, int N, int C,
                               int H, int W) {
  pdl_trigger();
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Ho * Wo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long q = i / C;
    const int ox = (int)(q % Wo); q /= Wo;
    const int oy = (int)(q % Ho);
    const long long n = q / Ho;
    float *p = dx + ((n * H + 2 * oy) * W + 2 * ox) * C + c;
    const float g = __ldg(dy + i);
    const int bi = idx[i];
    p[0] = bi == 0 ? g : 0.f;
    p[C] = bi == 1 ? g : 0.f;
    p[(long long)W * C] = bi == 2 ? g : 0.f;
    p[(long long)W * C + C] = bi == 3 ? g : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// BCELoss (mean): loss = -mean(t * log(y) + (1 - t) * log(1 - y)), logs clamped at -100 (ATen); one block (n is a batch size)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bce_fwd(const float *__restrict__ y, const float *__restrict__ t, int n, float *loss) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float yy = y[i], tt = t[i];
    const float l1 = fmaxf(logf(yy), -100.f), l0 = fmaxf(logf(1.f - yy), -100.f);
    s -= tt * l1 + (1.f - tt) * l0;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = red[0] / (float)n;
}

__global__ void k_bce_bwd(const float *__restrict__ y, const float *__restrict__ t, int n, const float *__restrict__ g, float *dy) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float yy = y[i], tt = t[i];
  dy[i] = __ldg(g) * (yy - tt) / fmaxf((1.f - yy) * yy, 1e-12f) / (float)n;
}

}  // namespace
}  // namespace srb

using namespace srb;

extern "C" {

size_t srb_bn_workspace_bytes(int32_t C) {
  // stats partials (148*4 blocks x 2 x C doubles) + PReLU partials + the two per-channel means of the backward
  return (size_t)148 * 4 * 2 * (size_t)C * sizeof(double) + (size_t)148 * 4 * ((C + 63) / 64) * sizeof(double) + 4 * (size_t)C * sizeof(float) + 1024;
}

static inline unsigned bn_gx(long long P) {
  long long gx = (P + 63) / 64;  // >= 16 pixels per pixel lane
  if (gx > 148 * 4) gx = 148 * 4;
  if (gx < 1) gx = 1;
  return (unsigned)gx;
}

int srb_bn_fwd(const float *x, float *y, int64_t P, int32_t C, const float *gamma, const float *beta, float *running_mean,
               float *running_var, int training, float momentum, float eps, float *save_mean, float *save_invstd, int act,
               float slope, const float *alpha, const float *residual, int round_to_tf32, void *ws, size_t ws_bytes, void *stream) {
  SRB_REQUIRE(x && y && gamma && beta && save_mean && save_invstd && P > 0 && C > 0 && (C % 4) == 0, SRB_EINVAL,
              "bad batch-norm arguments (dense NHWC, C %% 4 == 0)");
  SRB_REQUIRE(act >= SRB_ACT_NONE && act <= SRB_ACT_LRELU && (act != SRB_ACT_PRELU || alpha), SRB_EINVAL, "bad activation");
  SRB_REQUIRE(ws && ws_bytes >= srb_bn_workspace_bytes(C), SRB_EWORKSPACE, "batch-norm workspace too small");
  SRB_REQUIRE(training || (running_mean && running_var), SRB_EINVAL, "eval-mode batch norm needs running statistics");
  cudaStream_t st = (cudaStream_t)stream;
  if (training) {
    double *partial = (double *)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    const unsigned gx = bn_gx(P);
    launch_pdl(k_bn_stats, dim3(dim3(gx, (C + 63) / 64)), dim3(256), 0, st, x, P, C, partial);
    launch_pdl(k_bn_finalize, dim3((C + 31) / 32), dim3(32, kBnFinLanes), 0, st, partial, (int)gx, C, (double)P, momentum, eps, running_mean, running_var, save_mean,
                                                   save_invstd);
    count_launch(2);
  } else {
    launch_pdl(k_bn_eval_stats, dim3((C + 127) / 128), dim3(128), 0, st, running_mean, running_var, C, eps, save_mean, save_invstd);
    count_launch();
  }
  const long long n4 = P * C / 4;
  launch_pdl(k_bn_apply, dim3(nblocks(n4, 256 * 4)), dim3(256), 0, st, (const float4 *)x, (float4 *)y, n4, C / 4, (const float4 *)save_mean,
                                                    (const float4 *)save_invstd, (const float4 *)gamma, (const float4 *)beta, act, slope,
                                                    alpha, (const float4 *)residual, round_to_tf32);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

int srb_bn_bwd(const float *x, const float *dy, float *dx, int64_t P, int32_t C, const float *gamma, const float *beta,
               const float *save_mean, const float *save_invstd, int act, float slope, const float *alpha, float *dgamma,
               float *dbeta, float *dalpha, float scale, int accumulate, int round_to_tf32, void *ws, size_t ws_bytes, void *stream) {
  SRB_REQUIRE(x && dy && dx && gamma && beta && save_mean && save_invstd && dgamma && dbeta && P > 0 && C > 0 && (C % 4) == 0, SRB_EINVAL,
              "bad batch-norm arguments (dense NHWC, C %% 4 == 0)");
  SRB_REQUIRE(act != SRB_ACT_PRELU || alpha, SRB_EINVAL, "PReLU needs alpha");
  SRB_REQUIRE(ws && ws_bytes >= srb_bn_workspace_bytes(C), SRB_EWORKSPACE, "batch-norm workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned gx = bn_gx(P), gy = (unsigned)((C + 63) / 64);
  double *partial = (double *)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  double *apart = partial + (size_t)148 * 4 * 2 * C;
  float *means = (float *)(apart + (size_t)148 * 4 * gy);
  launch_pdl(k_bn_bwd_stats, dim3(dim3(gx, gy)), dim3(256), 0, st, x, dy, P, C, save_mean, save_invstd, gamma, beta, act, slope, alpha, partial,
                                               (act == SRB_ACT_PRELU && dalpha) ? apart : nullptr);
  launch_pdl(k_bn_bwd_finalize, dim3((C + 31) / 32), dim3(32, kBnFinLanes), 0, st, partial, (int)gx, C, (double)P, dgamma, dbeta, means, means + C,
                                                     (act == SRB_ACT_PRELU && dalpha) ? apart : nullptr, (int)(gx * gy), dalpha, scale,
                                                     accumulate);
  const long long n4 = P * C / 4;
  launch_pdl(k_bn_bwd_apply, dim3(nblocks(n4, 256 * 4)), dim3(256), 0, st, (const float4 *)x, (const float4 *)dy, (float4 *)dx, n4, C / 4,
                                                        (const float4 *)save_mean, (const float4 *)save_invstd, (const float4 *)gamma,
                                                        (const float4 *)beta, (const float4 *)means, (const float4 *)(means + C), act,
                                                        slope, alpha, round_to_tf32);
  count_launch(3);
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

size_t srb_linear_workspace_bytes(int32_t I, int32_t O) {
  // forward: up to 16 K-splits x 16 rows x O; dx: up to 8 O-splits x 16 rows x I
  const size_t a = (size_t)16 * 16 * (size_t)O * sizeof(float), b = (size_t)8 * 16 * (size_t)I * sizeof(float);
  return (a > b ? a : b) + 1024;
}

int srb_linear_fwd(const float *x, const float *w, const float *bias, float *y, int32_t B, int32_t I, int32_t O, void *ws,
                   size_t ws_bytes, void *stream) {
  SRB_REQUIRE(x && w && y && B > 0 && I > 0 && O > 0 && (I % 4) == 0, SRB_EINVAL, "bad linear arguments (I %% 4 == 0)");
  SRB_REQUIRE(ws && ws_bytes >= srb_linear_workspace_bytes(I, O), SRB_EWORKSPACE, "linear workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float *partial = (float *)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  const int otiles = (O + 31) / 32;
  int splits = 296 / otiles;  // ~2 blocks per SM
  if (splits < 1) splits = 1;
  if (splits > 16) splits = 16;
  int kps = ((I + splits - 1) / splits + kLinKC - 1) / kLinKC * kLinKC;
  splits = (I + kps - 1) / kps;
  for (int b0 = 0; b0 < B; b0 += kLinB) {
    const int bn = B - b0 < kLinB ? B - b0 : kLinB;
    launch_pdl(k_linear_fwd, dim3(dim3(otiles, splits)), dim3(256), 0, st, x + (size_t)b0 * I, w, bn, I, O, kps, partial);
    launch_pdl(k_linear_fwd_finish, dim3((bn * O + 255) / 256), dim3(256), 0, st, partial, splits, bn, O, bias, y + (size_t)b0 * O);
    count_launch(2);
  }
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

int srb_linear_bwd(const float *x, const float *w, const float *dy, float *dx, float *dw, float *db, int32_t B, int32_t I, int32_t O,
                   float scale, int accumulate, void *ws, size_t ws_bytes, void *stream) {
  SRB_REQUIRE(x && w && dy && B > 0 && I > 0 && O > 0 && (I % 4) == 0, SRB_EINVAL, "bad linear arguments (I %% 4 == 0)");
  SRB_REQUIRE(ws && ws_bytes >= srb_linear_workspace_bytes(I, O), SRB_EWORKSPACE, "linear workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float *partial = (float *)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  for (int b0 = 0; b0 < B; b0 += kLinB) {
    const int bn = B - b0 < kLinB ? B - b0 : kLinB;
    if (dx) {
      const int itiles = (I / 4 + 255) / 256;
      int splits = 296 / itiles;
      if (splits < 1) splits = 1;
      if (splits > 8) splits = 8;
      int ops = (O + splits - 1) / splits;
      if (ops > 512) ops = 512;  // shared memory: ops x 16 floats
      splits = (O + ops - 1) / ops;
      SRB_REQUIRE(splits <= 8, SRB_EUNSUPPORTED, "linear backward: more than 4096 output features");
      launch_pdl(k_linear_dx, dim3(dim3(itiles, splits)), dim3(256), (size_t)ops * kLinB * sizeof(float), st, dy + (size_t)b0 * O, w, bn, I, O, ops, partial);
      launch_pdl(k_linear_dx_finish, dim3((unsigned)(((long long)bn * (I / 4) + 255) / 256)), dim3(256), 0, st, (const float4 *)partial, splits, bn, I / 4,
                                                                                         (float4 *)(dx + (size_t)b0 * I));
      count_launch(2);
    }
    if (dw) {
      launch_pdl(k_linear_dw, dim3(dim3((I / 4 + 255) / 256, (O + 31) / 32)), dim3(256), 0, st, x + (size_t)b0 * I, dy + (size_t)b0 * O, bn, I, O, dw, db, scale,
                                                                           (accumulate || b0 > 0) ? 1 : 0);
      count_launch();
    }
  }
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

int srb_maxpool2_fwd(const float *x, float *y, uint8_t *idx, int32_t N, int32_t C, int32_t H, int32_t W, void *stream) {
  SRB_REQUIRE(x && y && idx && N > 0 && C > 0 && H >= 2 && W >= 2, SRB_EINVAL, "bad maxpool arguments");
  launch_pdl(k_maxpool2_fwd, dim3(nblocks((long long)N * (H / 2) * (W / 2) * C, 256)), dim3(256), 0, (cudaStream_t)stream, x, y, idx, N, C, H, W);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

int srb_maxpool2_bwd(const float *dy, const uint8_t *idx, float *dx, int32_t N, int32_t C, int32_t H, int32_t W, void *stream) {
  SRB_REQUIRE(dy && dx && idx && N > 0 && C > 0 && H >= 2 && W >= 2 && (H % 2) == 0 && (W % 2) == 0, SRB_EINVAL,
              "bad maxpool arguments (even H, W)");
  launch_pdl(k_maxpool2_bwd, dim3(nblocks((long long)N * (H / 2) * (W / 2) * C, 256)), dim3(256), 0, (cudaStream_t)stream, dy, idx, dx, N, C, H, W);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

int srb_bce_fwd(const float *y, const float *t, int32_t n, float *loss, void *stream) {
  SRB_REQUIRE(y && t && loss && n > 0, SRB_EINVAL, "bad BCE arguments");
  launch_pdl(k_bce_fwd, dim3(1), dim3(256), 0, (cudaStream_t)stream, y, t, n, loss);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

int srb_bce_bwd(const float *y, const float *t, int32_t n, const float *grad_loss, float *dy, void *stream) {
  SRB_REQUIRE(y && t && grad_loss && dy && n > 0, SRB_EINVAL, "bad BCE arguments");
  launch_pdl(k_bce_bwd, dim3((n + 255) / 256), dim3(256), 0, (cudaStream_t)stream, y, t, n, grad_loss, dy);
  count_launch();
  SRB_CHECK_CUDA(cudaGetLastError());
  return SRB_OK;
}

}  // extern "C"
