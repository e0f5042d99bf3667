/*
 * TEST INFRASTRUCTURE ONLY -- plain-C restatement of the arithmetic behind the hot path, independent
 * of torch.  The reference (pure Python) reaches this arithmetic through torch.nn modules:
 *   Conv2d            base_networks.py:42,66        y[n,o,p,q] = b[o] + sum_{c,r,s} x[n,c,p*st-pad+r,q*st-pad+s] w[o,c,r,s]
 *   ConvTranspose2d   base_networks.py:77, fsrcnn.py:33   (adjoint of the above w.r.t. x, weight (Cin,Cout,kh,kw))
 *   PixelShuffle      base_networks.py:157          out[n,c,h*r+i,w*r+j] = in[n,c*r*r+i*r+j,h,w]
 *   PReLU / ReLU / LeakyReLU  base_networks.py:51-56      y = x > 0 ? x : a*x
 * These are the published ATen definitions (torch 2.11 docs: nn.Conv2d, nn.ConvTranspose2d,
 * nn.PixelShuffle, nn.PReLU); tests/test_oracle.py checks this file against torch on CPU.
 * All tensors NCHW contiguous fp32, accumulation in double.  Build: make -C oracle  (-> oracle/_build/).
 */
#include <stddef.h>

#define IDX4(n, c, h, w, C, H, W) ((((size_t)(n) * (C) + (c)) * (H) + (h)) * (W) + (w))

void ref_conv2d_fwd(const float *x, const float *w, const float *b, float *y, int N, int C, int H, int W, int O,
                    int kh, int kw, int st, int pad) {
  int Ho = (H + 2 * pad - kh) / st + 1, Wo = (W + 2 * pad - kw) / st + 1;
  for (int n = 0; n < N; ++n)
    for (int o = 0; o < O; ++o)
      for (int p = 0; p < Ho; ++p)
        for (int q = 0; q < Wo; ++q) {
          double acc = b ? b[o] : 0.0;
          for (int c = 0; c < C; ++c)
            for (int r = 0; r < kh; ++r) {
              int iy = p * st - pad + r;
              if (iy < 0 || iy >= H) continue;
              for (int s = 0; s < kw; ++s) {
                int ix = q * st - pad + s;
                if (ix < 0 || ix >= W) continue;
                acc += (double)x[IDX4(n, c, iy, ix, C, H, W)] * w[IDX4(o, c, r, s, C, kh, kw)];
              }
            }
          y[IDX4(n, o, p, q, O, Ho, Wo)] = (float)acc;
        }
}

/* dx = adjoint of conv w.r.t. x applied to dy (N,O,Ho,Wo); dx has shape (N,C,H,W). */
void ref_conv2d_bwd_data(const float *dy, const float *w, float *dx, int N, int C, int H, int W, int O, int kh,
                         int kw, int st, int pad, int Ho, int Wo) {
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c)
      for (int iy = 0; iy < H; ++iy)
        for (int ix = 0; ix < W; ++ix) {
          double acc = 0.0;
          for (int o = 0; o < O; ++o)
            for (int r = 0; r < kh; ++r) {
              int t = iy + pad - r;
              if (t < 0 || t % st) continue;
              int p = t / st;
              if (p >= Ho) continue;
              for (int s = 0; s < kw; ++s) {
                int u = ix + pad - s;
                if (u < 0 || u % st) continue;
                int q = u / st;
                if (q >= Wo) continue;
                acc += (double)dy[IDX4(n, o, p, q, O, Ho, Wo)] * w[IDX4(o, c, r, s, C, kh, kw)];
              }
            }
          dx[IDX4(n, c, iy, ix, C, H, W)] = (float)acc;
        }
}

void ref_conv2d_bwd_weight(const float *x, const float *dy, float *dw, float *db, int N, int C, int H, int W, int O,
                           int kh, int kw, int st, int pad) {
  int Ho = (H + 2 * pad - kh) / st + 1, Wo = (W + 2 * pad - kw) / st + 1;
  for (int o = 0; o < O; ++o) {
    for (int c = 0; c < C; ++c)
      for (int r = 0; r < kh; ++r)
        for (int s = 0; s < kw; ++s) {
          double acc = 0.0;
          for (int n = 0; n < N; ++n)
            for (int p = 0; p < Ho; ++p) {
              int iy = p * st - pad + r;
              if (iy < 0 || iy >= H) continue;
              for (int q = 0; q < Wo; ++q) {
                int ix = q * st - pad + s;
                if (ix < 0 || ix >= W) continue;
                acc += (double)dy[IDX4(n, o, p, q, O, Ho, Wo)] * x[IDX4(n, c, iy, ix, C, H, W)];
              }
            }
          dw[IDX4(o, c, r, s, C, kh, kw)] = (float)acc;
        }
    if (db) {
      double acc = 0.0;
      for (int n = 0; n < N; ++n)
        for (int p = 0; p < Ho; ++p)
          for (int q = 0; q < Wo; ++q) acc += dy[IDX4(n, o, p, q, O, Ho, Wo)];
      db[o] = (float)acc;
    }
  }
}

/* ConvTranspose2d forward: x (N,Ci,H,W), w (Ci,Co,kh,kw) -> y (N,Co,Ho,Wo), Ho = (H-1)*st - 2*pad + kh + out_pad.
 * It is ref_conv2d_bwd_data of a Conv2d(Co -> Ci) whose "dy" is x. */
void ref_conv_transpose2d_fwd(const float *x, const float *w, const float *b, float *y, int N, int Ci, int H, int W,
                              int Co, int kh, int kw, int st, int pad, int out_pad) {
  int Ho = (H - 1) * st - 2 * pad + kh + out_pad, Wo = (W - 1) * st - 2 * pad + kw + out_pad;
  ref_conv2d_bwd_data(x, w, y, N, Co, Ho, Wo, Ci, kh, kw, st, pad, H, W);
  if (b)
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < Co; ++c)
        for (int i = 0; i < Ho * Wo; ++i) y[((size_t)n * Co + c) * Ho * Wo + i] += b[c];
}

void ref_pixel_shuffle(const float *in, float *out, int N, int C, int H, int W, int r) {
  /* in (N, C*r*r, H, W) -> out (N, C, H*r, W*r) */
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c)
      for (int h = 0; h < H; ++h)
        for (int i = 0; i < r; ++i)
          for (int w = 0; w < W; ++w)
            for (int j = 0; j < r; ++j)
              out[IDX4(n, c, h * r + i, w * r + j, C, H * r, W * r)] =
                  in[IDX4(n, c * r * r + i * r + j, h, w, C * r * r, H, W)];
}

/* act: 1 relu, 2 prelu(a), 3 lrelu(a) */
void ref_act_fwd(const float *x, float *y, size_t n, int act, float a) {
  for (size_t i = 0; i < n; ++i) {
    float v = x[i];
    y[i] = act == 1 ? (v > 0.f ? v : 0.f) : (v > 0.f ? v : a * v);
  }
}

/* dx = dy * act'(x); returns d(alpha) = sum dy * x * [x <= 0] (PReLU) */
double ref_act_bwd(const float *x, const float *dy, float *dx, size_t n, int act, float a) {
  double da = 0.0;
  for (size_t i = 0; i < n; ++i) {
    float v = x[i], g = dy[i];
    if (act == 1) dx[i] = v > 0.f ? g : 0.f;
    else {
      dx[i] = v > 0.f ? g : a * g;
      if (!(v > 0.f)) da += (double)g * v;
    }
  }
  return da;
}
