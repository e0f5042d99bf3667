// Strided and transposed convolutions on the tcgen05 slot-linear kernel (tc_conv_sl.cu) by PHASE DECOMPOSITION.
//
// gather geometry (srb_common.cuh):  small[n,co,oy,ox] = sum big[n,ci, oy*st - pad + r, ox*st - pad + s] * w[co,ci,r,s]
//
// (1) tc_strided_gather -- computes `small` from `big` (Conv2d forward with stride st; ConvTranspose2d backward-data):
//     write r - pad = st*d + a, a in [0,st): the tap reads phase image P_a[i] = big[st*i + a] at i = oy + d.  All st*st phase
//     images become extra input CHANNELS (k = (a*st+b)*Ci + c) of ONE stride-1 convolution with a ceil-sized kernel over d;
//     the phase images are never materialised -- the A-operand TMA map walks `big` with element strides (1, st, st, 1) -- and
//     (phase, tap) pairs that do not exist in the original filter are skipped through per-phase tap masks (k3 s2: 9 of 16).
// (2) tc_strided_scatter -- computes `big` from `small` (Conv2d backward-data with stride st; ConvTranspose2d forward):
//     output phase (a,b) (big[st*i + a, st*j + b]) only receives taps r = ra + st*t, ra = (a + pad) mod st: it is a stride-1
//     convolution of `small` with the sub-sampled, flipped filter; st*st launches (1 + 2 + 2 + 4 = 9 taps in total for k3 s2,
//     none wasted), each storing through a T4 view of `big` whose spatial strides are multiplied by st.
// Weight gradients of these layers stay on the CUDA-core wgrad (simt_conv.cu).
#include "srb_common.cuh"

namespace srb {

namespace {

inline int floordiv(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

struct GatherPhase {
  Geom g2;       // the equivalent stride-1 launch
  ConvOpt opt;
};

bool plan_gather(const Geom &g, GatherPhase *gp) {
  if (g.st < 2 || g.st > 4 || g.ps != 1) return false;
  const int st = g.st;
  const int dmin_r = floordiv(-g.pad, st), dmax_r = floordiv(g.kh - 1 - g.pad, st);
  const int dmin_s = floordiv(-g.pad, st), dmax_s = floordiv(g.kw - 1 - g.pad, st);
  const int kh2 = dmax_r - dmin_r + 1, kw2 = dmax_s - dmin_s + 1;
  if (kh2 * kw2 > 16 || dmin_r != dmin_s) return false;
  Geom g2 = g;
  g2.Ci = g.Ci * st * st;
  g2.Hi = (g.Hi + st - 1) / st;
  g2.Wi = (g.Wi + st - 1) / st;
  g2.kh = kh2; g2.kw = kw2;
  g2.st = 1;
  g2.pad = -dmin_r;
  gp->g2 = g2;
  ConvOpt o;
  o.in_ps = st; o.in_h = g.Hi; o.in_w = g.Wi;
  o.wmode = 1; o.st = st; o.pad0 = g.pad;  // r = st*(tr + dmin) + a + pad
  o.kh0 = g.kh; o.kw0 = g.kw; o.dmin_r = dmin_r; o.dmin_s = dmin_s;
  gp->opt = o;
  return true;
}

struct ScatterPhase {
  Geom g2;
  ConvOpt opt;
  int a, b;
  bool empty;  // no tap reaches this output phase (kernel smaller than the stride): the phase is bias / zero only
};

// phase (a, b) of the scatter: stride-1 gather of `small` (g.Co channels, g.Ho x g.Wo) into big phase (g.Ci channels)
bool plan_scatter_phase(const Geom &g, int a, int b, ScatterPhase *sp) {
  const int st = g.st;
  const int ra = (a + g.pad) % st, rb = (b + g.pad) % st;
  const int da = (a + g.pad) / st, db = (b + g.pad) / st;
  const int kt_a = ra < g.kh ? (g.kh - 1 - ra) / st + 1 : 0, kt_b = rb < g.kw ? (g.kw - 1 - rb) / st + 1 : 0;
  const int Hp = a < g.Hi ? (g.Hi - a + st - 1) / st : 0, Wp = b < g.Wi ? (g.Wi - b + st - 1) / st : 0;
  sp->a = a; sp->b = b;
  sp->empty = kt_a == 0 || kt_b == 0;
  if (Hp == 0 || Wp == 0) { sp->g2 = Geom{g.N, g.Co, g.Ho, g.Wo, g.Ci, 0, 0, 1, 1, 1, 0, 1}; return true; }
  const int tmax_a = kt_a - 1, tmax_b = kt_b - 1;
  const int pad_h = tmax_a - da, pad_w = tmax_b - db;
  if (!sp->empty && (pad_h < 0 || pad_w < 0)) return false;
  sp->g2 = Geom{g.N, g.Co, g.Ho, g.Wo, g.Ci, Hp, Wp, sp->empty ? 1 : kt_a, sp->empty ? 1 : kt_b, 1, sp->empty ? 0 : pad_h, 1};
  ConvOpt o;
  o.wmode = 2; o.st = st; o.kh0 = g.kh; o.kw0 = g.kw;
  o.ra = ra; o.rb = rb; o.tmax_a = tmax_a; o.tmax_b = tmax_b;
  o.pad_w = sp->empty ? 0 : pad_w;
  sp->opt = o;
  return true;
}

inline T4 phase_view(const T4 &t, int st, int a, int b) {
  T4 v = t;
  if (!t.p) return v;
  const long long es = t.dt == SRB_BF16 ? 2 : 4;
  v.p = (float *)((char *)t.p + ((long long)a * t.sh + (long long)b * t.sw) * es);
  v.sh = t.sh * st;
  v.sw = t.sw * st;
  return v;
}

}  // namespace

bool tc_strided_gather_supported(const Geom &g, const T4 &big, const T4 &small) {
  GatherPhase gp;
  if (!plan_gather(g, &gp)) return false;
  if (big.dt != SRB_F32 || small.dt != SRB_F32) return false;
  return tc_conv_supported(gp.g2, big, small, false, g.st);
}

int tc_strided_gather(const Geom &g, const T4 &big, const float *w, const T4 &small, const Epi &epi, void *ws, size_t ws_bytes,
                      cudaStream_t st) {
  GatherPhase gp;
  SRB_REQUIRE(plan_gather(g, &gp), SRB_EUNSUPPORTED, "strided conv: no phase plan");
  return tc_conv_gather(gp.g2, big, w, false, small, epi, ws, ws_bytes, st, gp.opt);
}

bool tc_strided_scatter_supported(const Geom &g, const T4 &small, const T4 &big) {
  if (g.st < 2 || g.st > 4 || g.ps != 1) return false;
  if (big.dt != SRB_F32 || small.dt != SRB_F32) return false;
  if (g.Co <= 4) return false;  // the launches' input is `small`: the Cin <= 4 operand flavour has no phase filter mapping
  for (int a = 0; a < g.st; ++a)
    for (int b = 0; b < g.st; ++b) {
      ScatterPhase sp;
      if (!plan_scatter_phase(g, a, b, &sp)) return false;
      if (sp.g2.Ho == 0) continue;
      if (sp.empty) return false;  // kernel smaller than the stride: not worth a special case (no such layer in the reference)
      if (!tc_conv_supported(sp.g2, small, phase_view(big, g.st, a, b), true)) return false;
    }
  return true;
}

int tc_strided_scatter(const Geom &g, const T4 &small, const float *w, const T4 &big, const Epi &epi, void *ws, size_t ws_bytes,
                       cudaStream_t st) {
  for (int a = 0; a < g.st; ++a)
    for (int b = 0; b < g.st; ++b) {
      ScatterPhase sp;
      SRB_REQUIRE(plan_scatter_phase(g, a, b, &sp), SRB_EUNSUPPORTED, "transposed conv: no phase plan");
      if (sp.g2.Ho == 0) continue;
      SRB_REQUIRE(!sp.empty, SRB_EUNSUPPORTED, "transposed conv: kernel smaller than the stride");
      Epi e = epi;
      e.residual = phase_view(epi.residual, g.st, a, b);
      e.preact = phase_view(epi.preact, g.st, a, b);
      e.mask = phase_view(epi.mask, g.st, a, b);
      SRB_REQUIRE(!epi.bits_in && !epi.bits_out, SRB_EUNSUPPORTED, "packed ReLU bits with a strided / transposed convolution");
      // sp.g2 is the launch's gather geometry: input = small (Ci := g.Co), output = this phase of big (Co := g.Ci)
      int rc = tc_conv_gather(sp.g2, small, w, true, phase_view(big, g.st, a, b), e, ws, ws_bytes, st, sp.opt);
      if (rc) return rc;
    }
  return SRB_OK;
}

// (3) tc_strided_wgrad -- weight gradient of a strided Conv2d: with r - pad = st*d + a the tap reads phase image
//     P_ab[i, j] = big[st*i + a, st*j + b] at (oy + d_r, ox + d_s): per phase a STRIDE-1 weight gradient between dz and a strided
//     VIEW of x (T4 strides x st: the TMA map walks it directly), with the sub-filter d in [dmin, dmax] expressed as a k' x k' filter
//     of padding pad' = -dmin (both taken as the maximum over the two axes and all phases; the surplus taps are computed and dropped
//     by the finish kernel, which scatters the others to their place in the kh x kw filter).  st*st launches on k_tc_wgrad.
namespace {
struct WgradPhases {
  Geom g2;      // the stride-1 geometry shared by all phases except Hi / Wi (set per phase)
  int kp, padp;
};
bool plan_wgrad_phases(const Geom &g, WgradPhases *wp) {
  const int st = g.st;
  if (st < 2 || st > 4 || g.ps != 1) return false;
  int dmin = 0, dmax = 0;
  bool any = false;
  for (int r = 0; r < (g.kh > g.kw ? g.kh : g.kw); ++r) {
    const int q = r - g.pad;
    const int d = q >= 0 ? q / st : -((-q + st - 1) / st);  // floor division
    if (!any || d < dmin) dmin = d;
    if (!any || d > dmax) dmax = d;
    any = true;
  }
  if (dmin > 0) dmin = 0;  // keep pad' >= 0
  wp->padp = -dmin;
  wp->kp = dmax - dmin + 1;
  if (wp->kp > 16) return false;
  wp->g2 = Geom{g.N, g.Ci, 0, 0, g.Co, g.Ho, g.Wo, wp->kp, wp->kp, 1, wp->padp, 1};
  return true;
}
}  // namespace

bool tc_strided_wgrad_supported(const Geom &g, const T4 &small, const T4 &big) {
  WgradPhases wp;
  if (!plan_wgrad_phases(g, &wp)) return false;
  if (big.dt != SRB_F32 || small.dt != SRB_F32 || g.Ci <= 4) return false;
  for (int a = 0; a < g.st; ++a)
    for (int b = 0; b < g.st; ++b) {
      Geom g2 = wp.g2;
      g2.Hi = a < g.Hi ? (g.Hi - a + g.st - 1) / g.st : 0;
      g2.Wi = b < g.Wi ? (g.Wi - b + g.st - 1) / g.st : 0;
      if (g2.Hi == 0 || g2.Wi == 0) return false;
      if (!tc_wgrad_supported(g2, small, phase_view(big, g.st, a, b))) return false;
    }
  return true;
}

size_t tc_strided_wgrad_ws_bytes(const Geom &g) {
  WgradPhases wp;
  if (!plan_wgrad_phases(g, &wp)) return 0;
  Geom g2 = wp.g2;
  g2.Hi = (g.Hi + g.st - 1) / g.st;
  g2.Wi = (g.Wi + g.st - 1) / g.st;
  return tc_wgrad_ws_bytes(g2);
}

int tc_strided_wgrad(const Geom &g, const T4 &small, const T4 &big, float *dw, float *db_small, float scale, int accumulate,
                     void *ws, size_t ws_bytes, cudaStream_t st) {
  WgradPhases wp;
  SRB_REQUIRE(plan_wgrad_phases(g, &wp), SRB_EUNSUPPORTED, "strided wgrad: no phase plan");
  bool first = true;
  for (int a = 0; a < g.st; ++a)
    for (int b = 0; b < g.st; ++b) {
      Geom g2 = wp.g2;
      g2.Hi = (g.Hi - a + g.st - 1) / g.st;
      g2.Wi = (g.Wi - b + g.st - 1) / g.st;
      const WgPhase ph{g.st, a, b, g.pad, g.kh, g.kw};
      // every phase sees the whole dz: the bias gradient comes from the first one only
      int rc = tc_conv_wgrad(g2, small, phase_view(big, g.st, a, b), dw, first ? db_small : nullptr, scale, accumulate, ws, ws_bytes, st, 1, &ph);
      if (rc) return rc;
      first = false;
    }
  return SRB_OK;
}

size_t tc_strided_ws_bytes(const Geom &g) {
  size_t best = 0;
  GatherPhase gp;
  if (plan_gather(g, &gp)) best = tc_conv_ws_bytes(gp.g2);
  if (g.st >= 2 && g.st <= 4)
    for (int a = 0; a < g.st; ++a)
      for (int b = 0; b < g.st; ++b) {
        ScatterPhase sp;
        if (!plan_scatter_phase(g, a, b, &sp) || sp.g2.Ho == 0 || sp.empty) continue;
        const size_t v = tc_conv_ws_bytes(sp.g2);
        if (v > best) best = v;
      }
  return best;
}

}  // namespace srb
