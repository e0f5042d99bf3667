"""GPU suite (-m gpu): the CUDA path, called through the C-ABI (ctypes) behind the block API, against
(1) the golden vectors produced from the unmodified reference, (2) the CPU oracle on seeded inputs,
(3) size-independent properties at BASELINE.json's full sizes.

Tolerances (north_star: 1e-3 relative fp32; PixelShuffle permutation bit-exact):
  math='fp32' (CUDA-core kernels)      rel-L2 <= 2e-5 (summation order only)
  math='auto' (tcgen05 TF32 kernels)   rel-L2 <= 1e-3 per op and on the shallow golden nets' outputs

Backward through ReLU/PReLU/LeakyReLU is discontinuous in the pre-activation: perturbing z by a relative eps
flips the mask of the ~eps fraction of elements with |z| < eps*sigma, and each flip changes a gradient entry
by O(1), so rel-L2 of dX/dW is ~sqrt(eps) (1-3e-2 for TF32; the cuDNN TF32 build the reference runs on
has the same property).  The op-level TF32 gate therefore evaluates the oracle's backward on the SAME
activation pattern the GPU forward produced (the derivative kernels themselves must still be 1e-3 exact);
the net-level TF32 gate checks outputs at 1e-3 and gradients by cosine similarity; the fp32 path is gated
at 1e-4 end to end with no conditioning.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu

import srb200
from srb200 import models as M
from oracle import torch_ref as R
from util import load_golden, load_state, rel_l2

DEV = "cuda:0"
TOL = {"fp32": 2e-5, "auto": 1e-3}


@pytest.fixture(autouse=True)
def _reset_math():
    srb200.set_math("auto")
    srb200.set_grad_scale(1.0)
    yield
    srb200.set_math("auto")


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")


BLOCKS = {
    "convblock_k3_relu": lambda: srb200.ConvBlock(8, 16, 3, 1, 1, activation="relu", norm=None),
    "convblock_k5_p0_prelu": lambda: srb200.ConvBlock(3, 12, 5, 1, 0, activation="prelu", norm=None),
    "convblock_k3_s2_lrelu": lambda: srb200.ConvBlock(6, 10, 3, 2, 1, activation="lrelu", norm=None),
    "convblock_default_k4s2_nobias": lambda: srb200.ConvBlock(4, 8, bias=False, activation=None, norm=None),
    "psblock_r4": lambda: srb200.PSBlock(8, 3, 4, 3, 1, 0, activation=None, norm=None),
    "psblock_r2_prelu": lambda: srb200.PSBlock(8, 8, 2, activation="prelu", norm=None),
    "resnetblock_relu": lambda: srb200.ResnetBlock(8, norm=None),
    "resnetblock_prelu": lambda: srb200.ResnetBlock(8, activation="prelu", norm=None),
    "deconvblock_k4s2": lambda: srb200.DeconvBlock(6, 4, activation="relu", norm=None),
    "upsample2x_ps": lambda: srb200.Upsample2xBlock(8, 8, upsample="ps", activation=None, norm=None),
    "upsample2x_deconv": lambda: srb200.Upsample2xBlock(4, 4, upsample="deconv", activation="lrelu", norm=None),
    "fsrcnn_tail": lambda: torch.nn.Sequential(srb200.PReLU(), srb200.ConvTranspose2d(6, 3, 9, 4, 3, output_padding=1)),
}


@pytest.mark.parametrize("name", sorted(BLOCKS))
@pytest.mark.parametrize("math", ["fp32", "auto"])
def test_block_matches_reference_golden(name, math):
    _need_gpu()
    srb200.set_math(math)
    g = load_golden("block_" + name)
    blk = BLOCKS[name]()
    load_state(blk, g)
    blk.to(DEV)
    x = torch.from_numpy(g["x"]).to(DEV).requires_grad_(True)
    y = blk(x)
    y.backward(torch.from_numpy(g["gy"]).to(DEV))
    tol = TOL[math]
    assert tuple(y.shape) == g["y"].shape
    assert rel_l2(y.detach(), g["y"]) < tol
    assert rel_l2(x.grad, g["gx"]) < tol
    for k, p in blk.named_parameters():
        assert rel_l2(p.grad, g["grad:" + k]) < tol, k


NETS = {"srcnn": "l2", "espcn": "l2", "fsrcnn": "l2", "vdsr": "l2", "edsr": "l1", "srgan_g": "l2"}


@pytest.mark.parametrize("name", sorted(NETS))
@pytest.mark.parametrize("math", ["fp32", "auto", "exact"])
def test_net_matches_reference_golden(name, math):
    """Reduced-width nets whose fixtures hold the reference's parameters, output, loss and gradients in full.
    Outputs/loss against the fixture; gradients against the oracle replayed on the GPU's activation pattern at the same
    tolerance class (netcheck.run_against_oracle) -- fp32 also straight against the fixture's gradients."""
    _need_gpu()
    from netcheck import run_against_oracle, tolerances
    g = load_golden("net_" + name)
    args = tuple(int(v) for v in g["args"])
    ref = R.build(name, args, init=False)
    load_state(ref, g)
    kind = "l1" if NETS[name] == "l1" else "mse"
    r = run_against_oracle(name, args, kind, torch.from_numpy(g["x"]), torch.from_numpy(g["target"]), math, ref=ref)
    ty, tg = tolerances(math, r["n_convs"])
    if math == "fp32":
        ty = tg = 1e-4
    assert rel_l2(r["y"], g["y"]) < ty
    loss = (TF.l1_loss if kind == "l1" else TF.mse_loss)(r["y"], torch.from_numpy(g["target"]))
    assert abs(loss.item() - float(g["loss"])) < ty * abs(float(g["loss"]))
    for k, e in r["grad_errs"].items():
        assert e < tg, (k, e)


# ---- op-level sweep against the CPU oracle ---------------------------------------------------------
#        Cin Cout k  s  p  act      res    ps  H   W   N
SWEEP = [
    (3, 64, 9, 1, 0, "relu", False, 1, 20, 22, 2),     # SRCNN first layer
    (64, 32, 5, 1, 0, "relu", False, 1, 14, 15, 2),    # SRCNN second
    (32, 3, 5, 1, 0, None, False, 1, 12, 12, 2),       # SRCNN last
    (3, 64, 5, 1, 0, "relu", False, 1, 18, 17, 2),     # ESPCN first
    (64, 32, 3, 1, 0, "relu", False, 1, 16, 16, 3),    # ESPCN second
    (32, 3, 3, 1, 0, None, False, 4, 13, 14, 2),       # ESPCN PSBlock (32 -> 48 -> PS4)
    (64, 64, 3, 1, 1, "relu", False, 1, 16, 16, 2),    # VDSR / EDSR body
    (64, 64, 3, 1, 1, None, True, 1, 12, 20, 2),       # ResnetBlock conv2 + skip
    (64, 3, 3, 1, 1, None, True, 1, 16, 16, 2),        # VDSR output conv + global residual
    (64, 64, 3, 1, 1, "prelu", False, 2, 8, 8, 2),     # SRGAN/EDSR upsampler (64 -> 256 -> PS2)
    (64, 3, 9, 1, 4, None, False, 1, 16, 16, 1),       # SRGAN output conv
    (64, 64, 3, 2, 1, "lrelu", False, 1, 16, 16, 2),   # SRGAN D stride-2
    (128, 128, 3, 2, 1, "lrelu", False, 1, 8, 8, 2),
    (64, 64, 3, 2, 1, "relu", False, 1, 17, 15, 2),    # stride 2 on odd sizes (phase images of unequal size)
    (32, 64, 4, 2, 1, None, False, 1, 12, 12, 1),      # ConvBlock defaults k4 s2 p1 (base_networks.py:40)
    (64, 32, 5, 3, 2, "lrelu", False, 1, 14, 13, 1),   # stride 3
    (56, 12, 1, 1, 0, "prelu", False, 1, 10, 10, 2),   # FSRCNN shrink
    (12, 12, 3, 1, 1, None, False, 1, 10, 10, 2),      # FSRCNN map
    (12, 56, 1, 1, 0, "prelu", False, 1, 10, 10, 2),   # FSRCNN expand
    (32, 32, 3, 1, 1, "relu", False, 1, 7, 5, 1),      # ragged: smaller than one tile
    (96, 160, 3, 1, 1, "relu", False, 1, 9, 9, 1),     # non power-of-two channel counts
    (256, 256, 3, 1, 1, "relu", False, 1, 8, 8, 1),    # EDSR-256 body
    (64, 64, 1, 1, 0, None, False, 1, 5, 6, 2),
    (40, 2, 3, 1, 1, "relu", False, 1, 45, 9, 2),      # skinny-output wgrad: ragged channel tile, several row bands
    (64, 4, 3, 1, 0, None, False, 1, 11, 13, 3),       # skinny-output wgrad, no padding
    (16, 1, 1, 1, 0, None, False, 1, 6, 7, 2),         # 1x1, single output channel
    (64, 64, 3, 1, 1, "relu", False, 1, 6, 150, 1),    # wide rows (column-tiled wgrad bands)
    (32, 64, 3, 1, 1, None, False, 1, 20, 140, 1),
]


def _ref_op(x, w, b, alpha, res, k, s, p, act, ps):
    z = TF.conv2d(x, w, b, s, p)
    if act == "relu":
        z = TF.relu(z)
    elif act == "prelu":
        z = TF.prelu(z, alpha)
    elif act == "lrelu":
        z = TF.leaky_relu(z, 0.2)
    if ps > 1:
        z = TF.pixel_shuffle(z, ps)
    if res is not None:
        z = z + res
    return z


@pytest.mark.parametrize("case", SWEEP)
@pytest.mark.parametrize("math", ["fp32", "auto"])
@pytest.mark.parametrize("cl", [False, True])
def test_fused_conv_vs_oracle(case, math, cl):
    _run_conv_case(case, math, cl)


# The planner only picks the row-stacked kernels (k_conv_rs; wgrad with filter rows stacked along N) for layers that give every
# CTA >= 8 rows, i.e. not for test-sized tensors: debug flag 1024 forces k_conv_rs wherever it is usable, 512 the stacked wgrad.
#           Cin Cout k  s  p  act      res    ps  H   W   N
SWEEP_RS = [c for c in SWEEP if c[3] == 1 and c[0] > 4] + [
    (64, 32, 3, 1, 0, "relu", False, 1, 30, 60, 5),    # two images per M tile, odd batch (last tile half empty)
    (48, 32, 3, 1, 2, None, False, 1, 28, 56, 3),      # dgrad-like full padding, partial last chunk (48 = 32 + 16 channels)
    (40, 16, 3, 1, 1, "lrelu", False, 1, 17, 33, 4),   # three images per M tile, 8 channels in the last chunk
    (40, 16, 3, 1, 1, None, True, 1, 17, 33, 4),       # the same with a residual (the reference never fuses act + residual)
    (32, 3, 3, 1, 0, None, False, 4, 29, 58, 2),       # PixelShuffle(4) -> NCHW
    (64, 64, 3, 1, 1, "prelu", False, 2, 16, 16, 8),   # PixelShuffle(2) -> NHWC, N = 4 * 64 * 3 > 256: not row-stacked, must still pass
    (64, 64, 3, 1, 1, "relu", False, 1, 12, 300, 1),   # three column strips per row
    (64, 16, 5, 1, 2, "relu", False, 1, 20, 40, 2),    # 5 x 5: N = 80
    (32, 64, 3, 1, 1, "relu", False, 1, 70, 24, 6),    # long strips: the TMEM ring wraps many times (R = 8 blocks)
    (64, 48, 3, 1, 1, None, False, 1, 40, 30, 4),      # NT = 48: ring of 10 blocks
    (64, 32, 3, 1, 1, "relu", False, 1, 5, 20, 1),     # one row per CTA: the second row stream of every CTA is empty
    (64, 32, 3, 1, 1, "relu", False, 1, 61, 20, 5),    # 305 rows over 148 CTAs: 2-3 rows per CTA, streams of unequal length
    (64, 64, 3, 1, 1, "relu", False, 1, 37, 150, 2),   # wide image: three column strips, Cout 64 (wgrad N = 192 when forced)
    (32, 80, 3, 1, 0, None, False, 1, 21, 45, 3),      # Cout 80 (padded to 96 = 3 co blocks: no NT = 64 tiling, the wgrad stays rows-stacked)
    (32, 128, 2, 1, 0, None, False, 1, 18, 26, 3),     # 2 x 2 filter: N = 2 x 128 = 256
]


@pytest.mark.parametrize("one_stream", [False, True], ids=["streams-auto", "one-stream"])
@pytest.mark.parametrize("case", SWEEP_RS)
def test_row_stacked_kernels_vs_oracle(case, one_stream):
    """k_conv_rs (forced regardless of problem size) + row-stacked wgrad; NT <= 32 layers run two row streams per CTA unless
    flag 65536 forces one.  The one-stream variant also forces the rows+co-stacked wgrad flavour (N = kh * NT, flag 524288)
    wherever it has a plan (Cout >= 64 with kh * 64 <= 256)."""
    from util import set_debug_flags as setf
    setf(1024 | 512 | ((65536 | 524288) if one_stream else 0))
    try:
        _run_conv_case(case, "auto", True)
    finally:
        setf(0)


def _run_conv_case(case, math, cl):
    _need_gpu()
    srb200.set_math(math)
    Cin, Cout, k, s, p, act, res, ps, H, W, N = case
    gen = torch.Generator().manual_seed(11)
    x = torch.randn(N, Cin, H, W, generator=gen)
    w = torch.randn(Cout * ps * ps, Cin, k, k, generator=gen) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout * ps * ps, generator=gen) * 0.1
    alpha = torch.tensor([0.25])
    xr = x.clone().requires_grad_(True)
    wr, br, ar = w.clone().requires_grad_(True), b.clone().requires_grad_(True), alpha.clone().requires_grad_(True)
    y0 = TF.conv2d(x, w, b, s, p)
    oshape = (N, Cout, y0.shape[2] * ps, y0.shape[3] * ps)
    r = torch.randn(oshape, generator=gen) if res else None
    rr = r.clone().requires_grad_(True) if res else None
    gy = torch.randn(oshape, generator=gen)

    xg = x.to(DEV)
    if cl:
        xg = xg.contiguous(memory_format=torch.channels_last)
    xg.requires_grad_(True)
    wg, bg, ag = (t.to(DEV).requires_grad_(True) for t in (w, b, alpha))
    rg = r.to(DEV).requires_grad_(True) if res else None
    y = srb200.conv2d(xg, wg, bg, s, p, activation=act, alpha=ag if act == "prelu" else None, residual=rg,
                      pixel_shuffle=ps)
    y.backward(gy.to(DEV))
    tol = TOL[math]
    yref = _ref_op(xr, wr, br, ar, rr, k, s, p, act, ps)
    assert rel_l2(y.detach(), yref.detach()) < tol
    if math == "auto" and act is not None:
        # oracle backward on the activation pattern of the GPU forward (module docstring)
        zr = TF.conv2d(xr, wr, br, s, p)
        if ps > 1:
            zr = TF.pixel_shuffle(zr, ps)
        yg = y.detach().cpu() - (r if res else 0)
        m = (yg > 0).float() if act != "prelu" else ((yg > 0) if alpha.item() > 0 else (yg < 0)).float()
        slope_t = {"relu": 0.0, "lrelu": 0.2}.get(act, ar)
        yref = zr * m + zr * (1 - m) * slope_t
        if res:
            yref = yref + rr
    yref.backward(gy)
    assert rel_l2(xg.grad, xr.grad) < tol
    assert rel_l2(wg.grad, wr.grad) < tol
    assert rel_l2(bg.grad, br.grad) < tol
    if act == "prelu":
        assert abs(ag.grad.item() - ar.grad.item()) < max(tol * 10 * abs(ar.grad.item()), 1e-4)
    if res:
        assert torch.equal(rg.grad.cpu(), gy)


@pytest.mark.parametrize("math", ["fp32", "auto"])
@pytest.mark.parametrize("cl", [False, True])
@pytest.mark.parametrize("k,s,p,op,Ci,Co", [(9, 4, 3, 1, 56, 3), (4, 2, 1, 0, 16, 8), (3, 1, 1, 0, 8, 8), (5, 3, 2, 2, 6, 5),
                                            (4, 2, 1, 0, 64, 64), (3, 2, 1, 1, 32, 16), (2, 2, 0, 0, 16, 16), (9, 4, 3, 1, 64, 32)])
def test_conv_transpose_vs_oracle(k, s, p, op, Ci, Co, math, cl):
    """ConvTranspose2d forward / backward.  channels_last inputs with Cin % 4 == 0 run as st x st output-phase convolutions on
    the tcgen05 kernel in 'auto' mode (csrc/strided.cu); everything else on the CUDA-core scatter kernel."""
    _need_gpu()
    srb200.set_math(math)
    # fp32 mode: exact-order-independent fp32; auto: NHWC outputs with C % 4 == 0 are stored tf32-rounded (2^-11 relative)
    tol = 2e-5 if math == "fp32" else 1e-3
    gen = torch.Generator().manual_seed(12)
    x = torch.randn(2, Ci, 7, 6, generator=gen)
    w = torch.randn(Ci, Co, k, k, generator=gen) * 0.1
    b = torch.randn(Co, generator=gen)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    yref = TF.conv_transpose2d(xr, wr, br, s, p, op)
    gy = torch.randn(yref.shape, generator=gen)
    yref.backward(gy)
    xg = x.to(DEV)
    if cl:
        xg = xg.contiguous(memory_format=torch.channels_last)
    xg.requires_grad_(True)
    wg, bg = (t.to(DEV).requires_grad_(True) for t in (w, b))
    y = srb200.conv_transpose2d(xg, wg, bg, s, p, op)
    y.backward(gy.to(DEV))
    assert rel_l2(y.detach(), yref.detach()) < tol
    assert rel_l2(xg.grad, xr.grad) < tol
    assert rel_l2(wg.grad, wr.grad) < tol
    assert rel_l2(bg.grad, br.grad) < tol


@pytest.mark.parametrize("r,C,H,W", [(2, 64, 8, 8), (4, 3, 9, 7), (3, 5, 4, 6), (2, 256, 4, 4)])
@pytest.mark.parametrize("math", ["fp32", "auto"])
def test_pixel_shuffle_permutation_bit_exact(r, C, H, W, math):
    """1x1 conv with an identity filter: the fused store must reproduce nn.PixelShuffle bit for bit
    (values are tf32-representable small integers, so the tensor path is exact too)."""
    _need_gpu()
    srb200.set_math(math)
    cin = C * r * r
    x = torch.randint(-64, 64, (2, cin, H, W)).float()
    w = torch.eye(cin).reshape(cin, cin, 1, 1)
    for cl in (False, True):
        xg = x.to(DEV)
        if cl:
            xg = xg.contiguous(memory_format=torch.channels_last)
        y = srb200.conv2d(xg, w.to(DEV), None, 1, 0, pixel_shuffle=r)
        assert torch.equal(y.cpu(), TF.pixel_shuffle(x, r))
        # and the un-shuffle in backward
        xg = xg.detach().requires_grad_(True)
        y = srb200.conv2d(xg, w.to(DEV), None, 1, 0, pixel_shuffle=r)
        gy = torch.randint(-8, 8, tuple(y.shape)).float()
        y.backward(gy.to(DEV))
        assert torch.equal(xg.grad.cpu(), TF.pixel_unshuffle(gy, r))


@pytest.mark.parametrize("math", ["fp32", "auto"])
@pytest.mark.parametrize("topology", ["chain", "two_convs", "conv_and_torch"])
def test_relu_backward_folded_into_consumer_dgrad(math, topology):
    """The consumer's dgrad epilogue applies the producer's ReLU mask (srb_conv_dgrad relu_mask).  Folding must not change
    a single bit of any gradient, and must stay correct when the ReLU output has more than one consumer."""
    _need_gpu()
    srb200.set_math(math)
    gen = torch.Generator().manual_seed(5)
    x0 = torch.randn(2, 32, 12, 14, generator=gen)
    ws = [torch.randn(32, 32, 3, 3, generator=gen) / 17 for _ in range(3)]
    bs = [torch.randn(32, generator=gen) * 0.1 for _ in range(3)]

    def run(fuse, use_srb=True):
        srb200.set_fuse_relu_backward(fuse)
        dev = DEV if use_srb else "cpu"
        x = x0.to(dev)
        if use_srb:
            x = x.contiguous(memory_format=torch.channels_last)
        x.requires_grad_(True)
        w = [t.to(dev).requires_grad_(True) for t in ws]
        b = [t.to(dev).requires_grad_(True) for t in bs]
        if use_srb:
            conv = lambda t, i, act: srb200.conv2d(t, w[i], b[i], 1, 1, activation=act)
        else:
            conv = lambda t, i, act: (TF.relu(TF.conv2d(t, w[i], b[i], 1, 1)) if act else TF.conv2d(t, w[i], b[i], 1, 1))
        h = conv(x, 0, "relu")
        if topology == "chain":
            out = conv(conv(h, 1, "relu"), 2, None)
        elif topology == "two_convs":
            out = conv(h, 1, None) + conv(h, 2, None)
        else:
            out = conv(h, 1, None) + 0.5 * h
        out.square().sum().backward()
        return [x.grad] + [t.grad for t in w + b]

    try:
        fused, plain = run(True), run(False)
    finally:
        srb200.set_fuse_relu_backward(True)
    for a, c in zip(fused, plain):
        assert (a is None and c is None) or torch.equal(a, c)  # unused layers (conv_and_torch) have no gradient
    if math == "fp32":
        for a, c in zip(fused, run(True, use_srb=False)):
            assert (a is None and c is None) or rel_l2(a, c) < 1e-4


@pytest.mark.parametrize("Cin,Cout,k,p,H,W", [(32, 64, 3, 1, 9, 13), (3, 64, 5, 0, 18, 17), (64, 32, 3, 0, 16, 16)])
def test_packed_relu_sign_bits_match_output(Cin, Cout, k, p, H, W):
    """srb_conv_fprop's relu_bits: bit c%16 of word [n, oy, ox, c/16] is set exactly where y > 0 (bit-exact)."""
    _need_gpu()
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(2, Cin, H, W, generator=gen).to(DEV)
    if Cin >= 8:
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    w = (torch.randn(Cout, Cin, k, k, generator=gen) / (Cin * k * k) ** 0.5).to(DEV).requires_grad_(True)
    b = (torch.randn(Cout, generator=gen) * 0.1).to(DEV).requires_grad_(True)
    y = srb200.conv2d(x, w, b, 1, p, activation="relu")
    tok = getattr(y, "_srb_relu", None)
    assert tok is not None and tok.bits is not None, "tensor-path ReLU layers must emit the packed mask"
    bits = tok.bits.cpu().numpy().view(np.uint16)                      # (N, Ho, Wo, Cout/16)
    pos = (y.detach().permute(0, 2, 3, 1) > 0).cpu().numpy()            # (N, Ho, Wo, Cout)
    unpacked = ((bits[..., None] >> np.arange(16, dtype=np.uint16)) & 1).reshape(pos.shape).astype(bool)
    assert np.array_equal(unpacked, pos)


@pytest.mark.parametrize("kind", ["mse", "l1"])
@pytest.mark.parametrize("shape", [(2, 3, 17, 19), (4, 3, 64, 64), (1, 1, 1, 3)])
def test_fused_losses_match_torch(kind, shape):
    """srb_loss_fwd / srb_loss_bwd vs nn.MSELoss / nn.L1Loss (mean), including the n % 4 tail and an upstream factor."""
    _need_gpu()
    gen = torch.Generator().manual_seed(9)
    y0, t0 = torch.randn(shape, generator=gen), torch.randn(shape, generator=gen)
    t0.view(-1)[0] = y0.view(-1)[0]  # an exact tie: l1 gradient must be 0 there
    yr = y0.clone().requires_grad_(True)
    ref = (TF.l1_loss if kind == "l1" else TF.mse_loss)(yr, t0)
    (3.0 * ref).backward()
    yg = y0.to(DEV).requires_grad_(True)
    out = (srb200.l1_loss if kind == "l1" else srb200.mse_loss)(yg, t0.to(DEV))
    (3.0 * out).backward()
    assert abs(out.item() - ref.item()) <= 2e-6 * abs(ref.item())
    assert rel_l2(yg.grad, yr.grad) < 1e-6
    assert yg.grad.view(-1)[0].item() == 0.0 if kind == "l1" else True


@pytest.mark.parametrize("name,args,shape", [("espcn", (3, 64, 4), (4, 3, 24, 24)), ("vdsr", (3, 64, 3), (2, 3, 32, 32))])
def test_cuda_graph_replay_equals_eager_training(name, args, shape):
    """srb200.TrainStepGraphs: three replayed steps must leave exactly the parameters that three eager steps leave
    (the library neither allocates nor synchronises, and every tensor map is encoded at capture time)."""
    _need_gpu()
    from srb200 import host

    def make():
        torch.manual_seed(0)
        net = M.MODELS[name](*args)
        host.init_model(name, net)
        net.to(DEV).train()
        opt = host.make_optimizer(name, net.parameters(), lr=1e-3, capturable=True)
        return net, opt, srb200.GradBucket(net, world_size=1)

    gen = torch.Generator().manual_seed(4)
    xs = [torch.rand(shape, generator=gen).to(DEV) for _ in range(2)]
    net_e, opt_e, bk_e = make()
    with torch.no_grad():
        oshape = net_e(xs[0]).shape
    ts = [torch.rand(oshape, generator=gen).to(DEV) for _ in range(2)]
    lossf = host.loss_for(name, fused=True)
    clip = host.VDSR_CLIP if name == "vdsr" else None

    def eager(net, opt, bk, i):
        bk.begin_step()
        l = lossf(net(xs[i % 2]), ts[i % 2])
        l.backward()
        if clip is not None:
            torch.nn.utils.clip_grad_norm_(net.parameters(), clip)
        opt.step()
        return l.item()

    losses_e = [eager(net_e, opt_e, bk_e, i) for i in range(4)]
    bk_e.detach()
    net_g, opt_g, bk_g = make()
    first = eager(net_g, opt_g, bk_g, 0)  # one eager step creates the optimizer state (not capturable: host-side init)
    st = srb200.TrainStepGraphs(net_g, lossf, opt_g, bk_g, slots=list(zip(xs, ts)), clip_norm=clip)
    assert st.launches_per_step >= 8
    losses_g = [first] + [st.step(i % 2).item() for i in range(1, 4)]
    assert losses_g == losses_e
    for (k, p), (_, q) in zip(net_g.named_parameters(), net_e.named_parameters()):
        assert torch.equal(p, q), k
    bk_g.detach()


def test_act_corner_cases_at_zero():
    """z == 0: ReLU grad is 0, PReLU/LeakyReLU take the slope branch (ATen semantics)."""
    _need_gpu()
    srb200.set_math("fp32")
    x = torch.zeros(1, 1, 4, 4)
    x[0, 0, 0, 0], x[0, 0, 1, 1] = 1.0, -2.0
    w = torch.ones(1, 1, 1, 1)
    gy = torch.ones(1, 1, 4, 4)
    for act, mod in (("relu", torch.nn.ReLU()), ("prelu", torch.nn.PReLU()), ("lrelu", torch.nn.LeakyReLU(0.2))):
        xr = x.clone().requires_grad_(True)
        mod(TF.conv2d(xr, w)).backward(gy)
        xg = x.to(DEV).requires_grad_(True)
        alpha = torch.tensor([0.25], device=DEV, requires_grad=True)
        y = srb200.conv2d(xg, w.to(DEV), None, 1, 0, activation=act, alpha=alpha if act == "prelu" else None)
        y.backward(gy.to(DEV))
        assert torch.equal(xg.grad.cpu(), xr.grad), act
        if act == "prelu":
            assert abs(alpha.grad.item() - mod.weight.grad.item()) < 1e-6


def test_empty_batch_and_tiny_images():
    _need_gpu()
    w = torch.randn(8, 4, 3, 3, device=DEV)
    w.requires_grad_(True)
    y = srb200.conv2d(torch.zeros(0, 4, 8, 8, device=DEV), w, None, 1, 1)
    assert tuple(y.shape) == (0, 8, 8, 8)
    y.sum().backward()
    assert torch.count_nonzero(w.grad).item() == 0
    w = w.detach()
    x = torch.randn(1, 4, 3, 3, device=DEV)
    y = srb200.conv2d(x, w, None, 1, 0)
    assert tuple(y.shape) == (1, 8, 1, 1)
    assert rel_l2(y, TF.conv2d(x.cpu(), w.cpu())) < 1e-3  # Cin = 4 runs on the tf32 tensor path in 'auto' mode
    with pytest.raises(srb200.SrbError):
        srb200.conv2d(torch.zeros(1, 4, 2, 2, device=DEV), w, None, 1, 0)  # kernel larger than input


# ---- size-independent properties at the BASELINE.json sizes ------------------------------------------
def test_espcn_cfg2_full_size_properties():
    """ESPCN x4, batch 128, 64x64 LR (BASELINE cfg2): adjointness <conv(x),g> == <x,dgrad(g)> and
    <dW,W> == <conv_W(x),g> (conv is linear in x and in W), plus the shuffle checksum."""
    _need_gpu()
    torch.manual_seed(0)
    net = M.ESPCN(3, 64, 4).to(DEV)
    R.init_normal(net)
    x = torch.rand(128, 3, 64, 64, device=DEV)
    y = net(x)
    assert tuple(y.shape) == (128, 3, 224, 224)
    # last block alone (linear: no activation): adjoint identities
    blk = net.layers[2]
    h = net.layers[1](net.layers[0](x)).detach().requires_grad_(True)
    out = blk(h)
    g = torch.randn_like(out)
    out.backward(g)
    lhs = (out.detach().double() * g.double()).sum().item()
    bias_term = (TF.pixel_shuffle(blk.conv.bias.detach().view(1, 48, 1, 1).expand(128, 48, 56, 56), 4).double()
                 * g.double()).sum().item()
    dx_term = (h.detach().double() * h.grad.double()).sum().item()
    dw_term = (blk.conv.weight.detach().double() * blk.conv.weight.grad.double()).sum().item()
    scale = abs(lhs) + abs(bias_term) + 1.0
    assert abs((lhs - bias_term) - dx_term) / scale < 2e-3
    assert abs((lhs - bias_term) - dw_term) / scale < 2e-3
    db_ref = TF.pixel_unshuffle(g, 4).sum(dim=(0, 2, 3))
    assert rel_l2(blk.conv.bias.grad, db_ref) < 1e-3  # sum of tf32-rounded dz over 401k pixels


def test_vdsr_body_layer_full_size_linearity():
    """64->64 k3 p1 at 128x128 (BASELINE cfg3 body layer, batch reduced to 8): conv(a*x1+b*x2) == a*conv(x1)+b*conv(x2)."""
    _need_gpu()
    torch.manual_seed(1)
    w = torch.randn(64, 64, 3, 3, device=DEV) / 24.0
    x1 = torch.randn(8, 64, 128, 128, device=DEV).contiguous(memory_format=torch.channels_last)
    x2 = torch.randn_like(x1)
    y1 = srb200.conv2d(x1, w, None, 1, 1)
    y2 = srb200.conv2d(x2, w, None, 1, 1)
    y12 = srb200.conv2d(2.0 * x1 - 0.5 * x2, w, None, 1, 1)
    assert rel_l2(y12, 2.0 * y1 - 0.5 * y2) < 1e-3
    # spot-check one image against the oracle
    ref = TF.conv2d(x1[:1].cpu().contiguous(), w.cpu(), None, 1, 1)
    assert rel_l2(y1[:1], ref) < 1e-3


def test_training_steps_track_the_oracle():
    """Restated loop body (espcn.py:126-131, vdsr.py:142-150 incl. clipping): 3 optimizer steps, parameters stay
    within tolerance of the CPU oracle's trajectory."""
    _need_gpu()
    for name, args, xs in (("espcn", (3, 64, 4), (4, 3, 20, 20)), ("vdsr", (3, 64, 3), (2, 3, 16, 16))):
        ref = R.build(name, args, seed=0)
        ours = M.MODELS[name](*args)
        ours.load_state_dict(ref.state_dict())
        ours.to(DEV)
        lr = 1e-5 if name == "espcn" else 1e-2  # Adam's g/sqrt(v) turns rounding of ~0 gradients into +-lr steps
        oref = R.make_optimizer(name, ref.parameters(), lr=lr)
        oours = R.make_optimizer(name, ours.parameters(), lr=lr)
        gen = torch.Generator().manual_seed(3)
        for _ in range(3):
            x = torch.rand(xs, generator=gen)
            t = torch.rand(R.output_shape(name, args, xs), generator=gen)
            l0 = R.train_step(name, ref, oref, x, t)
            l1 = R.train_step(name, ours, oours, x.to(DEV), t.to(DEV))
            assert abs(l0.item() - l1.item()) < 1e-3 * abs(l0.item())
        # Adam divides by sqrt(v): a bias whose gradient is ~0 moves by +-lr on rounding noise, so the
        # trajectory bound is looser than the per-op 1e-3 (biases start at exactly 0: error relative to lr)
        for (k, a), (_, b) in zip(ref.state_dict().items(), ours.state_dict().items()):
            assert rel_l2(b, a) < (5e-3 if name == "espcn" else 1e-3), (name, k)
