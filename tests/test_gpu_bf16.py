"""GPU suite (-m gpu): the bf16 STORAGE path (math='bf16', BASELINE cfg4 "EDSR 32 resblocks x 256 ch, bf16").

Contract: activations / activation gradients are bf16 channels_last tensors, operands feed tcgen05
kind::f16 (bf16 x bf16 products are exact in fp32, accumulated in fp32 in TMEM), parameters and parameter gradients stay fp32.
  op level:  against an fp64 convolution of the SAME bf16-rounded operands -- the only differences left are the fp32
             accumulation order and the final bf16 rounding of the stored tensor (2^-9 relative, rel-L2 ~1.2e-3): gate 3e-3
             on bf16 tensors, 1e-3 on fp32 results (dW, db, 3-channel outputs).
  net level: against the reference nets under torch.autocast(bfloat16) (same storage rounding points) and against the fp32
             oracle, at the measured bf16 error level (2.4e-3 per layer; ~sqrt(depth) growth).
"""
import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu

import srb200
from srb200 import models as M
from srb200 import functional as F
from oracle import torch_ref as R
from util import rel_l2

DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _reset():
    srb200.set_math("auto")
    srb200.set_grad_scale(1.0)
    yield
    srb200.set_math("auto")
    F.record_activation_masks(None)


def bf(t):
    return t.to(torch.bfloat16).to(t.dtype)


CASES = [
    # Cin Cout k  p  act      res    ps  H   W   N
    (256, 256, 3, 1, "relu", False, 1, 12, 12, 2),    # EDSR-256 body conv1
    (256, 256, 3, 1, None, True, 1, 10, 14, 2),       # EDSR-256 body conv2 + skip
    (256, 256, 3, 1, None, False, 2, 8, 8, 2),        # upsampler 256 -> 1024 -> PS2
    (3, 256, 3, 1, None, False, 1, 12, 12, 2),        # input conv (fp32 in, bf16 out)
    (256, 3, 3, 1, None, False, 1, 16, 16, 2),        # output conv (bf16 in, fp32 out)
    (64, 64, 3, 1, "relu", False, 1, 16, 16, 2),      # 1 ci-block (tap-pair wgrad)
    (64, 64, 3, 1, "prelu", False, 2, 8, 8, 2),       # SRGAN upsampler with PReLU
    (128, 64, 3, 1, "lrelu", False, 1, 9, 11, 1),
    (64, 128, 5, 2, "relu", False, 1, 12, 12, 1),     # k5, pad 2
    (256, 256, 3, 1, "relu", False, 1, 6, 140, 1),    # wide rows (column-tiled wgrad bands)
    (192, 96, 3, 1, None, False, 1, 9, 9, 2),         # channel counts that are not multiples of 64
]


@pytest.mark.parametrize("case", CASES)
def test_bf16_fused_conv_vs_bf16_operand_truth(case):
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    srb200.set_math("bf16")
    Cin, Cout, k, p, act, res, ps, H, W, N = case
    gen = torch.Generator().manual_seed(31)
    x = torch.randn(N, Cin, H, W, generator=gen)
    w = torch.randn(Cout * ps * ps, Cin, k, k, generator=gen) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout * ps * ps, generator=gen) * 0.1
    alpha = torch.tensor([0.25])
    x_in_bf16 = Cin % 8 == 0
    y_bf16 = Cout % 8 == 0
    if x_in_bf16:
        x = bf(x)  # what the previous layer would have stored
    z0 = TF.conv2d(x, w, b, 1, p)
    oshape = (N, Cout, z0.shape[2] * ps, z0.shape[3] * ps)
    r = bf(torch.randn(oshape, generator=gen)) if res else None
    gy = torch.randn(oshape, generator=gen)
    if y_bf16:
        gy = bf(gy)
    xg = x.to(DEV)
    if x_in_bf16:
        xg = xg.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    xg.requires_grad_(True)
    wg, bg, ag = (t.to(DEV).requires_grad_(True) for t in (w, b, alpha))
    rg = r.to(DEV).to(torch.bfloat16) if res else None
    y = srb200.conv2d(xg, wg, bg, 1, p, activation=act, alpha=ag if act == "prelu" else None, residual=rg, pixel_shuffle=ps)
    assert y.dtype == (torch.bfloat16 if y_bf16 else torch.float32)
    y.backward(gy.to(DEV).to(y.dtype))
    # truth: fp64 conv of the operands the tensor cores saw (x as stored, w rounded to bf16 / tf32), fp32 bias
    wq = bf(w) if x_in_bf16 else w
    xr, wr, br = (t.double().requires_grad_(True) for t in (x, wq, b))
    zr = TF.conv2d(xr, wr, br, 1, p)
    if ps > 1:
        zr = TF.pixel_shuffle(zr, ps)
    if act is not None:
        yg = y.detach().float().cpu().double() - (r.double() if res else 0)
        m = (yg > 0).double()
        slope = {"relu": 0.0, "lrelu": 0.2, "prelu": 0.25}[act]
        yr = zr * m + zr * (1 - m) * slope
    else:
        yr = zr
    if res:
        yr = yr + r.double()
    yr.backward(gy.double())
    t_store = 3e-3   # tensors stored in bf16 (one final rounding, 2^-9)
    t_f32 = 1e-3     # fp32 results of bf16 / tf32 products
    assert rel_l2(y.detach().float(), yr.detach()) < (t_store if y_bf16 else t_f32)
    # backward truth uses dz as the kernels saw it: for bf16 y the activation backward / un-shuffle re-round dz to bf16 only
    # when an activation is applied (dz = dy * act'), which is exact for relu and one more rounding for prelu/lrelu
    # (the dgrad of a bf16 dz multiplies bf16-rounded weights even when dx itself is an fp32 3-channel tensor)
    assert rel_l2(xg.grad.float(), xr.grad) < t_store
    assert rel_l2(wg.grad, wr.grad) < (4e-3 if act in ("prelu", "lrelu") else t_f32 * 2)
    assert rel_l2(bg.grad, br.grad) < (4e-3 if act in ("prelu", "lrelu") else t_f32 * 2)


def test_bf16_pixel_shuffle_permutation_bit_exact():
    """PixelShuffle(2) fused into the bf16 epilogue store and the bf16 un-shuffle of the gradient: exact permutations."""
    assert torch.cuda.is_available()
    srb200.set_math("bf16")
    C, r, H, W = 64, 2, 6, 5
    cin = C * r * r
    x = torch.randint(-64, 64, (2, cin, H, W)).float()
    w = torch.eye(cin).reshape(cin, cin, 1, 1)
    xg = x.to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = srb200.conv2d(xg, w.to(DEV), None, 1, 0, pixel_shuffle=r)
    assert y.dtype == torch.bfloat16
    assert torch.equal(y.float().cpu(), TF.pixel_shuffle(x, r))
    gy = torch.randint(-8, 8, tuple(y.shape)).float()
    y.backward(gy.to(DEV).to(torch.bfloat16))
    assert torch.equal(xg.grad.float().cpu(), TF.pixel_unshuffle(gy, r))


@pytest.mark.parametrize("name,args,xshape,loss_kind", [
    ("edsr", (3, 64, 4), (2, 3, 16, 16), "l1"),
    ("edsr", (3, 256, 32), (1, 3, 12, 10), "l1"),     # BASELINE cfg4 depth and width
    ("vdsr", (3, 64, 18), (2, 3, 24, 20), "mse"),
])
def test_bf16_net_vs_autocast_reference(name, args, xshape, loss_kind):
    assert torch.cuda.is_available()
    ref = R.build(name, args, seed=0)
    ref.train()
    gen = torch.Generator().manual_seed(1)
    x = torch.rand(xshape, generator=gen)
    y32 = ref(x).detach()
    tgt = torch.rand(y32.shape, generator=torch.Generator().manual_seed(2))
    lossf = TF.l1_loss if loss_kind == "l1" else TF.mse_loss
    # the reference under autocast(bf16): conv operands and outputs rounded to bf16 at the same points
    ref.zero_grad()
    with torch.autocast("cpu", dtype=torch.bfloat16):
        yac = ref(x)
    lossf(yac.float(), tgt).backward()
    gac = {k: p.grad.clone() for k, p in ref.named_parameters()}
    srb200.set_math("bf16")
    net = M.MODELS[name](*args)
    net.load_state_dict(ref.state_dict())
    net.to(DEV).train()
    l0 = srb200.launch_count()
    y = net(x.to(DEV))
    assert y.dtype == torch.float32 and srb200.launch_count() > l0
    lossf(y, tgt.to(DEV)).backward()
    torch.cuda.synchronize()
    n_convs = R.count_convs(ref)
    e32, eac = rel_l2(y.detach(), y32), rel_l2(y.detach(), yac.detach().float())
    eself = rel_l2(yac.detach().float(), y32)
    import numpy as np
    cos = []
    for k, p in net.named_parameters():
        a, b = p.grad.detach().double().cpu().flatten(), gac[k].double().flatten()
        if b.norm() > 0:
            cos.append((a @ b / (a.norm() * b.norm() + 1e-300)).item())
    print("\nbf16 %s%s: y vs fp32 oracle %.3e, vs autocast(bf16) oracle %.3e (autocast oracle vs fp32 oracle: %.3e), %d convs, "
          "min grad cosine vs autocast %.4f" % (name, args, e32, eac, eself, n_convs, min(cos)))
    # bf16 storage error level: 2.4e-3 per layer, random-walk growth with depth; the reference's own
    # autocast run sits at the same distance from fp32 (eself), which is the yardstick
    tol = 2.5e-3 * n_convs ** 0.5
    assert e32 < max(tol, 2.0 * eself)
    assert eac < max(tol, 2.0 * eself)
    assert min(cos) > 0.98
