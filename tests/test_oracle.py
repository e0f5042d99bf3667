"""CPU suite: pins the oracle (oracle/torch_ref.py, oracle/conv_ref.c) against the golden vectors generated
from the unmodified reference (oracle/make_golden.py) and, when /root/reference is present, against the
live reference classes."""
import numpy as np
import pytest
import torch
import torch.nn.functional as TF

from oracle import torch_ref as R
from oracle import conv_ref as C
from oracle import ref_import
from util import load_golden, load_state, rel_l2

NETS = {"srcnn": "l2", "espcn": "l2", "fsrcnn": "l2", "vdsr": "l2", "edsr": "l1", "srgan_g": "l2"}


@pytest.mark.parametrize("name", sorted(NETS))
def test_oracle_net_matches_golden(name):
    g = load_golden("net_" + name)
    torch.set_num_threads(1)
    net = R.build(name, tuple(int(v) for v in g["args"]), init=False)
    load_state(net, g)
    net.train()
    x = torch.from_numpy(g["x"])
    y = net(x)
    tgt = torch.from_numpy(g["target"])
    loss = TF.l1_loss(y, tgt) if NETS[name] == "l1" else TF.mse_loss(y, tgt)
    loss.backward()
    # same torch build, same op sequence -> equal up to thread-count dependent summation order
    assert rel_l2(y.detach(), g["y"]) < 1e-6
    assert abs(loss.item() - float(g["loss"])) < 1e-6 * max(1.0, abs(float(g["loss"])))
    for k, p in net.named_parameters():
        assert rel_l2(p.grad, g["grad:" + k]) < 1e-5, k


@pytest.mark.parametrize("name", sorted(NETS))
def test_oracle_init_matches_golden(name):
    """weight_init restatement: same seed -> same parameters as the reference's own weight_init()."""
    g = load_golden("net_" + name)
    net = R.build(name, tuple(int(v) for v in g["args"]), seed=0)
    for k, v in net.state_dict().items():
        assert torch.equal(v, torch.from_numpy(g["param:" + k])), k


BLOCKS = {
    "convblock_k3_relu": lambda: R.ConvBlock(8, 16, 3, 1, 1, activation="relu", norm=None),
    "convblock_k5_p0_prelu": lambda: R.ConvBlock(3, 12, 5, 1, 0, activation="prelu", norm=None),
    "convblock_k3_s2_lrelu": lambda: R.ConvBlock(6, 10, 3, 2, 1, activation="lrelu", norm=None),
    "convblock_default_k4s2_nobias": lambda: R.ConvBlock(4, 8, bias=False, activation=None, norm=None),
    "psblock_r4": lambda: R.PSBlock(8, 3, 4, 3, 1, 0, activation=None, norm=None),
    "psblock_r2_prelu": lambda: R.PSBlock(8, 8, 2, activation="prelu", norm=None),
    "resnetblock_relu": lambda: R.ResnetBlock(8, norm=None),
    "resnetblock_prelu": lambda: R.ResnetBlock(8, activation="prelu", norm=None),
    "deconvblock_k4s2": lambda: R.DeconvBlock(6, 4, activation="relu", norm=None),
    "upsample2x_ps": lambda: R.Upsample2xBlock(8, 8, upsample="ps", activation=None, norm=None),
    "upsample2x_deconv": lambda: R.Upsample2xBlock(4, 4, upsample="deconv", activation="lrelu", norm=None),
    "fsrcnn_tail": lambda: torch.nn.Sequential(torch.nn.PReLU(), torch.nn.ConvTranspose2d(6, 3, 9, 4, 3, output_padding=1)),
}


@pytest.mark.parametrize("name", sorted(BLOCKS))
def test_oracle_block_matches_golden(name):
    g = load_golden("block_" + name)
    blk = BLOCKS[name]()
    load_state(blk, g)
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    y = blk(x)
    y.backward(torch.from_numpy(g["gy"]))
    assert rel_l2(y.detach(), g["y"]) < 1e-6
    assert rel_l2(x.grad, g["gx"]) < 1e-6
    for k, p in blk.named_parameters():
        assert rel_l2(p.grad, g["grad:" + k]) < 1e-5, k


# ---- plain-C restatement vs torch (ATen) on CPU -------------------------------------------------
CONV_CASES = [  # N, C, H, W, O, k, st, pad
    (2, 3, 9, 8, 5, 3, 1, 1), (1, 4, 10, 10, 6, 5, 1, 0), (2, 2, 11, 9, 3, 3, 2, 1), (1, 3, 12, 12, 4, 9, 1, 4),
    (1, 5, 7, 7, 2, 1, 1, 0), (2, 4, 8, 8, 4, 4, 2, 1),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_c_conv_matches_aten(case):
    N, Cc, H, W, O, k, st, pad = case
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(N, Cc, H, W, generator=gen, requires_grad=True)
    w = torch.randn(O, Cc, k, k, generator=gen, requires_grad=True)
    b = torch.randn(O, generator=gen, requires_grad=True)
    y = TF.conv2d(x, w, b, st, pad)
    gy = torch.randn(y.shape, generator=gen)
    y.backward(gy)
    yc = C.conv2d_fwd(x.detach().numpy(), w.detach().numpy(), b.detach().numpy(), st, pad)
    assert rel_l2(yc, y.detach()) < 2e-6
    dxc = C.conv2d_bwd_data(gy.numpy(), w.detach().numpy(), tuple(x.shape), st, pad)
    assert rel_l2(dxc, x.grad) < 2e-6
    dwc, dbc = C.conv2d_bwd_weight(x.detach().numpy(), gy.numpy(), tuple(w.shape), st, pad)
    assert rel_l2(dwc, w.grad) < 2e-6
    assert rel_l2(dbc, b.grad) < 2e-6


@pytest.mark.parametrize("st,pad,op,k", [(2, 1, 0, 4), (4, 3, 1, 9), (1, 0, 0, 3), (3, 2, 2, 5)])
def test_c_conv_transpose_matches_aten(st, pad, op, k):
    gen = torch.Generator().manual_seed(6)
    x = torch.randn(2, 5, 6, 7, generator=gen)
    w = torch.randn(5, 3, k, k, generator=gen)
    b = torch.randn(3, generator=gen)
    y = TF.conv_transpose2d(x, w, b, st, pad, op)
    yc = C.conv_transpose2d_fwd(x.numpy(), w.numpy(), b.numpy(), st, pad, op)
    assert yc.shape == tuple(y.shape)
    assert rel_l2(yc, y) < 2e-6


@pytest.mark.parametrize("r", [2, 3, 4])
def test_pixel_shuffle_law_bit_exact(r):
    t = torch.randn(2, 3 * r * r, 5, 4)
    ref = torch.nn.PixelShuffle(r)(t)
    assert torch.equal(R.pixel_shuffle_law(t, r), ref)
    assert np.array_equal(C.pixel_shuffle(t.numpy(), r), ref.numpy())


def test_c_act_matches_aten():
    gen = torch.Generator().manual_seed(7)
    x = torch.randn(1000, generator=gen)
    x[::17] = 0.0  # the z == 0 corner: slope branch, ReLU grad 0
    gy = torch.randn(1000, generator=gen)
    for act, mod in ((1, torch.nn.ReLU()), (2, torch.nn.PReLU()), (3, torch.nn.LeakyReLU(0.2))):
        xt = x.clone().requires_grad_(True)
        y = mod(xt)
        y.backward(gy)
        a = 0.25 if act == 2 else 0.2
        assert np.array_equal(C.act_fwd(x.numpy(), act, a), y.detach().numpy())
        dx, da = C.act_bwd(x.numpy(), gy.numpy(), act, a)
        assert np.array_equal(dx, xt.grad.numpy())
        if act == 2:
            assert abs(da - mod.weight.grad.item()) < 1e-4 * max(1.0, abs(da))


# ---- oracle vs the live reference (dev container only) ---------------------------------------------
needs_ref = pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present (GPU box)")


@needs_ref
@pytest.mark.parametrize("name,mod,cls,args,xs", [
    ("srcnn", "srcnn", "Net", (3, 64), (1, 3, 24, 24)), ("espcn", "espcn", "Net", (3, 64, 4), (1, 3, 16, 16)),
    ("fsrcnn", "fsrcnn", "Net", (3, 4, 56, 12, 4), (1, 3, 12, 12)), ("vdsr", "vdsr", "Net", (3, 64, 18), (1, 3, 12, 12)),
    ("edsr", "edsr", "Net", (3, 64, 16), (1, 3, 8, 8)), ("srgan_g", "srgan", "Generator", (3, 64, 16), (2, 3, 8, 8)),
    ("srgan_d", "srgan", "Discriminator", (3, 16, 32), (2, 3, 32, 32)),
])
def test_oracle_equals_live_reference(name, mod, cls, args, xs):
    import warnings
    mods = ref_import.load()
    torch.set_num_threads(1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(0)
        ref = getattr(mods[mod], cls)(*args)
        ref.weight_init()
    ours = R.build(name, args, seed=0)
    sd_ref, sd_ours = ref.state_dict(), ours.state_dict()
    assert list(sd_ref.keys()) == list(sd_ours.keys())
    for k in sd_ref:
        assert torch.equal(sd_ref[k], sd_ours[k]), k
    x = torch.rand(xs, generator=torch.Generator().manual_seed(1))
    ya, yb = ref(x), ours(x)
    assert torch.equal(ya, yb)
    ya.square().mean().backward()
    yb.square().mean().backward()
    for (k, pa), (_, pb) in zip(ref.named_parameters(), ours.named_parameters()):
        assert torch.equal(pa.grad, pb.grad), k


# ---- the BASELINE depths (VDSR-18, EDSR 256x32, SRGAN G-16 / D): digest fixtures ------------------------------------
from util import DEEP, digest_close, loss_of  # noqa: E402


@pytest.mark.parametrize("case", sorted(DEEP))
def test_oracle_deep_net_matches_reference_digest(case):
    """Full-depth/width nets of BASELINE configs 3-5: the seeded oracle init reproduces the reference's parameters, and
    its forward/backward reproduces the reference's output, loss and every gradient (digests: norm + 64 samples)."""
    name, loss_kind = DEEP[case]
    g = load_golden("deep_" + case)
    net = R.build(name, tuple(int(v) for v in g["args"]), seed=0)
    for k, v in net.state_dict().items():
        if v.dtype.is_floating_point:
            assert digest_close(v, g, "param:" + k, 1e-12), k
    net.train()
    y = net(torch.from_numpy(g["x"]))
    loss = loss_of(loss_kind, y, torch.from_numpy(g["target"]))
    loss.backward()
    assert rel_l2(y.detach(), g["y"]) < 1e-5
    assert abs(loss.item() - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    gscale = max(float(g[k]) for k in g if k.startswith("grad:") and k.endswith(":norm"))
    for k, p in net.named_parameters():
        if float(g["grad:" + k + ":norm"]) < 1e-5 * gscale:
            continue  # mathematically-zero gradients (bias in front of BatchNorm)
        assert digest_close(p.grad, g, "grad:" + k, 2e-4), k


def test_forced_activation_replay_is_identity_on_own_pattern():
    """oracle.with_forced_activations fed with the net's own sign pattern changes nothing (the tool the GPU gradient gates use)."""
    net = R.build("srgan_g", (3, 16, 2), seed=0)
    x = torch.rand(2, 3, 8, 8, generator=torch.Generator().manual_seed(1))
    masks = []
    hooks = [m.register_forward_hook(lambda mod, i, o: masks.append(o.detach() > 0)) for m in net.modules()
             if isinstance(m, (torch.nn.ReLU, torch.nn.PReLU, torch.nn.LeakyReLU))]
    y = net(x)
    for h in hooks:
        h.remove()
    y.square().mean().backward()
    ref = {k: p.grad.clone() for k, p in net.named_parameters()}
    net.zero_grad()
    R.with_forced_activations(net, masks)
    y2 = net(x)
    y2.square().mean().backward()
    assert rel_l2(y2.detach(), y.detach()) < 1e-6
    for k, p in net.named_parameters():
        assert rel_l2(p.grad, ref[k]) < 1e-5, k


# ---- utils.img_interp (utils.py:242-269): the Pillow 8-bit bicubic resample restated in oracle/pil_bicubic.py -------------------
from oracle import pil_bicubic as PB  # noqa: E402

INTERP_CASES = ["x2", "x3", "x4", "half", "gray_x4"]


def _interp_golden():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "pil_bicubic.npz"))


@pytest.mark.parametrize("case", INTERP_CASES)
def test_pil_bicubic_oracle_matches_reference_golden(case):
    """Bit-exact against vectors produced by the unmodified utils.img_interp (oracle/make_golden.py interp)."""
    g = _interp_golden()
    y = PB.img_interp(torch.from_numpy(g[case + ":x"]), float(g[case + ":scale"]))
    assert torch.equal(y, torch.from_numpy(g[case + ":y"]))


@pytest.mark.parametrize("hw,out", [((13, 17), (52, 68)), ((16, 16), (48, 48)), ((31, 9), (15, 27)), ((8, 40), (8, 80))])
def test_pil_bicubic_oracle_matches_pillow(hw, out):
    """Non-square and mixed up/down scales straight against Pillow (Image.resize, BICUBIC), uint8 in / uint8 out."""
    from PIL import Image
    rng = np.random.default_rng(hw[0] * 100 + hw[1])
    img = rng.integers(0, 256, size=(hw[0], hw[1], 3), dtype=np.uint8)
    img[: hw[0] // 3] = np.where(rng.random((hw[0] // 3, hw[1], 3)) < 0.5, 0, 255)  # hard edges: overshoot + clipping
    want = np.asarray(Image.fromarray(img).resize((out[1], out[0]), Image.BICUBIC))
    got = PB.resize_u8(img, out[0], out[1])
    assert np.array_equal(got, want)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")
def test_pil_bicubic_oracle_equals_live_reference():
    u = ref_import.load()["utils"]
    x = torch.rand((3, 3, 14, 14), generator=torch.Generator().manual_seed(11))
    for sf in (2, 3, 4):
        assert torch.equal(PB.img_interp(x, sf), u.img_interp(x, sf))
    y = torch.rand((2, 3, 20, 20))
    assert torch.equal(u.shave(y, 3), y[..., 3:-3, 3:-3])  # utils.py:197-205


def test_byte_over_255_sequence_is_the_exact_division():
    """The device computes ToTensor's byte / 255 as q0 = b * RN(1/255); rem = fma(-q0, 255, b); q = fma(rem, RN(1/255), q0)
    (csrc/srb_common.cuh byte_over_255).  Exhaustive check in exact rational arithmetic that this is the correctly rounded
    quotient -- i.e. torch's `.div(255)` -- for every byte, and that the plain product is not."""
    from fractions import Fraction as Fr
    import torch

    def rn32(x):
        c = np.float32(float(x))
        cands = sorted([c, np.nextafter(c, np.float32(np.inf)), np.nextafter(c, np.float32(-np.inf))], key=lambda v: abs(Fr(float(v)) - x))
        a, b = cands[0], cands[1]
        if abs(Fr(float(a)) - x) == abs(Fr(float(b)) - x):
            return a if int(np.array([a]).view(np.uint32)[0]) % 2 == 0 else b
        return a

    r = np.float32(1.0 / 255.0)
    assert rn32(Fr(1, 255)) == r
    ref = torch.arange(256, dtype=torch.uint8).float().div(255).numpy()
    plain_off = 0
    for b in range(256):
        q0 = rn32(Fr(b) * Fr(float(r)))
        rem = rn32(Fr(b) - Fr(float(q0)) * 255)
        q = rn32(Fr(float(rem)) * Fr(float(r)) + Fr(float(q0)))
        assert q == rn32(Fr(b, 255)) == ref[b]
        plain_off += int(q0 != ref[b])
    assert plain_off == 126
