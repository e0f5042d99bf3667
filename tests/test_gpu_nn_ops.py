"""GPU suite (-m gpu): the non-conv layers of SRGAN on libsrb200 -- BatchNorm2d (+activation, +residual), Linear, MaxPool2d(2),
BCELoss -- against torch on CPU (fp32 arithmetic on both sides: 1e-5 class), the VGG feature extractor and one full SRGAN
adversarial step (srgan.py:256-310) against the CPU oracle."""
import copy

import numpy as np

import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu

import srb200
from srb200 import models as M, host, nn_ops
from oracle import torch_ref as R
from util import rel_l2

DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _reset():
    srb200.set_math("fp32")
    srb200.set_grad_scale(1.0)
    yield
    srb200.set_math("auto")


@pytest.mark.parametrize("act", [None, "relu", "lrelu", "prelu"])
@pytest.mark.parametrize("res", [False, True])
@pytest.mark.parametrize("shape", [(4, 64, 9, 7), (2, 128, 5, 5), (3, 8, 6, 6)])
def test_batch_norm_act_matches_torch(act, res, shape):
    assert torch.cuda.is_available()
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(shape, generator=gen) * 2 + 0.5
    r = torch.randn(shape, generator=gen) if res else None
    gy = torch.randn(shape, generator=gen)
    C = shape[1]
    bn_ref = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        bn_ref.weight.normal_(1.0, 0.2, generator=gen)
        bn_ref.bias.normal_(0.0, 0.2, generator=gen)
    bn = copy.deepcopy(bn_ref).to(DEV)
    alpha_ref = torch.tensor([0.25], requires_grad=True)
    alpha = alpha_ref.detach().clone().to(DEV).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    rr = r.clone().requires_grad_(True) if res else None
    z = bn_ref(xr)
    if act == "relu":
        z = TF.relu(z)
    elif act == "lrelu":
        z = TF.leaky_relu(z, 0.2)
    elif act == "prelu":
        z = TF.prelu(z, alpha_ref)
    if res:
        z = z + rr
    z.backward(gy)
    xg = x.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    rg = r.to(DEV).requires_grad_(True) if res else None
    y = srb200.batch_norm_act(xg, bn, activation=act, alpha=alpha, residual=rg)
    y.backward(gy.to(DEV))
    assert rel_l2(y.detach(), z.detach()) < 2e-5
    assert rel_l2(xg.grad, xr.grad) < 5e-5
    assert rel_l2(bn.weight.grad, bn_ref.weight.grad) < 5e-5
    assert rel_l2(bn.bias.grad, bn_ref.bias.grad) < 5e-5
    assert rel_l2(bn.running_mean, bn_ref.running_mean) < 1e-5 and rel_l2(bn.running_var, bn_ref.running_var) < 1e-5
    assert int(bn.num_batches_tracked) == int(bn_ref.num_batches_tracked) == 1
    if act == "prelu":
        assert abs(alpha.grad.item() - alpha_ref.grad.item()) < 1e-4 * max(1.0, abs(alpha_ref.grad.item()))
    if res:
        assert torch.equal(rg.grad.cpu(), gy)
    # eval mode uses the running statistics
    bn.eval(); bn_ref.eval()
    assert rel_l2(srb200.batch_norm_act(xg.detach(), bn), bn_ref(x)) < 2e-5


@pytest.mark.parametrize("B,I,O", [(16, 32768, 1024), (16, 1024, 1), (5, 2048, 37), (20, 512, 64)])
def test_linear_matches_torch(B, I, O):
    assert torch.cuda.is_available()
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(B, I, generator=gen)
    w = torch.randn(O, I, generator=gen) / I ** 0.5
    b = torch.randn(O, generator=gen)
    gy = torch.randn(B, O, generator=gen)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    TF.linear(xr, wr, br).backward(gy)
    xg, wg, bg = (t.to(DEV).requires_grad_(True) for t in (x, w, b))
    y = srb200.linear(xg, wg, bg)
    y.backward(gy.to(DEV))
    assert rel_l2(y.detach(), TF.linear(x, w, b)) < 2e-5
    assert rel_l2(xg.grad, xr.grad) < 2e-5
    assert rel_l2(wg.grad, wr.grad) < 2e-5
    assert rel_l2(bg.grad, br.grad) < 2e-5


def test_max_pool_and_bce_match_torch():
    assert torch.cuda.is_available()
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(3, 64, 10, 12, generator=gen)
    x[0, 0, 0, 0] = x[0, 0, 0, 1] = 5.0  # a tie: the first maximum wins
    gy = torch.randn(3, 64, 5, 6, generator=gen)
    xr = x.clone().requires_grad_(True)
    TF.max_pool2d(xr, 2).backward(gy)
    xg = x.to(DEV).requires_grad_(True)
    y = srb200.max_pool2(xg)
    y.backward(gy.to(DEV))
    assert torch.equal(y.detach().cpu(), TF.max_pool2d(x, 2))
    assert torch.equal(xg.grad.cpu(), xr.grad)
    p = torch.rand(16, 1, generator=gen) * 0.98 + 0.01
    t = (torch.rand(16, generator=gen) > 0.5).float()
    pr = p.clone().requires_grad_(True)
    (3.0 * TF.binary_cross_entropy(pr, t.reshape(16, 1))).backward()
    pg = p.to(DEV).requires_grad_(True)
    l = srb200.bce_loss(pg, t.to(DEV))
    (3.0 * l).backward()
    assert abs(l.item() - TF.binary_cross_entropy(p, t.reshape(16, 1)).item()) < 1e-6
    assert rel_l2(pg.grad, pr.grad) < 1e-6


def test_feature_extractor_matches_oracle():
    assert torch.cuda.is_available()
    ref = R.build_feature_extractor()
    fe = M.FeatureExtractor()
    fe.load_state_dict(ref.state_dict())
    fe.to(DEV)
    x = torch.rand(2, 3, 32, 32, generator=torch.Generator().manual_seed(6))
    xr = x.clone().requires_grad_(True)
    gy = torch.randn(2, 128, 16, 16, generator=torch.Generator().manual_seed(7))
    ref(xr).backward(gy)
    xg = x.to(DEV).requires_grad_(True)
    y = fe(xg)
    y.backward(gy.to(DEV))
    assert rel_l2(y.detach(), ref(x).detach()) < 1e-4
    assert rel_l2(xg.grad, xr.grad) < 1e-4
    for (k, p), (_, q) in zip(fe.named_parameters(), ref.named_parameters()):
        assert rel_l2(p.grad, q.grad) < 1e-4, k


@pytest.mark.parametrize("math", ["fp32", "exact", "auto"])
def test_srgan_adversarial_step_tracks_the_oracle(math):
    """srgan.py:256-310 on the engine: both losses and every parameter of G and D after one step against the CPU oracle."""
    assert torch.cuda.is_available()
    srb200.set_math(math)
    Gr, Dr, FEr = R.build("srgan_g", (3, 32, 2), seed=0), R.build("srgan_d", (3, 16, 32), seed=1), R.build_feature_extractor()
    G, D, FE = M.SRGANGenerator(3, 32, 2), M.SRGANDiscriminator(3, 16, 32), M.FeatureExtractor()
    G.load_state_dict(Gr.state_dict()); D.load_state_dict(Dr.state_dict()); FE.load_state_dict(FEr.state_dict())
    for m in (G, D, FE):
        m.to(DEV).train()
    lr_ = 1e-3
    gor, dor = R.make_srgan_optimizers(Gr, Dr, lr=lr_)
    go, do = host.make_srgan_optimizers(G, D, lr=lr_)
    gen = torch.Generator().manual_seed(8)
    lr_img, hr_img = torch.rand(4, 3, 8, 8, generator=gen), torch.rand(4, 3, 32, 32, generator=gen)
    dl_r, gl_r = R.srgan_step(Gr, Dr, FEr, gor, dor, lr_img, hr_img)
    l0 = srb200.launch_count()
    dl, gl = host.srgan_step(G, D, FE, go, do, lr_img.to(DEV), hr_img.to(DEV))
    assert srb200.launch_count() - l0 > 100
    tol = {"fp32": 1e-4, "exact": 2e-4, "auto": 2e-3}[math]
    assert abs(dl.item() - dl_r.item()) < tol * abs(dl_r.item())
    assert abs(gl.item() - gl_r.item()) < tol * abs(gl_r.item())
    # D (SGD): parameters move by lr/100 * grad -- compare the updates; G (Adam): sign-like steps of size lr, compared on the
    # parameters themselves (a rounding-level gradient flips a +-lr step, so the bound is a few lr relative to |p| ~ 0.02)
    d0 = R.build("srgan_d", (3, 16, 32), seed=1).state_dict()
    upd = {k: (q.detach() - d0[k]) for k, q in Dr.named_parameters()}
    U = max(u.abs().max().item() for u in upd.values())  # largest SGD update of the step
    for (k, p), (_, q) in zip(D.named_parameters(), Dr.named_parameters()):
        # biases start at exactly 0 and the ones in front of a BatchNorm have a mathematically zero gradient: compare the
        # UPDATES on the scale of the step's largest update instead of relative to the (near-zero) parameter
        # TF32 bound: D's first-layer bias gradient is a heavily cancelling sum of dz behind LeakyReLU sign flips and 16-sample
        # BatchNorm statistics; the CPU oracle itself moves by 0.11 U there when its conv operands are rounded to tf32
        # (0.036 U on the first-layer weights, <= 0.013 U elsewhere), so `auto` is gated at 0.25 U and the fp32 / 3xTF32
        # modes carry the tight bound.
        assert (p.detach().cpu() - q.detach()).abs().max().item() <= {"fp32": 5e-3, "exact": 1e-2, "auto": 0.25}[math] * U, k
    for (k, p), (_, q) in zip(G.named_parameters(), Gr.named_parameters()):
        assert (p.detach().cpu() - q.detach()).abs().max().item() <= 2.1 * lr_, k


# ---- utils.img_interp / utils.shave on the device (utils.py:197-205, 242-269) ---------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", ["x2", "x3", "x4", "half", "gray_x4"])
def test_img_interp_matches_reference_golden(case):
    import os
    from srb200 import host
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pil_bicubic.npz"))
    y = host.img_interp(torch.from_numpy(g[case + ":x"]).cuda(), float(g[case + ":scale"]))
    assert torch.equal(y.cpu(), torch.from_numpy(g[case + ":y"]))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,sf,border", [((4, 3, 32, 32), 4, 0), ((2, 3, 17, 23), 3, 0), ((3, 3, 24, 24), 2, 6),
                                             ((2, 1, 40, 28), 0.5, 2), ((0, 3, 8, 8), 2, 0)])
def test_img_interp_matches_oracle(shape, sf, border):
    """Bit-exact against the Pillow restatement (non-square, downscale, empty batch), optionally fused with shave."""
    from oracle import pil_bicubic as PB
    from srb200 import host
    x = torch.rand(shape, generator=torch.Generator().manual_seed(5))
    if shape[0]:
        x[0, :, : shape[2] // 2] = (x[0, :, : shape[2] // 2] > 0.5).float()  # saturated edges: clip to [0, 255]
    want = PB.img_interp(x, sf) if shape[0] else torch.empty((0, shape[1], int(shape[2] * sf), int(shape[3] * sf)))
    if border:
        want = want[..., border:-border, border:-border]
    got = host.img_interp(x.cuda(), sf, shave=border)
    assert got.shape == want.shape and torch.equal(got.cpu(), want)
    assert torch.equal(host.shave(got, 1).cpu(), want[..., 1:-1, 1:-1])


@pytest.mark.gpu
def test_img_interp_full_size_properties():
    """cfg3-sized batch (vdsr.py:137: 64 LR crops x4 -> 128^2): constant images stay constant, output lies on the 1/255 grid,
    and the batch result equals the per-image results (no cross-image reads)."""
    from srb200 import host
    x = torch.rand((64, 3, 32, 32), device="cuda", generator=torch.Generator("cuda").manual_seed(9))
    x[1] = 0.5
    y = host.img_interp(x, 4)
    assert y.shape == (64, 3, 128, 128)
    assert torch.equal(y[1], torch.full_like(y[1], float(int(0.5 * 255)) / 255.0))
    assert ((y * 255) - (y * 255).round()).abs().max().item() < 1e-3
    assert torch.equal(host.img_interp(x[5:7].clone(), 4), y[5:7])


# ---- srb200.FlatAdam: torch.optim.Adam (espcn.py:79, edsr.py:93) as one launch over flat buffers -------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_flat_adam_matches_torch_adam(wd):
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.PReLU(), torch.nn.Conv2d(8, 4, 3, bias=False)).to(DEV)
    ref = copy.deepcopy(net)
    bucket = srb200.GradBucket(net, world_size=1, direct=False)
    opt = srb200.FlatAdam(bucket, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    for (p, q) in zip(net.parameters(), ref.parameters()):
        assert torch.equal(p, q)  # re-seating the parameters as views keeps their values
    gen = torch.Generator(device="cuda").manual_seed(4)
    for step in range(5):
        x = torch.randn(2, 3, 9, 9, device=DEV, generator=gen)
        bucket.begin_step()
        ropt.zero_grad()
        net(x).square().mean().backward()
        ref(x).square().mean().backward()
        opt.step()
        ropt.step()
        for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
            assert (p - q).abs().max().item() <= 2e-6 * max(1.0, q.abs().max().item()), (step, k)
    assert opt.step_count == 5
    sd = opt.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    rs = ropt.state_dict()["state"]
    for i in range(len(list(net.parameters()))):
        assert rel_l2(sd["state"][i]["exp_avg"], rs[i]["exp_avg"]) < 1e-4
        assert rel_l2(sd["state"][i]["exp_avg_sq"], rs[i]["exp_avg_sq"]) < 1e-4  # the two nets' gradients differ at the 1e-6 level (cuDNN vs autograd order)


@pytest.mark.gpu
def test_flat_adam_is_graph_capturable_and_counts_replays():
    torch.manual_seed(5)
    lin = torch.nn.Conv2d(2, 2, 1).to(DEV)
    bucket = srb200.GradBucket(lin, world_size=1, direct=False)
    opt = srb200.FlatAdam(bucket, lr=1e-3)
    bucket.flat.fill_(0.5)
    opt.step()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            opt.step()
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert opt.step_count == 4  # one eager step + three replays (the capture itself executes nothing): the counter lives on the device


@pytest.mark.gpu
def test_srgan_step_with_grad_buckets_matches_plain_autograd():
    """srgan.py:256-310 with the flat gradient buckets (direct wgrad writes, modules applied several times before one backward:
    D sees real, fake and fake again) against the same step with ordinary autograd accumulation: identical parameter updates."""
    srb200.set_math("fp32")

    def run(with_buckets):
        torch.manual_seed(0)
        G, D, FE = M.SRGANGenerator(3, 32, 2), M.SRGANDiscriminator(3, 16, 32), M.FeatureExtractor()
        host.init_model("srgan", G); host.init_model("srgan", D)
        for m in (G, D, FE):
            m.to(DEV).train()
        go, do = host.make_srgan_optimizers(G, D, lr=1e-3)
        bg = srb200.GradBucket(G, world_size=1) if with_buckets else None
        bd = srb200.GradBucket(D, world_size=1) if with_buckets else None
        gen = torch.Generator().manual_seed(8)
        lr_img, hr_img = torch.rand(4, 3, 8, 8, generator=gen).to(DEV), torch.rand(4, 3, 32, 32, generator=gen).to(DEV)
        for _ in range(2):
            dl, gl = host.srgan_step(G, D, FE, go, do, lr_img, hr_img, bucket_g=bg, bucket_d=bd)
        if bg is not None:
            bg.detach(); bd.detach()
        return dl.item(), gl.item(), [p.detach().clone() for p in list(G.parameters()) + list(D.parameters())]

    dl0, gl0, p0 = run(False)
    dl1, gl1, p1 = run(True)
    assert abs(dl0 - dl1) <= 1e-5 * abs(dl0) and abs(gl0 - gl1) <= 1e-5 * abs(gl0)
    for a, b in zip(p0, p1):
        assert (a - b).abs().max().item() <= 1e-5 * max(1e-3, a.abs().max().item())


@pytest.mark.parametrize("shape", [(5, 16, 24, 3), (3, 9, 8, 1), (2, 7, 12, 4), (2, 6, 20, 2), (4, 10, 13, 3), (1, 4, 4, 5), (0, 8, 8, 3)],
                         ids=lambda s: "x".join(map(str, s)))
def test_image_to_tensor_matches_totensor(shape):
    """srb200.image_to_tensor == torchvision ToTensor (dataset.py:90 of the reference), bit for bit: uint8 HWC -> float CHW / 255;
    W % 4 == 0 with C <= 4 takes the 4-pixels-per-thread kernel, everything else (odd W, C = 5, unaligned views) the scalar one."""
    g = torch.Generator().manual_seed(3)
    img = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8).to(DEV)
    # exactly what ToTensor does ON THE HOST (dataset.py:90): a true division (torch's CUDA div-by-scalar multiplies by 1/255
    # and is 1 ulp off for 126 byte values -- the library kernel is not)
    want = img.cpu().permute(0, 3, 1, 2).float().div(255).contiguous().to(DEV)
    got = srb200.image_to_tensor(img)
    assert got.shape == want.shape and torch.equal(got, want)
    if shape[0] > 1:  # a batch slice that starts at an odd byte offset when N*H*W*C is odd: falls back, still exact
        sub = img[1:]
        assert torch.equal(srb200.image_to_tensor(sub), want[1:])
    out = torch.full_like(want, -1.0)
    srb200.image_to_tensor(img, out=out)
    assert torch.equal(out, want)
