"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the UNMODIFIED reference classes.

Run in the dev container (needs /root/reference):   python oracle/make_golden.py
For each of the six networks on the hot path (reduced widths where the default would make the fixture
large) it seeds torch, builds the reference Net, runs the net's own weight_init(), one forward and one
backward of the model's loss on CPU fp32, and stores input, target, every parameter, the output, the
loss and every parameter gradient.  Block-level cases (ConvBlock / PSBlock / ResnetBlock / DeconvBlock /
raw PReLU+ConvTranspose2d) are stored the same way.  The GPU box has no /root/reference: tests there
use only these files.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

# name -> (module, class, ctor args, input shape, loss)
NET_CASES = {
    "srcnn": ("srcnn", "Net", (3, 64), (2, 3, 40, 40), "mse"),
    "espcn": ("espcn", "Net", (3, 64, 4), (2, 3, 24, 24), "mse"),
    "fsrcnn": ("fsrcnn", "Net", (3, 4, 56, 12, 4), (2, 3, 20, 20), "mse"),
    "vdsr": ("vdsr", "Net", (3, 64, 4), (2, 3, 32, 32), "mse"),
    "edsr": ("edsr", "Net", (3, 32, 3), (2, 3, 16, 16), "l1"),
    "srgan_g": ("srgan", "Generator", (3, 32, 2), (2, 3, 16, 16), "mse"),
}


# Full BASELINE depths/widths (BASELINE.json configs 3-5).  Spatial sizes are small -- depth and width are
# what matter for error accumulation.  These nets have up to 43 M parameters, so the fixture stores DIGESTS: the input,
# target, output and loss in full, and per parameter / per gradient its L2 norm, its sum and 64 sampled entries
# (indices drawn by a seeded generator).  The parameters themselves are regenerated at test time by the seeded init.
DEEP_CASES = {
    "vdsr18": ("vdsr", "Net", (3, 64, 18), (2, 3, 24, 20), "mse"),                 # vdsr.py:13-32, cfg3
    "edsr256x32": ("edsr", "Net", (3, 256, 32), (1, 3, 12, 10), "l1"),             # edsr.py:13-45, cfg4
    "srgan_g16": ("srgan", "Generator", (3, 64, 16), (2, 3, 16, 12), "mse"),       # srgan.py:14-42, cfg5
    "srgan_d": ("srgan", "Discriminator", (3, 64, 128), (2, 3, 128, 128), "bce"),  # srgan.py:49-77, cfg5
}
DIGEST_SAMPLES = 64


def digest(t, tag):
    """{tag:norm, tag:sum, tag:sample} of a tensor (sample indices: generator seeded by the element count)."""
    flat = t.detach().reshape(-1).double()
    n = flat.numel()
    idx = torch.randint(0, n, (DIGEST_SAMPLES,), generator=torch.Generator().manual_seed(n % 2147483647))
    return {tag + ":norm": np.float64(flat.norm().item()), tag + ":sum": np.float64(flat.sum().item()),
            tag + ":sample": flat[idx].numpy().astype(np.float64)}


def _run(model, x, loss_kind):
    model.train()
    y = model(x)
    tgt = torch.rand(y.shape, generator=torch.Generator().manual_seed(2))
    if loss_kind == "bce":  # srgan.py:157,276: BCELoss of the decision against the "real" label
        tgt = torch.ones_like(y)
        loss = torch.nn.functional.binary_cross_entropy(y, tgt)
    else:
        loss = torch.nn.functional.l1_loss(y, tgt) if loss_kind == "l1" else torch.nn.functional.mse_loss(y, tgt)
    loss.backward()
    return y, tgt, loss


def main():
    mods = ref_import.load()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    for name, (mod, cls, args, xshape, loss_kind) in NET_CASES.items():
        torch.manual_seed(0)
        net = getattr(mods[mod], cls)(*args)
        net.weight_init()
        if name == "srcnn":
            # the reference's N(0, 1e-3) init makes |y| ~ 1e-5; keep it (relative error is what tests use)
            pass
        x = torch.rand(xshape, generator=torch.Generator().manual_seed(1))
        params = {k: v.detach().clone() for k, v in net.state_dict().items()}
        y, tgt, loss = _run(net, x, loss_kind)
        blob = {"x": x.numpy(), "target": tgt.numpy(), "y": y.detach().numpy(), "loss": np.float64(loss.item()),
                "args": np.array(args, dtype=np.int64)}
        for k, v in params.items():
            blob["param:" + k] = v.numpy()
        for k, p in net.named_parameters():
            blob["grad:" + k] = p.grad.numpy()
        np.savez_compressed(os.path.join(OUT, "net_%s.npz" % name), **blob)
        print(name, "y", tuple(y.shape), "loss", loss.item(), "params", sum(p.numel() for p in net.parameters()))

    for name, (mod, cls, args, xshape, loss_kind) in DEEP_CASES.items():
        torch.manual_seed(0)
        net = getattr(mods[mod], cls)(*args)
        net.weight_init()
        x = torch.rand(xshape, generator=torch.Generator().manual_seed(1))
        blob = {"x": x.numpy(), "args": np.array(args, dtype=np.int64)}
        for k, v in net.state_dict().items():
            if v.dtype.is_floating_point:
                blob.update(digest(v, "param:" + k))
        torch.set_num_threads(8)
        y, tgt, loss = _run(net, x, loss_kind)
        torch.set_num_threads(1)
        blob.update({"target": tgt.numpy(), "y": y.detach().numpy(), "loss": np.float64(loss.item())})
        for k, p in net.named_parameters():
            blob.update(digest(p.grad, "grad:" + k))
        np.savez_compressed(os.path.join(OUT, "deep_%s.npz" % name), **blob)
        print("deep", name, "y", tuple(y.shape), "loss", loss.item(), "params", sum(p.numel() for p in net.parameters()))

    # block-level cases straight from base_networks.py
    B = mods["base_networks"]
    block_cases = {
        "convblock_k3_relu": (lambda: B.ConvBlock(8, 16, 3, 1, 1, activation="relu", norm=None), (2, 8, 9, 11)),
        "convblock_k5_p0_prelu": (lambda: B.ConvBlock(3, 12, 5, 1, 0, activation="prelu", norm=None), (2, 3, 12, 10)),
        "convblock_k3_s2_lrelu": (lambda: B.ConvBlock(6, 10, 3, 2, 1, activation="lrelu", norm=None), (2, 6, 11, 12)),
        "convblock_default_k4s2_nobias": (lambda: B.ConvBlock(4, 8, bias=False, activation=None, norm=None),
                                          (2, 4, 10, 10)),
        "psblock_r4": (lambda: B.PSBlock(8, 3, 4, 3, 1, 0, activation=None, norm=None), (2, 8, 7, 9)),
        "psblock_r2_prelu": (lambda: B.PSBlock(8, 8, 2, activation="prelu", norm=None), (2, 8, 6, 5)),
        "resnetblock_relu": (lambda: B.ResnetBlock(8, norm=None), (2, 8, 7, 7)),
        "resnetblock_prelu": (lambda: B.ResnetBlock(8, activation="prelu", norm=None), (2, 8, 6, 8)),
        "deconvblock_k4s2": (lambda: B.DeconvBlock(6, 4, activation="relu", norm=None), (2, 6, 5, 6)),
        "upsample2x_ps": (lambda: B.Upsample2xBlock(8, 8, upsample="ps", activation=None, norm=None), (1, 8, 5, 5)),
        "upsample2x_deconv": (lambda: B.Upsample2xBlock(4, 4, upsample="deconv", activation="lrelu", norm=None),
                              (1, 4, 5, 4)),
        "fsrcnn_tail": (lambda: torch.nn.Sequential(torch.nn.PReLU(),
                                                    torch.nn.ConvTranspose2d(6, 3, 9, 4, 3, output_padding=1)),
                        (2, 6, 5, 5)),
    }
    for name, (ctor, xshape) in block_cases.items():
        torch.manual_seed(0)
        blk = ctor()
        x = torch.randn(xshape, generator=torch.Generator().manual_seed(1), requires_grad=True)
        y = blk(x)
        gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(3))
        y.backward(gy)
        blob = {"x": x.detach().numpy(), "y": y.detach().numpy(), "gy": gy.numpy(), "gx": x.grad.numpy()}
        for k, v in blk.state_dict().items():
            blob["param:" + k] = v.numpy()
        for k, p in blk.named_parameters():
            blob["grad:" + k] = p.grad.numpy()
        np.savez_compressed(os.path.join(OUT, "block_%s.npz" % name), **blob)
        print(name, tuple(y.shape))


# utils.img_interp (utils.py:242-269) vectors: square images (see ref_import: Scale/Resize tuple order), the scale factors the
# drivers use (srcnn.py:121, vdsr.py:137: x scale_factor; 2, 3, 4) and the 1/r downscale of the test loops.
INTERP_CASES = {"x2": ((2, 3, 16, 16), 2), "x3": ((1, 3, 11, 11), 3), "x4": ((2, 3, 12, 12), 4), "half": ((1, 3, 24, 24), 0.5),
                "gray_x4": ((2, 1, 9, 9), 4)}


def make_interp():
    mods = ref_import.load()
    blob = {}
    for name, (shape, sf) in INTERP_CASES.items():
        x = torch.rand(shape, generator=torch.Generator().manual_seed(7))
        if name == "x2":
            x[0, :, :4] = 1.0  # saturated rows: exercises the clip to 255 on bicubic overshoot
            x[0, :, 4:8] = 0.0
        y = mods["utils"].img_interp(x, sf)
        blob[name + ":x"] = x.numpy()
        blob[name + ":scale"] = np.float64(sf)
        blob[name + ":y"] = y.numpy()
        print("img_interp", name, tuple(x.shape), "->", tuple(y.shape))
    np.savez_compressed(os.path.join(OUT, "pil_bicubic.npz"), **blob)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "interp":
        make_interp()
    else:
        main()
        make_interp()
