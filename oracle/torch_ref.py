"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the conv + PixelShuffle hot path.

Restates, on plain torch.nn (CPU, fp32 or fp64), the reference's building blocks, the six networks
that use them, their weight initialisers and their training-loop bodies.  Nothing under the product
package imports this file; it is used by tests/, __graft_entry__.smoke() and bench.py's CPU baseline
(`--impl reference`, `cpu_baseline`) only.

Where the arithmetic lives: the reference (pure Python, /root/reference) delegates every FLOP of this
path to PyTorch -- torch.nn.Conv2d / ConvTranspose2d / PixelShuffle / PReLU (ATen -> oneDNN on CPU,
cuDNN on GPU).  The reference pins no version (its API use dates it to torch ~0.3); the oracle is
anchored to what is installed here and on the GPU box: torch 2.11.0+cu128.  The ATen conv semantics
are additionally restated without torch in oracle/conv_ref.c and cross-checked in tests/.

PINNING: the reference ships no tests, seeds or golden vectors.  The pins are
(1) tests/golden/*.npz, produced by oracle/make_golden.py from the *unmodified reference classes*
imported in the dev container (oracle/ref_import.py), and (2) tests/test_oracle_vs_reference.py, which
compares this restatement with the live reference bit-for-bit whenever /root/reference is present.

Citations are file:line in /root/reference.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as TF


# --------------------------------------------------------------------------------------------
# building blocks (base_networks.py)
# --------------------------------------------------------------------------------------------
def _activation(kind):
    # base_networks.py:50-60 (identical table in every block)
    table = {"relu": lambda: nn.ReLU(True), "prelu": nn.PReLU, "lrelu": lambda: nn.LeakyReLU(0.2, True),
             "tanh": nn.Tanh, "sigmoid": nn.Sigmoid}
    return table[kind]() if kind in table else None


class _Block(nn.Module):
    """Shared tail of every block: optional norm module `bn`, optional activation module `act`."""

    def _finish_init(self, norm, activation, width, bn_cls, in_cls, norm_arg_for_instance=None):
        self.norm = norm
        which = norm if norm_arg_for_instance is None else norm_arg_for_instance
        if norm == "batch":
            self.bn = bn_cls(width)
        elif which == "instance":
            self.bn = in_cls(width)
        self.activation = activation
        a = _activation(activation)
        if a is not None:
            self.act = a

    def _tail(self, t):
        if self.norm is not None:
            t = self.bn(t)
        return self.act(t) if self.activation is not None else t


class DenseBlock(_Block):  # base_networks.py:4-36
    def __init__(self, input_size, output_size, bias=True, activation="relu", norm="batch"):
        super().__init__()
        self.fc = nn.Linear(input_size, output_size, bias=bias)
        self._finish_init(norm, activation, output_size, nn.BatchNorm1d, nn.InstanceNorm1d)

    def forward(self, x):
        return self._tail(self.fc(x))


class ConvBlock(_Block):  # base_networks.py:39-71
    def __init__(self, input_size, output_size, kernel_size=4, stride=2, padding=1, bias=True, activation="relu",
                 norm="batch"):
        super().__init__()
        self.conv = nn.Conv2d(input_size, output_size, kernel_size, stride, padding, bias=bias)
        self._finish_init(norm, activation, output_size, nn.BatchNorm2d, nn.InstanceNorm2d)

    def forward(self, x):
        return self._tail(self.conv(x))


class DeconvBlock(_Block):  # base_networks.py:74-106
    def __init__(self, input_size, output_size, kernel_size=4, stride=2, padding=1, bias=True, activation="relu",
                 norm="batch"):
        super().__init__()
        self.deconv = nn.ConvTranspose2d(input_size, output_size, kernel_size, stride, padding, bias=bias)
        self._finish_init(norm, activation, output_size, nn.BatchNorm2d, nn.InstanceNorm2d)

    def forward(self, x):
        return self._tail(self.deconv(x))


class ResnetBlock(_Block):  # base_networks.py:109-150
    def __init__(self, num_filter, kernel_size=3, stride=1, padding=1, bias=True, activation="relu", norm="batch"):
        super().__init__()
        self.conv1 = nn.Conv2d(num_filter, num_filter, kernel_size, stride, padding, bias=bias)
        self.conv2 = nn.Conv2d(num_filter, num_filter, kernel_size, stride, padding, bias=bias)
        # :118 tests the ctor argument (not self.norm) for 'instance' -- same object either way
        self._finish_init(norm, activation, num_filter, nn.BatchNorm2d, nn.InstanceNorm2d, norm_arg_for_instance=norm)

    def forward(self, x):
        # ONE bn module is applied after conv1 and again after conv2 (:137,:145); act only after the first (:141-142)
        t = self.conv1(x)
        if self.norm is not None:
            t = self.bn(t)
        if self.activation is not None:
            t = self.act(t)
        t = self.conv2(t)
        if self.norm is not None:
            t = self.bn(t)
        return torch.add(t, x)  # :149


class PSBlock(_Block):  # base_networks.py:153-185
    def __init__(self, input_size, output_size, scale_factor, kernel_size=3, stride=1, padding=1, bias=True,
                 activation="relu", norm="batch"):
        super().__init__()
        self.conv = nn.Conv2d(input_size, output_size * scale_factor ** 2, kernel_size, stride, padding, bias=bias)
        self.ps = nn.PixelShuffle(scale_factor)
        self._finish_init(norm, activation, output_size, nn.BatchNorm2d, nn.InstanceNorm2d, norm_arg_for_instance=norm)

    def forward(self, x):
        return self._tail(self.ps(self.conv(x)))


class Upsample2xBlock(nn.Module):  # base_networks.py:188-214
    def __init__(self, input_size, output_size, bias=True, upsample="deconv", activation="relu", norm="batch"):
        super().__init__()
        if upsample == "deconv":
            self.upsample = DeconvBlock(input_size, output_size, 4, 2, 1, bias=bias, activation=activation, norm=norm)
        elif upsample == "ps":
            self.upsample = PSBlock(input_size, output_size, 2, bias=bias, activation=activation, norm=norm)
        elif upsample == "rnc":
            self.upsample = nn.Sequential(
                nn.Upsample(scale_factor=2, mode="nearest"),
                ConvBlock(input_size, output_size, 3, 1, 1, bias=bias, activation=activation, norm=norm))

    def forward(self, x):
        return self.upsample(x)


class _Blocks:
    """Namespace handed to the net builders (so the same topology can be built over any block set)."""
    DenseBlock, ConvBlock, DeconvBlock, ResnetBlock, PSBlock, Upsample2xBlock = (
        DenseBlock, ConvBlock, DeconvBlock, ResnetBlock, PSBlock, Upsample2xBlock)
    PReLU = nn.PReLU
    ConvTranspose2d = nn.ConvTranspose2d


REF_BLOCKS = _Blocks


# --------------------------------------------------------------------------------------------
# weight initialisers (utils.py:76-113, fsrcnn.py:45-55)
# --------------------------------------------------------------------------------------------
def _init_by_classname(m, conv_like, norm_like):
    # utils.py:76-93 / :96-113 dispatch on substrings of the class name, in this order
    name = type(m).__name__
    for key in ("Linear", "Conv2d", "ConvTranspose2d"):
        if key in name:
            conv_like(m)
            return
    if "Norm" in name:
        norm_like(m)


def _zero_bias(m):
    if m.bias is not None:
        m.bias.data.zero_()


def init_normal(net, mean=0.0, std=0.02):
    """utils.weights_init_normal applied to every module (e.g. espcn.py:27-29)."""
    def conv_like(m):
        m.weight.data.normal_(mean, std)
        _zero_bias(m)

    def norm_like(m):
        m.weight.data.normal_(1.0, 0.02)
        _zero_bias(m)
    for m in net.modules():
        _init_by_classname(m, conv_like, norm_like)


def init_kaiming(net):
    """utils.weights_init_kaming applied to every module (vdsr.py:34-36)."""
    def conv_like(m):
        nn.init.kaiming_normal_(m.weight)
        _zero_bias(m)

    def norm_like(m):
        m.weight.data.normal_(1.0, 0.02)
        _zero_bias(m)
    for m in net.modules():
        _init_by_classname(m, conv_like, norm_like)


def init_fsrcnn(net, mean=0.0, std=0.02):
    """fsrcnn.py:45-55 -- isinstance based; the deconv gets N(0, 1e-4)."""
    for m in net.modules():
        if isinstance(m, nn.Conv2d):
            m.weight.data.normal_(mean, std)
            _zero_bias(m)
        if isinstance(m, nn.ConvTranspose2d):
            m.weight.data.normal_(0.0, 0.0001)
            _zero_bias(m)


# --------------------------------------------------------------------------------------------
# networks.  Attribute names reproduce the reference's state_dict keys.
# --------------------------------------------------------------------------------------------
class SRCNNNet(nn.Module):  # srcnn.py:13-29
    def __init__(self, num_channels, base_filter, B=REF_BLOCKS):
        super().__init__()
        f = base_filter
        self.layers = nn.Sequential(B.ConvBlock(num_channels, f, 9, 1, 0, norm=None),
                                    B.ConvBlock(f, f // 2, 5, 1, 0, norm=None),
                                    B.ConvBlock(f // 2, num_channels, 5, 1, 0, activation=None, norm=None))

    def forward(self, x):
        return self.layers(x)

    def weight_init(self):
        init_normal(self, 0.0, 0.001)  # srcnn.py:27


class ESPCNNet(nn.Module):  # espcn.py:13-29
    def __init__(self, num_channels, base_filter, scale_factor, B=REF_BLOCKS):
        super().__init__()
        f = base_filter
        self.layers = nn.Sequential(B.ConvBlock(num_channels, f, 5, 1, 0, activation="relu", norm=None),
                                    B.ConvBlock(f, f // 2, 3, 1, 0, activation="relu", norm=None),
                                    B.PSBlock(f // 2, num_channels, scale_factor, 3, 1, 0, activation=None, norm=None))

    def forward(self, x):
        return self.layers(x)

    def weight_init(self):
        init_normal(self)  # espcn.py:27-29 (defaults 0, 0.02)


class FSRCNNNet(nn.Module):  # fsrcnn.py:13-55
    def __init__(self, num_channels, scale_factor, d, s, m, B=REF_BLOCKS):
        super().__init__()
        self.first_part = B.ConvBlock(num_channels, d, 5, 1, 0, activation="prelu", norm=None)
        mid = [B.ConvBlock(d, s, 1, 1, 0, activation="prelu", norm=None)]
        mid += [B.ConvBlock(s, s, 3, 1, 1, activation=None, norm=None) for _ in range(m)]
        mid += [B.PReLU(), B.ConvBlock(s, d, 1, 1, 0, activation="prelu", norm=None)]
        self.mid_part = nn.Sequential(*mid)
        self.last_part = B.ConvTranspose2d(d, num_channels, 9, scale_factor, 3, output_padding=1)  # :33

    def forward(self, x):
        return self.last_part(self.mid_part(self.first_part(x)))

    def weight_init(self):
        init_fsrcnn(self)


class VDSRNet(nn.Module):  # vdsr.py:13-36
    def __init__(self, num_channels, base_filter, num_residuals, B=REF_BLOCKS):
        super().__init__()
        f = base_filter
        self.input_conv = B.ConvBlock(num_channels, f, 3, 1, 1, norm=None, bias=False)
        self.residual_layers = nn.Sequential(*[B.ConvBlock(f, f, 3, 1, 1, norm=None, bias=False)
                                               for _ in range(num_residuals)])
        self.output_conv = B.ConvBlock(f, num_channels, 3, 1, 1, activation=None, norm=None, bias=False)

    def forward(self, x):
        return torch.add(self.output_conv(self.residual_layers(self.input_conv(x))), x)  # :27-32

    def weight_init(self):
        init_kaiming(self)


class EDSRNet(nn.Module):  # edsr.py:13-45
    def __init__(self, num_channels, base_filter, num_residuals, B=REF_BLOCKS):
        super().__init__()
        f = base_filter
        self.input_conv = B.ConvBlock(num_channels, f, 3, 1, 1, activation=None, norm=None)
        self.residual_layers = nn.Sequential(*[B.ResnetBlock(f, norm=None) for _ in range(num_residuals)])
        self.mid_conv = B.ConvBlock(f, f, 3, 1, 1, activation=None, norm=None)
        self.upscale4x = nn.Sequential(B.Upsample2xBlock(f, f, upsample="ps", activation=None, norm=None),
                                       B.Upsample2xBlock(f, f, upsample="ps", activation=None, norm=None))
        self.output_conv = B.ConvBlock(f, num_channels, 3, 1, 1, activation=None, norm=None)

    def forward(self, x):
        head = self.input_conv(x)
        body = torch.add(self.mid_conv(self.residual_layers(head)), head)  # :38-42
        return self.output_conv(self.upscale4x(body))

    def weight_init(self):
        init_normal(self)


class SRGANGenerator(nn.Module):  # srgan.py:14-46
    def __init__(self, num_channels, base_filter, num_residuals, B=REF_BLOCKS):
        super().__init__()
        f = base_filter
        self.input_conv = B.ConvBlock(num_channels, f, 9, 1, 4, activation="prelu", norm=None)
        self.residual_layers = nn.Sequential(*[B.ResnetBlock(f, activation="prelu") for _ in range(num_residuals)])
        self.mid_conv = B.ConvBlock(f, f, 3, 1, 1, activation=None)  # norm defaults to 'batch'
        self.upscale4x = nn.Sequential(B.Upsample2xBlock(f, f, upsample="ps", activation="prelu", norm=None),
                                       B.Upsample2xBlock(f, f, upsample="ps", activation="prelu", norm=None))
        self.output_conv = B.ConvBlock(f, num_channels, 9, 1, 4, activation=None, norm=None)

    def forward(self, x):
        head = self.input_conv(x)
        body = torch.add(self.mid_conv(self.residual_layers(head)), head)
        return self.output_conv(self.upscale4x(body))

    def weight_init(self):
        init_normal(self)


class SRGANDiscriminator(nn.Module):  # srgan.py:49-81
    def __init__(self, num_channels, base_filter, image_size, B=REF_BLOCKS):
        super().__init__()
        f = base_filter
        self.image_size = image_size
        self.input_conv = B.ConvBlock(num_channels, f, 3, 1, 1, activation="lrelu", norm=None)
        widths = [(f, f, 2), (f, 2 * f, 1), (2 * f, 2 * f, 2), (2 * f, 4 * f, 1), (4 * f, 4 * f, 2),
                  (4 * f, 8 * f, 1), (8 * f, 8 * f, 2)]
        self.conv_blocks = nn.Sequential(*[B.ConvBlock(a, b, 3, s, 1, activation="lrelu") for a, b, s in widths])
        flat = f * 8 * image_size // 16 * image_size // 16  # :66, evaluated left to right exactly as written
        self.dense_layers = nn.Sequential(B.DenseBlock(flat, f * 16, activation="lrelu", norm=None),
                                          B.DenseBlock(f * 16, 1, activation="sigmoid", norm=None))

    def forward(self, x):
        t = self.conv_blocks(self.input_conv(x))
        return self.dense_layers(t.view(t.size()[0], -1))

    def weight_init(self):
        init_normal(self)


# name -> (constructor, default ctor args as instantiated by the reference drivers, input scale rule)
NETS = {
    "srcnn": (SRCNNNet, (3, 64)),                 # srcnn.py:73
    "espcn": (ESPCNNet, (3, 64, 4)),              # espcn.py:73 with scale_factor=4 (main.py:26)
    "fsrcnn": (FSRCNNNet, (3, 4, 56, 12, 4)),     # fsrcnn.py:99
    "vdsr": (VDSRNet, (3, 64, 18)),               # vdsr.py:80
    "edsr": (EDSRNet, (3, 64, 16)),               # edsr.py:87   (BASELINE cfg4: (3, 256, 32))
    "srgan_g": (SRGANGenerator, (3, 64, 16)),     # srgan.py:136
    "srgan_d": (SRGANDiscriminator, (3, 64, 128)),
}


def build(name, args=None, B=REF_BLOCKS, seed=0, init=True):
    """Construct + weight_init under a fixed seed (the reference has no seeds; tests pin seed 0)."""
    cls, default = NETS[name]
    torch.manual_seed(seed)
    net = cls(*(default if args is None else args), B=B)
    if init:
        net.weight_init()
    return net


# --------------------------------------------------------------------------------------------
# training-loop bodies (restated: the reference drivers no longer run on torch 2.x)
# --------------------------------------------------------------------------------------------
def make_optimizer(name, params, lr=1e-5):
    if name == "srcnn":
        return torch.optim.SGD(params, lr=lr)                                   # srcnn.py:79
    if name == "espcn":
        return torch.optim.Adam(params, lr=lr)                                  # espcn.py:79
    if name == "fsrcnn":
        return torch.optim.SGD(params, lr=lr, momentum=0.9)                     # fsrcnn.py:105-106
    if name == "vdsr":
        return torch.optim.SGD(params, lr=lr, momentum=0.9, weight_decay=1e-4)  # vdsr.py:86-90
    if name in ("edsr", "srgan_g"):
        return torch.optim.Adam(params, lr=lr, betas=(0.9, 0.999), eps=1e-8)    # edsr.py:93
    raise KeyError(name)


def loss_fn(name):
    return TF.l1_loss if name == "edsr" else TF.mse_loss  # edsr.py:98 L1; MSE elsewhere (srcnn.py:84 ...)


def train_step(name, model, optimizer, x, target):
    """zero_grad -> forward -> loss -> backward -> (clip) -> step   (espcn.py:126-131, vdsr.py:142-150, edsr.py:150-155)."""
    optimizer.zero_grad()
    out = model(x)
    loss = loss_fn(name)(out, target)
    loss.backward()
    if name == "vdsr":
        nn.utils.clip_grad_norm_(model.parameters(), 0.4)  # vdsr.py:149 (clip = 0.4, vdsr.py:87)
    optimizer.step()
    return loss.detach()


def output_shape(name, args, in_shape):
    """Shape of Net(x) for x of in_shape (N,3,H,W), from the layer arithmetic."""
    n, c, h, w = in_shape
    if name == "srcnn":
        return (n, c, h - 16, w - 16)
    if name == "espcn":
        r = args[2]
        return (n, c, (h - 8) * r, (w - 8) * r)
    if name == "fsrcnn":
        r = args[1]
        hh = (h - 4 - 1) * r - 6 + 9 + 1
        ww = (w - 4 - 1) * r - 6 + 9 + 1
        return (n, c, hh, ww)
    if name == "vdsr":
        return in_shape
    if name in ("edsr", "srgan_g"):
        return (n, c, 4 * h, 4 * w)
    raise KeyError(name)


def pixel_shuffle_law(t, r):
    """out[n,c,h*r+i,w*r+j] = in[n,c*r*r+i*r+j,h,w]  (ATen pixel_shuffle; used by base_networks.py:157) -- index form."""
    n, crr, h, w = t.shape
    c = crr // (r * r)
    return t.reshape(n, c, r, r, h, w).permute(0, 1, 4, 2, 5, 3).reshape(n, c, h * r, w * r)


# --------------------------------------------------------------------------------------------
# mask-conditioned replay (test tool): the derivative of ReLU / PReLU / LeakyReLU is discontinuous in the
# pre-activation, so a 1e-3 comparison of GRADIENTS is only meaningful on a common activation pattern.
# --------------------------------------------------------------------------------------------
class _ForcedAct(nn.Module):
    """Stands in for one nn.ReLU / nn.LeakyReLU / nn.PReLU of an oracle net: y = z where the recorded pattern says
    'positive', slope*z elsewhere -- the pattern comes from the implementation under test, not from sign(z)."""

    def __init__(self, orig, feed):
        super().__init__()
        self.feed = feed
        if isinstance(orig, nn.PReLU):
            self.weight = orig.weight  # the same Parameter object: d(alpha) lands where the tests look for it
            self.kind = "prelu"
        elif isinstance(orig, nn.LeakyReLU):
            self.kind, self.slope = "lrelu", orig.negative_slope
        else:
            self.kind, self.slope = "relu", 0.0

    def forward(self, z):
        m = self.feed.pop(0).to(z.dtype)
        assert m.shape == z.shape, (m.shape, z.shape)
        slope = self.weight if self.kind == "prelu" else self.slope
        return z * m + (z * slope) * (1.0 - m)


def with_forced_activations(net, masks):
    """Swap every ReLU/LeakyReLU/PReLU module of `net` (in place; parameters untouched) for a _ForcedAct that consumes
    `masks` (list of bool tensors in execution order).  Returns net."""
    feed = list(masks)

    def swap(mod):
        for name, child in list(mod.named_children()):
            if isinstance(child, (nn.ReLU, nn.LeakyReLU, nn.PReLU)):
                setattr(mod, name, _ForcedAct(child, feed))
            else:
                swap(child)
    swap(net)
    net._forced_feed = feed
    return net


def count_convs(net):
    return sum(1 for m in net.modules() if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)))


# --------------------------------------------------------------------------------------------
# SRGAN remainder: VGG feature extractor, input normalisation and the adversarial step (srgan.py:84-90, 249-310)
# --------------------------------------------------------------------------------------------
class FeatureExtractor(nn.Module):  # srgan.py:84-90 over torchvision's vgg19 layout (features[:feature_layer + 1])
    CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]

    def __init__(self, feature_layer=8):
        super().__init__()
        layers, cin = [], 3
        for v in self.CFG:  # torchvision.models.vgg.make_layers(cfgs['E'], batch_norm=False)
            if v == "M":
                layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
            else:
                layers += [nn.Conv2d(cin, v, kernel_size=3, padding=1), nn.ReLU(inplace=True)]
                cin = v
        self.features = nn.Sequential(*layers[:feature_layer + 1])

    def forward(self, x):
        return self.features(x)


def build_feature_extractor(seed=2):
    """The reference loads ImageNet weights (srgan.py:144, needs network); tests and benches use torchvision's own default
    initialisation (kaiming_normal_ fan_out / relu, zero bias: torchvision/models/vgg.py _initialize_weights) under a seed."""
    torch.manual_seed(seed)
    fe = FeatureExtractor()
    for m in fe.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            nn.init.constant_(m.bias, 0)
    return fe


VGG_MEAN, VGG_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def norm_vgg(img):
    """utils.norm(img, vgg=True) (utils.py:219-229) on a 4-D batch: per-channel (x - mean) / std."""
    mean = torch.tensor(VGG_MEAN, dtype=img.dtype, device=img.device).view(1, 3, 1, 1)
    std = torch.tensor(VGG_STD, dtype=img.dtype, device=img.device).view(1, 3, 1, 1)
    return (img - mean) / std


def make_srgan_optimizers(G, D, lr=1e-5):
    g_opt = torch.optim.Adam(G.parameters(), lr=lr, betas=(0.9, 0.999))                  # srgan.py:147
    d_opt = torch.optim.SGD(D.parameters(), lr=lr / 100, momentum=0.9, nesterov=True)    # srgan.py:149
    return g_opt, d_opt


def srgan_step(G, D, FE, g_opt, d_opt, lr_img, hr_img, mse=TF.mse_loss, bce=None):
    """One adversarial iteration exactly as written in srgan.py:256-310 (including its redundancies: recon is not detached
    for the D update, VGG features of recon.data carry no gradient to G), with the two fixes modern torch forces: labels are
    shaped like the decision (N,1), and `.data` is `.detach()`.  Returns (D_loss, G_loss) detached."""
    if bce is None:
        bce = lambda y, t: TF.binary_cross_entropy(y, t.reshape(y.shape))  # noqa: E731
    x_, y_ = norm_vgg(hr_img), norm_vgg(lr_img)                             # :256-257
    n = x_.shape[0]
    real_label = torch.ones(n, device=x_.device)
    fake_label = torch.zeros(n, device=x_.device)
    d_opt.zero_grad()                                                      # :272
    D_real_loss = bce(D(x_), real_label)                                   # :275-276
    recon = G(y_)                                                          # :279
    D_fake_loss = bce(D(recon), fake_label)                                # :280-281
    D_loss = D_real_loss + D_fake_loss
    D_loss.backward()                                                      # :286
    d_opt.step()
    g_opt.zero_grad()                                                      # :290
    recon = G(y_)                                                          # :293
    GAN_loss = bce(D(recon), real_label)                                   # :294-297
    mse_loss = mse(recon, x_)                                              # :300
    x_VGG = norm_vgg(hr_img)
    recon_VGG = norm_vgg(recon.detach())                                   # :302 (recon_image.data)
    real_feature = FE(x_VGG)
    fake_feature = FE(recon_VGG)
    vgg_loss = mse(fake_feature, real_feature.detach())                    # :305
    G_loss = mse_loss + 6e-3 * vgg_loss + 1e-3 * GAN_loss                  # :308
    G_loss.backward()
    g_opt.step()
    return D_loss.detach(), G_loss.detach()
