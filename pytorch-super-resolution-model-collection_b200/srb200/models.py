"""Host-side mirrors of the six reference `Net` definitions, built over srb200.base_networks.

In the reference these classes live in the per-model scripts (srcnn.py:13, espcn.py:13, fsrcnn.py:13,
vdsr.py:13, edsr.py:13, srgan.py:14/49) and stay untouched when `base_networks` is swapped for
srb200.base_networks; they are mirrored here only because the GPU box has no /root/reference, so the
benchmark, smoke test and GPU parity tests need the topologies in-tree.  Attribute names (and so the
state_dict keys) and forward dataflow are the reference's; initialisation lives with the reference's
utils.py and is restated only in the test oracle.
"""
import torch
import torch.nn as nn

from .base_networks import ConvBlock, PSBlock, ResnetBlock, Upsample2xBlock, DenseBlock, prepare
from .convert import PReLU, ConvTranspose2d


def _seq(blocks):
    return nn.Sequential(*blocks)


class SRCNN(nn.Module):
    """srcnn.py:13-25  conv9+ReLU -> conv5+ReLU -> conv5, all pad 0."""

    def __init__(self, num_channels=3, base_filter=64):
        super().__init__()
        self.layers = _seq([
            ConvBlock(num_channels, base_filter, 9, 1, 0, norm=None),
            ConvBlock(base_filter, base_filter // 2, 5, 1, 0, norm=None),
            ConvBlock(base_filter // 2, num_channels, 5, 1, 0, activation=None, norm=None)])

    def forward(self, x):
        return self.layers(x)


class ESPCN(nn.Module):
    """espcn.py:13-25  conv5+ReLU -> conv3+ReLU -> conv3 -> PixelShuffle(r) (fused in the last conv's store)."""

    def __init__(self, num_channels=3, base_filter=64, scale_factor=4):
        super().__init__()
        self.layers = _seq([
            ConvBlock(num_channels, base_filter, 5, 1, 0, activation='relu', norm=None),
            ConvBlock(base_filter, base_filter // 2, 3, 1, 0, activation='relu', norm=None),
            PSBlock(base_filter // 2, num_channels, scale_factor, 3, 1, 0, activation=None, norm=None)])

    def forward(self, x):
        return self.layers(x)


class FSRCNN(nn.Module):
    """fsrcnn.py:13-43  conv5+PReLU -> [conv1+PReLU, m x conv3, PReLU, conv1+PReLU] -> deconv9 stride r."""

    def __init__(self, num_channels=3, scale_factor=4, d=56, s=12, m=4):
        super().__init__()
        self.first_part = ConvBlock(num_channels, d, 5, 1, 0, activation='prelu', norm=None)
        mid = [ConvBlock(d, s, 1, 1, 0, activation='prelu', norm=None)]
        for _ in range(m):
            mid.append(ConvBlock(s, s, 3, 1, 1, activation=None, norm=None))
        mid.append(PReLU())
        mid.append(ConvBlock(s, d, 1, 1, 0, activation='prelu', norm=None))
        self.mid_part = _seq(mid)
        self.last_part = ConvTranspose2d(d, num_channels, 9, scale_factor, 3, output_padding=1)

    def forward(self, x):
        return self.last_part(self.mid_part(self.first_part(x)))


class VDSR(nn.Module):
    """vdsr.py:13-32  20 bias-free conv3 (+ReLU) layers and a global residual."""

    def __init__(self, num_channels=3, base_filter=64, num_residuals=18):
        super().__init__()
        self.input_conv = ConvBlock(num_channels, base_filter, 3, 1, 1, norm=None, bias=False)
        self.residual_layers = _seq([ConvBlock(base_filter, base_filter, 3, 1, 1, norm=None, bias=False)
                                     for _ in range(num_residuals)])
        self.output_conv = ConvBlock(base_filter, num_channels, 3, 1, 1, activation=None, norm=None, bias=False)

    def forward(self, x):
        residual = x
        out = self.input_conv(x)
        out = self.residual_layers(out)
        out = self.output_conv(out)
        return torch.add(out, residual)


class EDSR(nn.Module):
    """edsr.py:13-45  conv3 -> N x ResnetBlock -> conv3 -> +skip -> 2 x (conv3 -> PS2) -> conv3."""

    def __init__(self, num_channels=3, base_filter=64, num_residuals=16):
        super().__init__()
        self.input_conv = ConvBlock(num_channels, base_filter, 3, 1, 1, activation=None, norm=None)
        self.residual_layers = _seq([ResnetBlock(base_filter, norm=None) for _ in range(num_residuals)])
        self.mid_conv = ConvBlock(base_filter, base_filter, 3, 1, 1, activation=None, norm=None)
        self.upscale4x = _seq([
            Upsample2xBlock(base_filter, base_filter, upsample='ps', activation=None, norm=None),
            Upsample2xBlock(base_filter, base_filter, upsample='ps', activation=None, norm=None)])
        self.output_conv = ConvBlock(base_filter, num_channels, 3, 1, 1, activation=None, norm=None)

    def forward(self, x):
        out = self.input_conv(x)
        residual = out
        out = self.residual_layers(out)
        out = self.mid_conv(out)
        out = torch.add(out, residual)
        out = self.upscale4x(out)
        return self.output_conv(out)


class SRGANGenerator(nn.Module):
    """srgan.py:14-42  conv9+PReLU -> 16 x ResnetBlock(BN, PReLU) -> conv3+BN -> +skip -> 2 x (conv3->PS2->PReLU) -> conv9."""

    def __init__(self, num_channels=3, base_filter=64, num_residuals=16):
        super().__init__()
        self.input_conv = ConvBlock(num_channels, base_filter, 9, 1, 4, activation='prelu', norm=None)
        self.residual_layers = _seq([ResnetBlock(base_filter, activation='prelu') for _ in range(num_residuals)])
        self.mid_conv = ConvBlock(base_filter, base_filter, 3, 1, 1, activation=None)
        self.upscale4x = _seq([
            Upsample2xBlock(base_filter, base_filter, upsample='ps', activation='prelu', norm=None),
            Upsample2xBlock(base_filter, base_filter, upsample='ps', activation='prelu', norm=None)])
        self.output_conv = ConvBlock(base_filter, num_channels, 9, 1, 4, activation=None, norm=None)

    def forward(self, x):
        out = self.input_conv(x)
        residual = out
        out = self.residual_layers(out)
        out = self.mid_conv(out)
        out = torch.add(out, residual)
        out = self.upscale4x(out)
        return self.output_conv(out)


class SRGANDiscriminator(nn.Module):
    """srgan.py:49-77  8 ConvBlocks (stride 1/2 alternating, LeakyReLU, BN on all but the first) + 2 DenseBlocks."""

    def __init__(self, num_channels=3, base_filter=64, image_size=128):
        super().__init__()
        self.image_size = image_size
        f = base_filter
        self.input_conv = ConvBlock(num_channels, f, 3, 1, 1, activation='lrelu', norm=None)
        self.conv_blocks = _seq([
            ConvBlock(f, f, 3, 2, 1, activation='lrelu'),
            ConvBlock(f, f * 2, 3, 1, 1, activation='lrelu'),
            ConvBlock(f * 2, f * 2, 3, 2, 1, activation='lrelu'),
            ConvBlock(f * 2, f * 4, 3, 1, 1, activation='lrelu'),
            ConvBlock(f * 4, f * 4, 3, 2, 1, activation='lrelu'),
            ConvBlock(f * 4, f * 8, 3, 1, 1, activation='lrelu'),
            ConvBlock(f * 8, f * 8, 3, 2, 1, activation='lrelu')])
        self.dense_layers = _seq([
            DenseBlock(f * 8 * image_size // 16 * image_size // 16, f * 16, activation='lrelu', norm=None),
            DenseBlock(f * 16, 1, activation='sigmoid', norm=None)])
        prepare(self)  # the last conv block returns NCHW so the reference's `.view` flatten (srgan.py:75) works unchanged

    def forward(self, x):
        out = self.input_conv(x)
        out = self.conv_blocks(out)
        out = out.view(out.size()[0], -1)  # srgan.py:75
        return self.dense_layers(out)


class MaxPool2d(nn.MaxPool2d):
    """nn.MaxPool2d(2, 2) on libsrb200 (class name unchanged: the reference never matches on it)."""

    def forward(self, x):
        from . import nn_ops
        if self.kernel_size not in (2, (2, 2)) or self.stride not in (2, (2, 2)) or self.padding not in (0, (0, 0)):
            raise RuntimeError("srb200.MaxPool2d implements MaxPool2d(2, 2) (VGG19 features[4])")
        return nn_ops.max_pool2(x)


class FeatureExtractor(nn.Module):
    """srgan.py:84-90  vgg19.features[:feature_layer + 1] (conv3-64, ReLU, conv3-64, ReLU, maxpool, conv3-128, ReLU,
    conv3-128, ReLU for the default 8).  Child indices -- and so the state_dict keys features.{0,2,5,7}.* of a torchvision
    checkpoint -- are kept; every Conv2d + ReLU pair runs as one fused libsrb200 kernel, the pool as srb_maxpool2.
    `netVGG`: a torchvision vgg19 (its layers are shared, not copied) or None for freshly initialised layers (the reference
    downloads ImageNet weights, srgan.py:144; no network here)."""
    CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]

    def __init__(self, netVGG=None, feature_layer=8):
        super().__init__()
        if netVGG is not None:
            layers = list(netVGG.features.children())[:feature_layer + 1]
        else:
            layers, cin = [], 3
            for v in self.CFG:
                if v == "M":
                    layers.append(nn.MaxPool2d(2, 2))
                else:
                    layers += [nn.Conv2d(cin, v, 3, padding=1), nn.ReLU(True)]
                    cin = v
            layers = layers[:feature_layer + 1]
        self.features = nn.Sequential(*layers)

    def forward(self, x):
        from . import functional as F
        from . import nn_ops
        mods = list(self.features.children())
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Conv2d):
                relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
                x = F.conv2d(x, m.weight, m.bias, m.stride[0], m.padding[0], activation="relu" if relu else None)
                i += 2 if relu else 1
            elif isinstance(m, nn.MaxPool2d):
                x = nn_ops.max_pool2(x)
                i += 1
            elif isinstance(m, nn.ReLU):
                x = torch.relu(x)
                i += 1
            else:
                x = m(x)
                i += 1
        return x


MODELS = {"srcnn": SRCNN, "espcn": ESPCN, "fsrcnn": FSRCNN, "vdsr": VDSR, "edsr": EDSR,
          "srgan_g": SRGANGenerator, "srgan_d": SRGANDiscriminator, "srgan_fe": FeatureExtractor}
