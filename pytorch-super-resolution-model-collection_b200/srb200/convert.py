"""convert(net): rewrite an already-built reference model in place so its base_networks blocks and raw
nn.PReLU / nn.ConvTranspose2d / nn.Conv2d layers (fsrcnn.py:26,33) run on libsrb200.  Parameters are
shared, not copied: state_dict keys, optimizers and checkpoints are unchanged."""
import torch

from . import base_networks as B
from . import functional as F

_BLOCKS = {"ConvBlock": B.ConvBlock, "DeconvBlock": B.DeconvBlock, "ResnetBlock": B.ResnetBlock,
           "PSBlock": B.PSBlock, "Upsample2xBlock": B.Upsample2xBlock, "DenseBlock": B.DenseBlock}


class PReLU(torch.nn.PReLU):
    """nn.PReLU() with one shared slope, forward/backward on libsrb200 (class name keeps 'PReLU')."""

    def forward(self, x):
        if self.weight.numel() != 1:
            raise RuntimeError("srb200.PReLU supports the single-slope nn.PReLU() the reference uses")
        return F.prelu(x, self.weight)


class ConvTranspose2d(torch.nn.ConvTranspose2d):
    def forward(self, x, output_size=None):
        if output_size is not None or self.groups != 1 or self.dilation != (1, 1):
            raise RuntimeError("srb200.ConvTranspose2d: output_size/groups/dilation are not used by the reference")
        return F.conv_transpose2d(x, self.weight, self.bias, self.stride[0], self.padding[0], self.output_padding[0])


class Conv2d(torch.nn.Conv2d):
    def forward(self, x):
        if self.groups != 1 or self.dilation != (1, 1) or self.padding_mode != "zeros":
            raise RuntimeError("srb200.Conv2d: groups/dilation/padding_mode are not used by the reference")
        return F.conv2d(x, self.weight, self.bias, self.stride[0], self.padding[0])


def _is_ours(m):
    return type(m).__module__ == B.__name__


def _swap_block(m):
    cls = _BLOCKS[type(m).__name__]
    new = cls.__new__(cls)
    torch.nn.Module.__init__(new)
    # share every child / parameter / buffer / plain attribute (norm, activation)
    new.__dict__.update({k: v for k, v in m.__dict__.items() if not k.startswith("_")})
    for k in ("_parameters", "_buffers", "_modules"):
        new.__dict__[k] = m.__dict__[k]
    new.training = m.training
    return new


def convert(net):
    """In-place conversion; returns net.  Ends with base_networks.prepare(net) (NCHW output in front of a `.view` flatten)."""
    _convert(net)
    return B.prepare(net)


def _convert(net):
    for name, child in list(net.named_children()):
        tname = type(child).__name__
        if tname == "FeatureExtractor" and type(child).__module__ != "srb200.models":
            from .models import FeatureExtractor  # srgan.py:84-90: same `features` Sequential, fused forward
            new = FeatureExtractor.__new__(FeatureExtractor)
            torch.nn.Module.__init__(new)
            new.features = child.features
            setattr(net, name, new)
            continue
        if tname in _BLOCKS and not _is_ours(child):
            new = _swap_block(child)
            _convert(new)  # Upsample2xBlock holds nested blocks
            setattr(net, name, new)
        elif type(child) is torch.nn.PReLU:
            child.__class__ = PReLU
        elif type(child) is torch.nn.ConvTranspose2d:
            child.__class__ = ConvTranspose2d
        elif type(child) is torch.nn.Conv2d and type(net).__name__ not in _BLOCKS:
            child.__class__ = Conv2d
        else:
            _convert(child)
    return net
