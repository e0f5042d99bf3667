"""Drop-in for the reference's base_networks.py: same class names, constructor signatures, defaults,
child-module names (conv / deconv / conv1 / conv2 / ps / bn / act / fc / upsample -> identical state_dict
keys) and forward semantics, with the Conv2d / ConvTranspose2d / PixelShuffle / ReLU /
PReLU / LeakyReLU / residual-add stacks executed by libsrb200's fused sm_100a kernels instead of ATen.

Usage (reference model files untouched; they do `from base_networks import *`, e.g. srcnn.py:5):
    import sys, srb200.base_networks as bn; sys.modules['base_networks'] = bn
or rewrite an already-built model with srb200.convert(net).

The parameter holders stay torch.nn.Conv2d / ConvTranspose2d / PReLU modules so that
utils.weights_init_* (class-name matching, utils.py:76-113), optimizers and checkpoints keep working;
only their forward is bypassed.  BatchNorm / InstanceNorm / Linear / tanh / sigmoid stay on torch
(SURVEY.md 8f "next" rows).
"""
import os

import torch  # re-exported on purpose: the model files get `torch` through the star import (srcnn.py:13)

from . import functional as F
from . import nn_ops as N

__all__ = ["torch", "DenseBlock", "ConvBlock", "DeconvBlock", "ResnetBlock", "PSBlock", "Upsample2xBlock"]  # (prepare() is not star-exported: the model files never call it)

_FUSABLE = (None, "relu", "prelu", "lrelu")


def _make_norm(norm_arg, self_norm, ctor2d, num):
    # mirrors the reference's quirk: ResnetBlock/PSBlock test the ctor argument for 'instance' (base_networks.py:118,162)
    if self_norm == "batch":
        return torch.nn.BatchNorm2d(num)
    if norm_arg == "instance":
        return torch.nn.InstanceNorm2d(num)
    return None


def _make_act(activation):
    if activation == "relu":
        return torch.nn.ReLU(True)
    if activation == "prelu":
        return torch.nn.PReLU()
    if activation == "lrelu":
        return torch.nn.LeakyReLU(0.2, True)
    if activation == "tanh":
        return torch.nn.Tanh()
    if activation == "sigmoid":
        return torch.nn.Sigmoid()
    return None


class _ActMixin:
    # Set by srb200.prepare()/convert() on the last conv block in front of a flatten (srgan.py:75 does
    # `out.view(N, -1)`, which needs NCHW-contiguous memory): the block then returns NCHW instead of channels_last.
    nchw_out = False
    # Set by srb200.FusedLoss for one forward pass on the network's last block: (target, kind, holder).  The block then runs
    # the conv with the loss fused into its epilogue (F.conv2d_loss) and leaves the loss in holder["loss"].
    _loss_req = None

    def _loss_forward(self, x, c, r):
        target, kind, holder = self._loss_req
        if not (x.is_cuda and F.loss_fusable(tuple(x.shape), F._is_cl(x), tuple(c.weight.shape), c.stride[0], c.padding[0], r)):
            return None
        loss, y = F.conv2d_loss(x, c.weight, c.bias, target, kind, c.stride[0], c.padding[0], r, need_output=True)
        holder["loss"], holder["y"] = loss, y
        return y

    def _layout(self, out):
        return out.contiguous() if self.nchw_out else out

    def _fused_act(self):
        """(activation code, alpha) if the block's activation can be fused into the conv epilogue."""
        if self.activation in _FUSABLE:
            return self.activation, (self.act.weight if self.activation == "prelu" else None)
        return None, None

    def _norm_act(self, out, residual=None, with_act=True):
        """norm -> activation (-> + residual) after a conv.  BatchNorm2d on CUDA runs as libsrb200's fused kernels
        (statistics, normalise + activation + residual in one pass; backward likewise); InstanceNorm and tanh / sigmoid
        stay on torch."""
        act = self.activation if with_act else None
        bn = self.bn
        if isinstance(bn, torch.nn.BatchNorm2d) and out.is_cuda and out.dtype == torch.float32 and out.shape[1] % 4 == 0:
            fus = act if act in _FUSABLE else None
            out = N.batch_norm_act(out, bn, activation=fus, alpha=self.act.weight if fus == "prelu" else None, residual=residual)
            if act is not None and fus is None:
                out = self.act(out)
            return out
        out = bn(out)
        if act is not None:
            out = self._post(out, fused=False)
        return out if residual is None else torch.add(out, residual)

    def _post(self, out, fused):
        # activation the conv kernel could not fuse (after a norm layer, or tanh/sigmoid)
        if self.activation is not None and not fused:
            if self.activation == "prelu":
                return F.prelu(out, self.act.weight)
            out = self.act(out)
            if self.activation in ("relu", "lrelu"):
                F._record(out)
            return out
        return out


class DenseBlock(torch.nn.Module):
    """base_networks.py:4-36 -- Linear (+BN1d) (+act); the Linear runs on libsrb200 (nn_ops.linear), BatchNorm1d on torch."""

    def __init__(self, input_size, output_size, bias=True, activation='relu', norm='batch'):
        super(DenseBlock, self).__init__()
        self.fc = torch.nn.Linear(input_size, output_size, bias=bias)
        self.norm = norm
        if self.norm == 'batch':
            self.bn = torch.nn.BatchNorm1d(output_size)
        elif self.norm == 'instance':
            self.bn = torch.nn.InstanceNorm1d(output_size)
        self.activation = activation
        act = _make_act(activation)
        if act is not None:
            self.act = act

    def forward(self, x):
        if x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.shape[1] % 4 == 0:
            out = N.linear(x, self.fc.weight, self.fc.bias)  # weights streamed once by libsrb200 (HBM-bound GEMV-like layer)
        else:
            out = self.fc(x)
        if self.norm is not None:
            out = self.bn(out)
        if self.activation is None:
            return out
        out = self.act(out)
        if self.activation in ("relu", "lrelu", "prelu"):
            F._record(out)
        return out


class ConvBlock(torch.nn.Module, _ActMixin):
    """base_networks.py:39-71 -- act(bn?(Conv2d(x))).  norm=None: one fused kernel."""

    def __init__(self, input_size, output_size, kernel_size=4, stride=2, padding=1, bias=True, activation='relu',
                 norm='batch'):
        super(ConvBlock, self).__init__()
        self.conv = torch.nn.Conv2d(input_size, output_size, kernel_size, stride, padding, bias=bias)
        self.norm = norm
        bn = _make_norm(norm, self.norm, None, output_size) if norm in ('batch', 'instance') else None
        if bn is not None:
            self.bn = bn
        self.activation = activation
        act = _make_act(activation)
        if act is not None:
            self.act = act

    def forward(self, x):
        c = self.conv
        if self._loss_req is not None and self.norm is None and self.activation is None:
            fused = self._loss_forward(x, c, 1)
            if fused is not None:
                return fused
        if self.norm is not None:
            return self._layout(self._norm_act(F.conv2d(x, c.weight, c.bias, c.stride[0], c.padding[0])))
        a, alpha = self._fused_act()
        out = F.conv2d(x, c.weight, c.bias, c.stride[0], c.padding[0], activation=a, alpha=alpha)
        return self._layout(self._post(out, fused=self.activation in _FUSABLE))


class DeconvBlock(torch.nn.Module, _ActMixin):
    """base_networks.py:74-106 -- act(bn?(ConvTranspose2d(x)))."""

    def __init__(self, input_size, output_size, kernel_size=4, stride=2, padding=1, bias=True, activation='relu',
                 norm='batch'):
        super(DeconvBlock, self).__init__()
        self.deconv = torch.nn.ConvTranspose2d(input_size, output_size, kernel_size, stride, padding, bias=bias)
        self.norm = norm
        bn = _make_norm(norm, self.norm, None, output_size) if norm in ('batch', 'instance') else None
        if bn is not None:
            self.bn = bn
        self.activation = activation
        act = _make_act(activation)
        if act is not None:
            self.act = act

    def forward(self, x):
        d = self.deconv
        if self.norm is not None:
            return self._layout(self._norm_act(F.conv_transpose2d(x, d.weight, d.bias, d.stride[0], d.padding[0], d.output_padding[0])))
        a, alpha = self._fused_act()
        out = F.conv_transpose2d(x, d.weight, d.bias, d.stride[0], d.padding[0], d.output_padding[0],
                                 activation=a, alpha=alpha)
        return self._layout(self._post(out, fused=self.activation in _FUSABLE))


class ResnetBlock(torch.nn.Module, _ActMixin):
    """base_networks.py:109-150 -- x + bn?(conv2(act(bn?(conv1(x))))); ONE bn module shared by both convs (:137,145).
    norm=None: two fused kernels (conv1+act, conv2+residual add)."""

    def __init__(self, num_filter, kernel_size=3, stride=1, padding=1, bias=True, activation='relu', norm='batch'):
        super(ResnetBlock, self).__init__()
        self.conv1 = torch.nn.Conv2d(num_filter, num_filter, kernel_size, stride, padding, bias=bias)
        self.conv2 = torch.nn.Conv2d(num_filter, num_filter, kernel_size, stride, padding, bias=bias)
        self.norm = norm
        if self.norm == 'batch':
            self.bn = torch.nn.BatchNorm2d(num_filter)
        elif norm == 'instance':
            self.bn = torch.nn.InstanceNorm2d(num_filter)
        self.activation = activation
        act = _make_act(activation)
        if act is not None:
            self.act = act

    def forward(self, x):
        c1, c2 = self.conv1, self.conv2
        if self.norm is not None:
            # conv1 -> bn -> act -> conv2 -> bn (the SAME module, :137,145) -> + x; the add rides in the second norm kernel
            out = self._norm_act(F.conv2d(x, c1.weight, c1.bias, c1.stride[0], c1.padding[0]))
            out = self._norm_act(F.conv2d(out, c2.weight, c2.bias, c2.stride[0], c2.padding[0]), residual=x, with_act=False)
            return self._layout(out)
        a, alpha = self._fused_act()
        # x feeds conv1 and the skip: its two gradients are summed in conv1's dgrad epilogue (F.SkipGrad), not by an add kernel
        tok = F.SkipGrad() if (_SKIPGRAD and x.requires_grad and torch.is_grad_enabled() and x.is_cuda) else None
        out = F.conv2d(x, c1.weight, c1.bias, c1.stride[0], c1.padding[0], activation=a, alpha=alpha, skip_sink=tok)
        out = self._post(out, fused=self.activation in _FUSABLE)
        return self._layout(F.conv2d(out, c2.weight, c2.bias, c2.stride[0], c2.padding[0], residual=x, skip_src=tok))


class PSBlock(torch.nn.Module, _ActMixin):
    """base_networks.py:153-185 -- act?(bn?(PixelShuffle_r(Conv2d(in -> out*r*r)))).
    The shuffle is never a kernel: it is the store addressing of the conv epilogue (bit-exact permutation)."""

    def __init__(self, input_size, output_size, scale_factor, kernel_size=3, stride=1, padding=1, bias=True,
                 activation='relu', norm='batch'):
        super(PSBlock, self).__init__()
        self.conv = torch.nn.Conv2d(input_size, output_size * scale_factor ** 2, kernel_size, stride, padding,
                                    bias=bias)
        self.ps = torch.nn.PixelShuffle(scale_factor)
        self.norm = norm
        if self.norm == 'batch':
            self.bn = torch.nn.BatchNorm2d(output_size)
        elif norm == 'instance':
            self.bn = torch.nn.InstanceNorm2d(output_size)
        self.activation = activation
        act = _make_act(activation)
        if act is not None:
            self.act = act

    def forward(self, x):
        c = self.conv
        r = self.ps.upscale_factor
        if self._loss_req is not None and self.norm is None and self.activation is None:
            fused = self._loss_forward(x, c, r)
            if fused is not None:
                return fused
        if self.norm is not None:
            return self._layout(self._norm_act(F.conv2d(x, c.weight, c.bias, c.stride[0], c.padding[0], pixel_shuffle=r)))
        a, alpha = self._fused_act()
        out = F.conv2d(x, c.weight, c.bias, c.stride[0], c.padding[0], activation=a, alpha=alpha, pixel_shuffle=r)
        return self._layout(self._post(out, fused=self.activation in _FUSABLE))


class Upsample2xBlock(torch.nn.Module):
    """base_networks.py:188-214 -- 'deconv' (k4 s2 p1) | 'ps' (r=2) | 'rnc' (nearest x2 + conv3)."""

    def __init__(self, input_size, output_size, bias=True, upsample='deconv', activation='relu', norm='batch'):
        super(Upsample2xBlock, self).__init__()
        scale_factor = 2
        if upsample == 'deconv':
            self.upsample = DeconvBlock(input_size, output_size, kernel_size=4, stride=2, padding=1, bias=bias,
                                        activation=activation, norm=norm)
        elif upsample == 'ps':
            self.upsample = PSBlock(input_size, output_size, scale_factor=scale_factor, bias=bias,
                                    activation=activation, norm=norm)
        elif upsample == 'rnc':
            self.upsample = torch.nn.Sequential(
                torch.nn.Upsample(scale_factor=scale_factor, mode='nearest'),
                ConvBlock(input_size, output_size, kernel_size=3, stride=1, padding=1, bias=bias,
                          activation=activation, norm=norm))

    def forward(self, x):
        return self.upsample(x)


_CONV_BLOCK_NAMES = ("ConvBlock", "DeconvBlock", "ResnetBlock", "PSBlock")


def prepare(net):
    """Layout fix-up for models that flatten a conv activation with `.view` (srgan.Discriminator, srgan.py:75): our
    blocks hand channels_last memory to each other, which `.view(N, -1)` rejects (and whose element order would not
    match the Linear weights anyway).  The last conv block registered before the first DenseBlock is told to return
    NCHW-contiguous memory.  Called by srb200.convert(); call it yourself after building a model through the
    `sys.modules['base_networks']` swap.  Returns net."""
    last_conv = None
    for m in net.modules():
        name = type(m).__name__
        if name == "DenseBlock":
            if last_conv is not None:
                last_conv.nchw_out = True
            break
        if name in _CONV_BLOCK_NAMES and type(m).__module__ == __name__:
            last_conv = m
    return net


# SRB_NO_SKIPGRAD=1: leave the residual blocks' fan-out sum to autograd (A/B measurements)
_SKIPGRAD = os.environ.get("SRB_NO_SKIPGRAD") != "1"


class FusedLoss(torch.nn.Module):
    """`loss = FusedLoss(net, 'mse' | 'l1')(x, target)`  ==  `criterion(net(x), target)` of the reference loops
    (espcn.py:128-129, srcnn.py:128-129, edsr.py:152-153) with the criterion evaluated inside the epilogue of the
    network's last convolution: no loss-forward, loss-backward or pixel-un-shuffle kernels, and dL/dy never exists in
    the shuffled layout.  Applies when the network's output IS the output of its last ConvBlock / PSBlock (no norm, no
    activation) -- checked on the first call; otherwise (VDSR adds its global residual after the last conv, FSRCNN ends
    in a transposed conv) the plain fused loss kernels are used."""

    def __init__(self, net, kind="mse"):
        super().__init__()
        self.net = net
        self.kind = {"l2": "mse"}.get(kind, kind)
        self.last = None
        for m in net.modules():
            if type(m).__name__ in ("ConvBlock", "PSBlock") and type(m).__module__ == __name__:
                self.last = m
        self.mode = None  # decided on the first call: "fused" | "plain"

    def _plain(self, x, target):
        y = self.net(x)
        return (F.l1_loss if self.kind == "l1" else F.mse_loss)(y, target)

    def forward(self, x, target):
        """target: the fp32 (N,C,H,W) tensor of the reference loops, or the decoded uint8 (N,H,W,C) image batch itself -- the
        fused epilogue then applies ToTensor's byte/255 on the fly (dataset.py:90) and the fp32 HR tensor never exists."""
        u8 = target.dtype == torch.uint8
        if self.mode == "plain" or self.last is None:
            return self._plain(x, F.image_to_tensor(target) if u8 else target)
        holder = {}
        self.last._loss_req = (target, self.kind, holder)
        try:
            y = self.net(x)
        finally:
            self.last._loss_req = None
        if "loss" in holder and y is holder["y"]:
            self.mode = "fused"
            return holder["loss"]
        if u8:
            target = F.image_to_tensor(target)
        if "loss" in holder:
            # the network post-processes its last block's output: the fused loss would be wrong -- never use it here again
            self.mode = "plain"
            return (F.l1_loss if self.kind == "l1" else F.mse_loss)(y, target) if y.requires_grad else self._plain(x, target)
        self.mode = "plain"
        return (F.l1_loss if self.kind == "l1" else F.mse_loss)(y, target)
