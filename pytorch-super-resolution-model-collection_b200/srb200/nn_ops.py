"""autograd.Functions over libsrb200 for the layers around the conv stacks of SRGAN:
train-mode BatchNorm2d fused with the activation / residual add that follows it (base_networks.py:46,117,137,145,161),
nn.Linear (DenseBlock, base_networks.py:7,29-31), MaxPool2d(2) (VGG19 features[4], srgan.py:84-90), BCELoss (srgan.py:157).
torch allocates tensors and supplies the stream; the arithmetic is the library's."""
import ctypes

import torch

from . import _lib
from ._lib import lib, check
from .functional import _state, _workspace, _stream, _ptr, _is_cl, _record

_ACT = {None: _lib.ACT_NONE, "relu": _lib.ACT_RELU, "prelu": _lib.ACT_PRELU, "lrelu": _lib.ACT_LRELU}


def _rounds():
    return 1 if _state["math"] in (_lib.MATH_AUTO, _lib.MATH_TF32) else 0


class _BatchNormAct(torch.autograd.Function):
    """y = act(BatchNorm2d(x)) + residual over a channels_last fp32 tensor."""

    @staticmethod
    def forward(ctx, x, gamma, beta, alpha, residual, running_mean, running_var, training, momentum, eps, act, slope):
        if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4):
            raise RuntimeError("srb200 batch norm needs a CUDA float32 NCHW-shaped tensor; there is no CPU path")
        N, C, H, W = x.shape
        if C % 4 != 0:
            raise RuntimeError("srb200 batch norm needs C % 4 == 0")
        if not _is_cl(x):
            x = x.contiguous(memory_format=torch.channels_last)
        if residual is not None and not _is_cl(residual):
            residual = residual.contiguous(memory_format=torch.channels_last)
        y = torch.empty_like(x)
        mean = torch.empty(C, dtype=torch.float32, device=x.device)
        invstd = torch.empty(C, dtype=torch.float32, device=x.device)
        ws = _workspace(x.device, int(lib.srb_bn_workspace_bytes(C)))
        rnd = _rounds() if C >= 8 else 0
        check(lib.srb_bn_fwd(_ptr(x), _ptr(y), N * H * W, C, _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var),
                             1 if training else 0, ctypes.c_float(momentum), ctypes.c_float(eps), _ptr(mean), _ptr(invstd),
                             _ACT[act], ctypes.c_float(slope), _ptr(alpha), _ptr(residual), rnd, _ptr(ws), ws.numel(),
                             _stream(x.device)))
        ctx.act, ctx.slope, ctx.rnd = act, slope, rnd
        ctx.has_res = residual is not None
        ctx.params = (gamma, beta)
        ctx.save_for_backward(x, gamma, beta, alpha, mean, invstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, alpha, mean, invstd = ctx.saved_tensors
        N, C, H, W = x.shape
        if not _is_cl(dy):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x)
        gparam, bparam = ctx.params
        direct = getattr(gparam, "_srb_direct", False) and gparam.grad is not None and \
            getattr(bparam, "_srb_direct", False) and bparam.grad is not None
        accumulate = 0
        if direct:
            dg_t, db_t, scale = gparam.grad, bparam.grad, _state["grad_scale"]
            accumulate = 1 if getattr(gparam, "_srb_written", False) else 0
            gparam._srb_written = True
            bparam._srb_written = True
            dgamma = dbeta = None
        else:
            dgamma = dg_t = torch.empty_like(gamma)
            dbeta = db_t = torch.empty_like(beta)
            scale = 1.0
        dalpha = torch.zeros_like(alpha) if (ctx.act == "prelu" and alpha is not None) else None
        ws = _workspace(x.device, int(lib.srb_bn_workspace_bytes(C)))
        check(lib.srb_bn_bwd(_ptr(x), _ptr(dy), _ptr(dx), N * H * W, C, _ptr(gamma), _ptr(beta), _ptr(mean), _ptr(invstd),
                             _ACT[ctx.act], ctypes.c_float(ctx.slope), _ptr(alpha), _ptr(dg_t), _ptr(db_t), _ptr(dalpha),
                             ctypes.c_float(scale), accumulate, ctx.rnd, _ptr(ws), ws.numel(), _stream(x.device)))
        dres = dy if (ctx.has_res and ctx.needs_input_grad[4]) else None
        return dx, dgamma, dbeta, dalpha, dres, None, None, None, None, None, None, None


def batch_norm_act(x, bn, activation=None, alpha=None, slope=0.2, residual=None):
    """act(bn(x)) + residual with `bn` an nn.BatchNorm2d module (its parameters, buffers, momentum, eps and training flag
    are used and its running statistics / num_batches_tracked updated exactly like the module's own forward)."""
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    momentum = 0.1 if bn.momentum is None else bn.momentum
    training = bn.training or not bn.track_running_stats
    y = _BatchNormAct.apply(x, bn.weight, bn.bias, alpha if activation == "prelu" else None, residual,
                            bn.running_mean if bn.track_running_stats else None,
                            bn.running_var if bn.track_running_stats else None, training, float(momentum), float(bn.eps),
                            activation, float(slope))
    if activation is not None and _state["act_recorder"] is not None:
        _record(y, residual)
    return y


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2):
            raise RuntimeError("srb200 linear needs a CUDA float32 (B, I) tensor; there is no CPU path")
        x = x.contiguous()
        weight = weight.contiguous()
        B, I = x.shape
        O = weight.shape[0]
        y = torch.empty((B, O), dtype=torch.float32, device=x.device)
        ws = _workspace(x.device, int(lib.srb_linear_workspace_bytes(I, O)))
        check(lib.srb_linear_fwd(_ptr(x), _ptr(weight), _ptr(bias), _ptr(y), B, I, O, _ptr(ws), ws.numel(), _stream(x.device)))
        ctx.params = (weight, bias)
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        B, I = x.shape
        O = weight.shape[0]
        dev = x.device
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        wparam, bparam = ctx.params
        dw = db = None
        dw_t = db_t = None
        scale, accumulate = 1.0, 0
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            direct = getattr(wparam, "_srb_direct", False) and wparam.grad is not None and \
                (bparam is None or (getattr(bparam, "_srb_direct", False) and bparam.grad is not None))
            if direct:
                dw_t, db_t, scale = wparam.grad, (bparam.grad if bparam is not None else None), _state["grad_scale"]
                accumulate = 1 if getattr(wparam, "_srb_written", False) else 0
                wparam._srb_written = True
                if bparam is not None:
                    bparam._srb_written = True
            else:
                dw = dw_t = torch.empty_like(weight)
                db = db_t = torch.empty(O, dtype=torch.float32, device=dev) if ctx.has_bias else None
        ws = _workspace(dev, int(lib.srb_linear_workspace_bytes(I, O)))
        check(lib.srb_linear_bwd(_ptr(x), _ptr(weight), _ptr(dy), _ptr(dx), _ptr(dw_t), _ptr(db_t), B, I, O, ctypes.c_float(scale),
                                 accumulate, _ptr(ws), ws.numel(), _stream(dev)))
        return dx, dw, db


def linear(x, weight, bias=None):
    """nn.Linear forward / backward on libsrb200 (weights streamed once per 16 batch rows)."""
    return _Linear.apply(x, weight, bias)


class _MaxPool2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4):
            raise RuntimeError("srb200 max_pool2 needs a CUDA float32 4-D tensor; there is no CPU path")
        if not _is_cl(x):
            x = x.contiguous(memory_format=torch.channels_last)
        N, C, H, W = x.shape
        y = torch.empty((N, C, H // 2, W // 2), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
        idx = torch.empty((N, H // 2, W // 2, C), dtype=torch.uint8, device=x.device)
        check(lib.srb_maxpool2_fwd(_ptr(x), _ptr(y), _ptr(idx), N, C, H, W, _stream(x.device)))
        ctx.shape = (N, C, H, W)
        ctx.save_for_backward(idx)
        return y

    @staticmethod
    def backward(ctx, dy):
        idx, = ctx.saved_tensors
        N, C, H, W = ctx.shape
        if H % 2 or W % 2:
            raise RuntimeError("srb200 max_pool2 backward needs even H and W")
        if not _is_cl(dy):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty((N, C, H, W), dtype=torch.float32, device=dy.device, memory_format=torch.channels_last)
        check(lib.srb_maxpool2_bwd(_ptr(dy), _ptr(idx), _ptr(dx), N, C, H, W, _stream(dy.device)))
        return dx


def max_pool2(x):
    """nn.MaxPool2d(kernel_size=2, stride=2) on a channels_last tensor (VGG19 features[4])."""
    return _MaxPool2.apply(x)


class _BCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, t):
        if not (y.is_cuda and y.dtype == torch.float32):
            raise RuntimeError("srb200 bce_loss needs CUDA float32 tensors; there is no CPU path")
        y = y.contiguous()
        t = t.to(torch.float32).reshape(y.shape).contiguous()
        loss = torch.empty((), dtype=torch.float32, device=y.device)
        check(lib.srb_bce_fwd(_ptr(y), _ptr(t), y.numel(), _ptr(loss), _stream(y.device)))
        ctx.save_for_backward(y, t)
        return loss

    @staticmethod
    def backward(ctx, g):
        y, t = ctx.saved_tensors
        dy = torch.empty_like(y)
        check(lib.srb_bce_bwd(_ptr(y), _ptr(t), y.numel(), _ptr(g.contiguous()), _ptr(dy), _stream(y.device)))
        return dy, None


def bce_loss(y, t):
    """nn.BCELoss() (mean).  The reference compares (N,1) decisions with (N,) labels (srgan.py:264-276), which modern torch
    rejects; the label is reshaped to the decision's shape here."""
    return _BCE.apply(y, t)
