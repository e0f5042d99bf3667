"""Host-side pieces of the reference drivers that surround the hot path and that the in-tree benchmark /
smoke test need (the GPU box has no /root/reference): weight init (utils.py:76-113, fsrcnn.py:45-55),
optimizer and loss choice per model (srcnn.py:79, espcn.py:79, fsrcnn.py:106, vdsr.py:89-90,149, edsr.py:93,98).
Plain torch host code, unchanged in spirit -- none of it is on the accelerated path."""
import torch
import torch.nn as nn
import torch.nn.functional as TF


def weights_init_normal(m, mean=0.0, std=0.02):
    cname = type(m).__name__
    if any(k in cname for k in ("Linear", "Conv2d", "ConvTranspose2d")):
        m.weight.data.normal_(mean, std)
        if m.bias is not None:
            m.bias.data.zero_()
    elif "Norm" in cname:
        m.weight.data.normal_(1.0, 0.02)
        if m.bias is not None:
            m.bias.data.zero_()


def weights_init_kaiming(m):
    cname = type(m).__name__
    if any(k in cname for k in ("Linear", "Conv2d", "ConvTranspose2d")):
        nn.init.kaiming_normal_(m.weight)
        if m.bias is not None:
            m.bias.data.zero_()
    elif "Norm" in cname:
        m.weight.data.normal_(1.0, 0.02)
        if m.bias is not None:
            m.bias.data.zero_()


def init_model(name, net):
    """The weight_init() each reference Net defines."""
    for m in net.modules():
        if name == "srcnn":
            weights_init_normal(m, 0.0, 0.001)
        elif name == "vdsr":
            weights_init_kaiming(m)
        elif name == "fsrcnn":
            if isinstance(m, nn.Conv2d):
                m.weight.data.normal_(0.0, 0.02)
                if m.bias is not None:
                    m.bias.data.zero_()
            if isinstance(m, nn.ConvTranspose2d):
                m.weight.data.normal_(0.0, 0.0001)
                if m.bias is not None:
                    m.bias.data.zero_()
        else:
            weights_init_normal(m)
    return net


def make_optimizer(name, params, lr=1e-5, capturable=False):
    """Same optimizers and hyper-parameters as the reference drivers (srcnn.py:79, espcn.py:79, fsrcnn.py:106,
    vdsr.py:89-90, edsr.py:93).  On CUDA parameters torch's single-kernel multi-tensor implementation is selected
    (`fused=True`): identical update rule, one launch instead of ~7 per step -- it matters for the tiny nets."""
    params = list(params)
    fused = {"fused": True} if (params and params[0].is_cuda) else {}
    adam = dict(fused, capturable=True) if (capturable and fused) else fused  # step counter on the device: CUDA-graph safe
    if name == "srcnn":
        return torch.optim.SGD(params, lr=lr, **fused)
    if name == "espcn":
        return torch.optim.Adam(params, lr=lr, **adam)
    if name == "fsrcnn":
        return torch.optim.SGD(params, lr=lr, momentum=0.9, **fused)
    if name == "vdsr":
        return torch.optim.SGD(params, lr=lr, momentum=0.9, weight_decay=1e-4, **fused)
    return torch.optim.Adam(params, lr=lr, betas=(0.9, 0.999), eps=1e-8, **adam)


def loss_for(name, fused=False):
    """nn.L1Loss for EDSR (edsr.py:98), nn.MSELoss for the others; fused=True selects libsrb200's one-pass kernels."""
    if fused:
        from . import functional as F
        return F.l1_loss if name == "edsr" else F.mse_loss
    return TF.l1_loss if name == "edsr" else TF.mse_loss


VDSR_CLIP = 0.4
