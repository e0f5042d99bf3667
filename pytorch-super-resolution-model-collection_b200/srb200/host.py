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
        elif name == "srgan":
            weights_init_normal(m, 0.0, 0.02)  # srgan.py:44-46,78-80
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


# ---- SRGAN (srgan.py:136-157, 249-310): optimizers, input normalisation and the adversarial step, restated because the
# reference driver no longer runs on torch 2.x; every layer inside G, D and the VGG feature extractor runs on libsrb200
VGG_MEAN, VGG_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


_vgg_consts = {}


def norm_vgg(img):
    """utils.norm(img, vgg=True) (utils.py:219-229) on a 4-D batch (host-side pre-processing in the reference).  The two
    constant tensors are built once per (device, dtype): a host->device copy is not allowed inside a CUDA-graph capture."""
    key = (img.device, img.dtype)
    if key not in _vgg_consts:
        if img.is_cuda and torch.cuda.is_current_stream_capturing():
            raise RuntimeError("srb200.host.norm_vgg: call once (or host.prepare_norm_vgg(device)) before capturing a CUDA graph")
        _vgg_consts[key] = (torch.tensor(VGG_MEAN, dtype=img.dtype, device=img.device).view(1, 3, 1, 1),
                            torch.tensor(VGG_STD, dtype=img.dtype, device=img.device).view(1, 3, 1, 1))
    mean, std = _vgg_consts[key]
    return (img - mean) / std


def prepare_norm_vgg(device, dtype=torch.float32):
    """Build norm_vgg's constants ahead of a CUDA-graph capture."""
    norm_vgg(torch.zeros((1, 3, 1, 1), device=device, dtype=dtype))


def make_srgan_optimizers(G, D, lr=1e-5, capturable=False):
    """Adam for G (srgan.py:147), SGD momentum 0.9 Nesterov lr/100 for D (srgan.py:149); torch's single-kernel implementations."""
    gp, dp = list(G.parameters()), list(D.parameters())
    fused = {"fused": True} if (gp and gp[0].is_cuda) else {}
    adam = dict(fused, capturable=True) if (capturable and fused) else fused
    return (torch.optim.Adam(gp, lr=lr, betas=(0.9, 0.999), **adam),
            torch.optim.SGD(dp, lr=lr / 100, momentum=0.9, nesterov=True, **fused))


def srgan_step(G, D, FE, g_opt, d_opt, lr_img, hr_img, bucket_g=None, bucket_d=None, after_update=None):
    """One adversarial iteration as written in srgan.py:256-310 (labels shaped like the decision; `.data` -> `.detach()`).
    With GradBuckets (data parallel) the D gradients are exchanged before D's step and the G gradients before G's step.
    `after_update` is called after each of the two optimizer steps (packed-weight cache: srb200.repack_weights).
    Returns (D_loss, G_loss), detached device scalars."""
    from . import functional as F
    from . import nn_ops
    x_, y_ = norm_vgg(hr_img), norm_vgg(lr_img)
    n = x_.shape[0]
    real_label = torch.ones(n, device=x_.device)
    fake_label = torch.zeros(n, device=x_.device)
    # ---- discriminator (srgan.py:272-287); recon is NOT detached in the reference: G's gradients are computed and discarded
    if bucket_d is not None:
        bucket_d.begin_step()
        if bucket_g is not None:
            bucket_g.begin_step()
    else:
        d_opt.zero_grad()
    D_real_loss = nn_ops.bce_loss(D(x_), real_label)
    recon = G(y_)
    D_fake_loss = nn_ops.bce_loss(D(recon), fake_label)
    D_loss = D_real_loss + D_fake_loss
    D_loss.backward()
    if bucket_d is not None:
        bucket_d.all_reduce()
    d_opt.step()
    if after_update is not None:
        after_update()  # e.g. srb200.repack_weights: the generator phase below already runs through the UPDATED discriminator
    # ---- generator (srgan.py:290-310); D's gradients are computed again and cleared by the next iteration's zero_grad
    if bucket_g is not None:
        bucket_g.begin_step()
        bucket_d.begin_step()
    else:
        g_opt.zero_grad()
    recon = G(y_)
    GAN_loss = nn_ops.bce_loss(D(recon), real_label)
    mse_loss = F.mse_loss(recon, x_)
    real_feature = FE(norm_vgg(hr_img))
    fake_feature = FE(norm_vgg(recon.detach()))
    vgg_loss = F.mse_loss(fake_feature, real_feature.detach())
    G_loss = mse_loss + 6e-3 * vgg_loss + 1e-3 * GAN_loss
    G_loss.backward()
    if bucket_g is not None:
        bucket_g.all_reduce()
    g_opt.step()
    if after_update is not None:
        after_update()
    return D_loss.detach(), G_loss.detach()


# ---- utils.img_interp / utils.shave on the device (utils.py:197-205, 242-269) -------------------------------------------------
_PIL_PRECISION_BITS = 32 - 8 - 2
_coeff_cache = {}


def _pil_bicubic_tables(in_size, out_size, device):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the BICUBIC filter (a = -0.5, support 2): per output index the
    first source index, the tap count and the taps as 22-bit fixed-point integers.  Host arithmetic in double, like Pillow."""
    import math
    key = (in_size, out_size, str(device))
    hit = _coeff_cache.get(key)
    if hit is not None:
        return hit

    def cubic(x, a=-0.5):
        x = abs(x)
        if x < 1.0:
            return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
        if x < 2.0:
            return (((x - 5) * x + 8) * x - 4) * a
        return 0.0

    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds, coeffs = [], []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        cnt = min(int(center + support + 0.5), in_size) - xmin
        w = [cubic((x + xmin - center + 0.5) / filterscale) for x in range(cnt)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        q = [int(-0.5 + v * (1 << _PIL_PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << _PIL_PRECISION_BITS)) for v in w]
        bounds += [xmin, cnt]
        coeffs += q + [0] * (ksize - cnt)
    out = (torch.tensor(bounds, dtype=torch.int32, device=device), torch.tensor(coeffs, dtype=torch.int32, device=device), ksize)
    _coeff_cache[key] = out
    return out


def img_interp(imgs, scale_factor, interpolation="bicubic", shave=0):
    """utils.img_interp(imgs, scale_factor) for a CUDA float (N,C,H,W) batch in [0,1], bit-exact with the reference's PIL
    loop (utils.py:242-269), optionally fused with utils.shave(., shave) (the `shave(img_interp(...))` of espcn.py:149)."""
    import ctypes
    from ._lib import lib, check
    if interpolation != "bicubic":
        raise RuntimeError("srb200.host.img_interp implements the reference's default 'bicubic' mode")
    if not (imgs.is_cuda and imgs.dtype == torch.float32 and imgs.dim() == 4):
        raise RuntimeError("img_interp needs a CUDA float32 (N,C,H,W) tensor; there is no CPU path")
    imgs = imgs.contiguous()
    n, c, h, w = imgs.shape
    th, tw = int(h * scale_factor), int(w * scale_factor)
    bw, kw, ks_w = _pil_bicubic_tables(w, tw, imgs.device)
    bh, kh, ks_h = _pil_bicubic_tables(h, th, imgs.device)
    assert ks_w == ks_h
    out = torch.empty((n, c, th - 2 * shave, tw - 2 * shave), dtype=torch.float32, device=imgs.device)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    check(lib.srb_img_interp_bicubic(vp(imgs), vp(out), n, c, h, w, th, tw, vp(bw), vp(kw), vp(bh), vp(kh), ks_w, shave,
                                     ctypes.c_void_p(torch.cuda.current_stream(imgs.device).cuda_stream)))
    return out


def shave(imgs, border_size=0):
    """utils.shave (utils.py:197-205): crop `border_size` pixels from every side (a view; no kernel needed)."""
    if border_size == 0:
        return imgs
    return imgs[..., border_size:-border_size, border_size:-border_size]
