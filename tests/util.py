"""Helpers shared by the tests (error metrics, golden loading)."""
import os

import numpy as np
import torch

from conftest import GOLDEN


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    d = (a - b).norm().item()
    n = b.norm().item()
    return d / n if n > 0 else d


def rel_max(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    n = b.abs().max().item()
    d = (a - b).abs().max().item()
    return d / n if n > 0 else d


def load_state(net, blob):
    sd = {k[len("param:"):]: torch.from_numpy(v) for k, v in blob.items() if k.startswith("param:")}
    missing = net.load_state_dict(sd, strict=True)
    return missing


DEEP = {  # fixture name -> (oracle / srb200.models key, loss)
    "vdsr18": ("vdsr", "mse"),
    "edsr256x32": ("edsr", "l1"),
    "srgan_g16": ("srgan_g", "mse"),
    "srgan_d": ("srgan_d", "bce"),
}


def digest_of(t):
    """Same digest as oracle/make_golden.py:digest (norm, sum, 64 seeded samples)."""
    flat = torch.as_tensor(t).detach().reshape(-1).double().cpu()
    n = flat.numel()
    idx = torch.randint(0, n, (64,), generator=torch.Generator().manual_seed(n % 2147483647))
    return flat.norm().item(), flat.sum().item(), flat[idx].numpy()


def digest_close(t, blob, tag, tol):
    """|norm - norm_ref| and the sampled entries agree to `tol` relative to the reference norm (per-element scale)."""
    norm, total, sample = digest_of(t)
    ref_norm = float(blob[tag + ":norm"])
    n = torch.as_tensor(t).numel()
    scale = max(ref_norm, 1e-30)
    if abs(norm - ref_norm) > tol * scale:
        return False
    rms = scale / max(n, 1) ** 0.5
    return bool(np.all(np.abs(sample - blob[tag + ":sample"]) <= tol * 40 * rms + tol * np.abs(blob[tag + ":sample"])))


def loss_of(kind, y, tgt):
    import torch.nn.functional as TF
    if kind == "l1":
        return TF.l1_loss(y, tgt)
    if kind == "bce":
        return TF.binary_cross_entropy(y, tgt)
    return TF.mse_loss(y, tgt)


def set_debug_flags(flags):
    """srb_debug_set_flags + drop srb200's cached workspace sizes: the planner's choice (and therefore the workspace a call
    needs) depends on the flags, and the Python side caches srb_conv_workspace_bytes per layer."""
    import ctypes
    from srb200 import _lib, functional
    f = _lib.lib.srb_debug_set_flags
    f.argtypes = [ctypes.c_int]
    f.restype = None
    f(int(flags))
    functional._ws_cache.clear()
