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
