"""GPU debugging aid: tensor-core (tcgen05) path vs the fp32 CUDA-core path of the same library, per shape."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-super-resolution-model-collection_b200"))
import torch
import srb200
from srb200 import _lib

torch.manual_seed(0)
dev = "cuda:0"
CASES = [  # N Cin H W Cout k pad act res ps
    (1, 32, 16, 8, 16, 1, 0, None, False, 1),
    (1, 32, 16, 8, 32, 3, 1, None, False, 1),
    (2, 64, 16, 16, 64, 3, 1, "relu", False, 1),
    (2, 64, 12, 20, 64, 3, 1, None, True, 1),
    (3, 64, 16, 16, 32, 3, 0, "relu", False, 1),
    (2, 32, 13, 14, 3, 3, 0, None, False, 4),
    (2, 64, 16, 16, 3, 3, 1, None, True, 1),
    (1, 64, 16, 16, 3, 9, 4, None, False, 1),
    (1, 96, 9, 9, 160, 3, 1, "relu", False, 1),
    (1, 256, 8, 8, 256, 3, 1, "relu", False, 1),
    (2, 64, 8, 8, 64, 3, 1, "prelu", False, 2),
    (8, 64, 128, 128, 64, 3, 1, "relu", False, 1),
    (4, 256, 32, 32, 256, 3, 1, "relu", False, 1),
]
only = int(sys.argv[1]) if len(sys.argv) > 1 else None
for idx, (N, Cin, H, W, Cout, k, pad, act, res, ps) in enumerate(CASES):
    if only is not None and idx != only:
        continue
    x = torch.randn(N, Cin, H, W, device=dev).contiguous(memory_format=torch.channels_last)
    w = torch.randn(Cout * ps * ps, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout * ps * ps, device=dev) * 0.1
    alpha = torch.tensor([0.25], device=dev)
    out = {}
    for math in ("fp32", "auto"):
        srb200.set_math(math)
        xx = x.clone().requires_grad_(True)
        ww = w.clone().requires_grad_(True)
        y = srb200.conv2d(xx, ww, b, 1, pad, activation=act, alpha=alpha if act == "prelu" else None, pixel_shuffle=ps)
        r = torch.randn_like(y) if res else None
        torch.manual_seed(1)
        gy = torch.randn_like(y)
        y.backward(gy)
        torch.cuda.synchronize()
        out[math] = (y.detach().float(), xx.grad.float(), ww.grad.float())
    p = _lib.ConvParams(N, Cin, H, W, Cout, k, k, 1, pad, 0, 0, ps, 0, 0.2, 2)
    tp = [_lib.lib.srb_conv_uses_tensor_path(ctypes.byref(p), i, 1, 1 if Cout % 32 == 0 else 0) for i in range(3)]
    def rel(a, b):
        return ((a.double() - b.double()).norm() / b.double().norm()).item()
    ey, ex, ew = (rel(a, b) for a, b in zip(out["auto"], out["fp32"]))
    print("case %2d N%d Cin%d %dx%d Cout%d k%d p%d act=%s ps%d tensor_path(f,d,w)=%s  rel y %.3e dx %.3e dw %.3e"
          % (idx, N, Cin, H, W, Cout, k, pad, act, ps, tp, ey, ex, ew), flush=True)
    if ey > 2e-3:
        d = (out["auto"][0] - out["fp32"][0]).abs()
        yy = out["fp32"][0]
        print("   y mismatch: max abs diff %.4e at %s ; ref absmax %.3e; auto absmax %.3e; frac bad %.3f"
              % (d.max().item(), tuple(torch.nonzero(d == d.max())[0].tolist()), yy.abs().max().item(),
                 out["auto"][0].abs().max().item(), (d > 1e-2 * yy.abs().max()).float().mean().item()))
        bad = (d > 1e-2 * yy.abs().max())
        print("   bad by channel (first 16):", bad.float().mean(dim=(0, 2, 3))[:16].tolist())
        print("   bad by row:", bad.float().mean(dim=(0, 1, 3)).tolist()[:20])
        print("   bad by col:", bad.float().mean(dim=(0, 1, 2)).tolist()[:20])
