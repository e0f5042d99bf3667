"""Debug: time single conv layers (fprop through srb200.conv2d) under the kernel debug flags.
flags: 0 normal, 1 empty epilogue, 2 operands loaded once, 4 no MMAs, 128 force the slot-linear kernel"""
import ctypes, sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-super-resolution-model-collection_b200"))
import srb200
from srb200 import _lib

LAYERS = [("espcn L2 64->32", 128, 64, 60, 60, 32, 3, 0, 1, "relu"), ("espcn L2d 32->64", 128, 32, 58, 58, 64, 3, 2, 1, None),
          ("espcn L3 32->48 PS4", 128, 32, 58, 58, 3, 3, 0, 4, None), ("vdsr body 64->64", 64, 64, 128, 128, 64, 3, 1, 1, "relu"),
          ("edsr64 body", 32, 64, 32, 32, 64, 3, 1, 1, "relu"), ("espcn L3d 48->32", 128, 48, 56, 56, 32, 3, 2, 1, None),
          ("srgan G body b16", 16, 64, 32, 32, 64, 3, 1, 1, "relu"), ("srgan D 128->128 64^2", 16, 128, 64, 64, 128, 3, 1, 1, "lrelu"),
          ("srgan D 256->256 32^2", 16, 256, 32, 32, 256, 3, 1, 1, "lrelu"), ("edsr64 up 64->256 ps2", 32, 64, 32, 32, 64, 3, 1, 2, None),
          ("espcn L1 3->64 k5", 128, 3, 64, 64, 64, 5, 0, 1, "relu")]
if os.environ.get("ONLY"):  # e.g. ONLY="espcn L1" under ncu
    LAYERS = [l for l in LAYERS if l[0].startswith(os.environ["ONLY"])]

def timeit(f, n=10, reps=5):
    """GPU time per call: n calls captured into one CUDA graph (no host launch overhead), best of `reps` replays."""
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(n):
                f()
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n)
    return best

if __name__ == "__main__" and not os.environ.get("WGRAD"):
    dbg = _lib.lib.srb_debug_set_flags
    dbg.argtypes = [ctypes.c_int]
    dbg.restype = None
    dev = torch.device("cuda:0")
    flags = [int(v) for v in sys.argv[1:]] or [0, 1, 2, 4, 3, 7, 128]
    for name, N, Ci, H, W, Co, k, p, ps, act in LAYERS:
        x = torch.randn(N, Ci, H, W, device=dev)
        if Ci > 4:
            x = x.contiguous(memory_format=torch.channels_last)
        w = torch.randn(Co * ps * ps, Ci, k, k, device=dev) * 0.05
        b = torch.randn(Co * ps * ps, device=dev)
        out = []
        for fl in flags:
            dbg(fl)
            out.append("%d: %6.1f us" % (fl, timeit(lambda: srb200.conv2d(x, w, b, 1, p, activation=act, pixel_shuffle=ps))))
        dbg(0)
        print("%-22s %s" % (name, "   ".join(out)), flush=True)


def time_wgrad(flags_list):
    """srb_conv_wgrad (main kernel + finish) per layer through the C-ABI, under the planner debug flags
    (256: never stack filter rows along N, 512: always)."""
    from srb200 import functional as F
    from srb200._lib import lib, check, t4
    dbg = _lib.lib.srb_debug_set_flags
    dbg.argtypes = [ctypes.c_int]
    dbg.restype = None
    dev = torch.device("cuda:0")
    layers = [("espcn L2 wgrad 64->32", 128, 64, 60, 60, 32, 3, 0), ("espcn L3 wgrad 32->48", 128, 32, 58, 58, 48, 3, 0),
              ("vdsr body wgrad 64->64", 64, 64, 128, 128, 64, 3, 1), ("edsr64 body wgrad", 32, 64, 32, 32, 64, 3, 1),
              ("vdsr out-like 64->32", 64, 64, 128, 128, 32, 3, 1), ("espcn L1 wgrad 3->64 k5", 128, 3, 64, 64, 64, 5, 0)]
    for name, N, Ci, H, W, Co, k, p in layers:
        x = torch.randn(N, Ci, H, W, device=dev)
        if Ci >= 8:
            x = x.contiguous(memory_format=torch.channels_last)
        w = torch.randn(Co, Ci, k, k, device=dev) * 0.05
        Ho, Wo = H + 2 * p - k + 1, W + 2 * p - k + 1
        dz = torch.randn(N, Co, Ho, Wo, device=dev).contiguous(memory_format=torch.channels_last)
        dw, db = torch.empty_like(w), torch.empty(Co, device=dev)
        prm = F._params(x, w, 1, p, 0, False, 1, None, 0.2)
        out = []
        for fl in flags_list:
            dbg(fl)
            ws = torch.empty(lib.srb_conv_workspace_bytes(ctypes.byref(prm), _lib.PASS_WGRAD) + 1024, dtype=torch.uint8, device=dev)
            tx, tdz = t4(x), t4(dz)
            def f():
                st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
                check(lib.srb_conv_wgrad(ctypes.byref(prm), ctypes.byref(tx), ctypes.byref(tdz), F._ptr(dw), F._ptr(db), ctypes.c_float(1.0), 0,
                                         F._ptr(ws), ws.numel(), st))
            out.append("%d: %6.1f us" % (fl, timeit(f)))
        dbg(0)
        print("%-30s %s" % (name, "   ".join(out)), flush=True)


def time_loss(flags_list):
    """ESPCN's last layer (32 -> 48, PixelShuffle(4)) fused with the MSE criterion (srb_conv_fprop_loss)."""
    dbg = _lib.lib.srb_debug_set_flags
    dbg.argtypes = [ctypes.c_int]
    dbg.restype = None
    dev = torch.device("cuda:0")
    x = torch.randn(128, 32, 58, 58, device=dev).contiguous(memory_format=torch.channels_last)
    w = torch.randn(48, 32, 3, 3, device=dev) * 0.05
    b = torch.randn(48, device=dev)
    t = torch.rand(128, 3, 224, 224, device=dev)
    out = []
    for fl in flags_list:
        dbg(fl)
        def f():
            with torch.no_grad():
                srb200.conv2d_loss(x, w, b, t, "mse", 1, 0, 4)
        out.append("%d: %6.1f us" % (fl, timeit(f)))
    dbg(0)
    print("%-30s %s" % ("espcn L3 + PS4 + MSE", "   ".join(out)), flush=True)


if __name__ == "__main__" and os.environ.get("LOSS"):
    time_loss([8, 9, 11, 12, 13, 14, 15, 128])

if __name__ == "__main__" and os.environ.get("WGRAD"):
    time_wgrad([int(v) for v in os.environ.get("WGRAD").split(",") if v] if "," in os.environ.get("WGRAD") else [0, 256, 512])
