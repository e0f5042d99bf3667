"""Debug: per-CTA phase timeline of k_conv_sl for one layer (uses the undeclared srb_debug_set_trace hook)."""
import ctypes, sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-super-resolution-model-collection_b200"))
import srb200
from srb200 import _lib
import numpy as np

def run(name, N, Ci, H, W, Co, k, p, ps=1, act="relu"):
    dev = torch.device("cuda:0")
    x = torch.randn(N, Ci, H, W, device=dev)
    if Ci % 4 == 0 and Ci >= 8:
        x = x.contiguous(memory_format=torch.channels_last)
    w = torch.randn(Co * ps * ps, Ci, k, k, device=dev) * 0.05
    b = torch.randn(Co * ps * ps, device=dev)
    for _ in range(3):
        y = srb200.conv2d(x, w, b, 1, p, activation=act, pixel_shuffle=ps)
    torch.cuda.synchronize()
    maxc = 1 << 16
    buf = torch.zeros(maxc * 8, dtype=torch.int64, device=dev)
    fn = _lib.lib.srb_debug_set_trace
    fn.argtypes = [ctypes.c_void_p, ctypes.c_longlong]
    fn.restype = None
    fn(ctypes.c_void_p(buf.data_ptr()), maxc)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = srb200.conv2d(x, w, b, 1, p, activation=act, pixel_shuffle=ps)
    e1.record()
    torch.cuda.synchronize()
    fn(None, 0)
    t = buf.cpu().numpy().reshape(-1, 8)
    t = t[t[:, 0] != 0]
    t0 = t[:, 0].min()
    d = lambda a, b: (t[:, b] - t[:, a]) / 1e3
    print("%s: %d CTAs, call %.1f us (incl. pack kernels), kernel span %.1f us" % (name, len(t), e0.elapsed_time(e1) * 1e3, (t[:, 6].max() - t0) / 1e3))
    cyc = (t[:, 3] >> 32).astype(np.float64)
    ns = (t[:, 3] & 0xffffffff).astype(np.float64)
    print("   MMA loop: mean %.2f us, %.0f SM cycles -> effective SM clock %.0f MHz" % (ns.mean() / 1e3, cyc.mean(), (cyc / ns).mean() * 1e3))
    for lbl, a_, b_ in (("setup (alloc+init+sync)", 0, 1), ("wait first operands", 1, 2),
                        ("first band epilogue", 4, 5), ("CTA lifetime", 0, 6)):
        v = d(a_, b_)
        print("   %-26s mean %7.2f us  p10 %7.2f  p50 %7.2f  p90 %7.2f" % (lbl, v.mean(), np.percentile(v, 10), np.percentile(v, 50), np.percentile(v, 90)))
    # concurrency per SM
    sm = t[:, 7]
    life = d(0, 6)
    print("   CTAs/SM %.1f, sum(lifetime)/SM/span = %.2f resident CTAs on average" % (len(t) / len(set(sm.tolist())), life.sum() / len(set(sm.tolist())) / ((t[:, 6].max() - t0) / 1e3)))

if __name__ == "__main__":
    dbg = _lib.lib.srb_debug_set_flags
    dbg.argtypes = [ctypes.c_int]
    dbg.restype = None
    if len(sys.argv) > 1:
        for flags in (0, 1, 2, 3):
            dbg(flags)
            print("#### debug flags %d (1: empty epilogue, 2: A tiles loaded once)" % flags)
            run("espcn L1 (c4)", 128, 3, 64, 64, 64, 5, 0)
            run("espcn L2", 128, 64, 60, 60, 32, 3, 0)
            run("vdsr body", 64, 64, 128, 128, 64, 3, 1)
        dbg(0)
        sys.exit(0)
    run("espcn L1 (c4)", 128, 3, 64, 64, 64, 5, 0)
    run("espcn L2", 128, 64, 60, 60, 32, 3, 0)
    run("espcn L3 +PS4", 128, 32, 58, 58, 3, 3, 0, ps=4, act=None)
    run("vdsr body", 64, 64, 128, 128, 64, 3, 1)
