"""Debug: per-role row timeline of k_conv_rs on CTA 0 (srb_debug_set_trace)."""
import ctypes, sys, os
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-super-resolution-model-collection_b200"))
import srb200
from srb200 import _lib

def run(name, N, Ci, H, W, Co, k, p, flags):
    dev = torch.device("cuda:0")
    x = torch.randn(N, Ci, H, W, device=dev).contiguous(memory_format=torch.channels_last)
    w = torch.randn(Co, Ci, k, k, device=dev) * 0.05
    b = torch.randn(Co, device=dev)
    dbg = _lib.lib.srb_debug_set_flags; dbg.argtypes = [ctypes.c_int]; dbg.restype = None
    fn = _lib.lib.srb_debug_set_trace; fn.argtypes = [ctypes.c_void_p, ctypes.c_longlong]; fn.restype = None
    dbg(flags)
    for _ in range(3):
        srb200.conv2d(x, w, b, 1, p, activation="relu")
    torch.cuda.synchronize()
    buf = torch.zeros(4 * 128, dtype=torch.int64, device=dev)
    fn(ctypes.c_void_p(buf.data_ptr()), 64)
    srb200.conv2d(x, w, b, 1, p, activation="relu")
    torch.cuda.synchronize()
    fn(None, 0); dbg(0)
    t = buf.cpu().numpy().reshape(4, 128).astype(np.float64)
    t0 = t[t > 0].min()
    print("== %s flags %d" % (name, flags))
    for role, lbl in enumerate(["producer (stage free, chunk 0)", "mma row start", "epilogue half 0 (row ready)", "epilogue half 1 (row ready)"]):
        v = t[role][t[role] > 0]
        rel = (v - t0) / 1e3
        d = np.diff(rel)
        print("  %-32s n=%3d first %.2f us last %.2f us; deltas (us): %s" % (lbl, len(v), rel[0] if len(v) else -1, rel[-1] if len(v) else -1,
              " ".join("%.2f" % q for q in d[:40])))

if __name__ == "__main__":
    for fl in (15, 11, 8):
        run("espcn L2", 128, 64, 60, 60, 32, 3, 0, fl)
    run("vdsr body", 64, 64, 128, 128, 64, 3, 1, 15)
