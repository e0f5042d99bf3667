"""Print the tile plans libsrb200 picks for the layers of the benchmark configs (host only)."""
import ctypes, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-super-resolution-model-collection_b200"))
from srb200 import _lib
L = [("espcn L1", 128, 3, 64, 64, 64, 5, 0, 1), ("espcn L2", 128, 64, 60, 60, 32, 3, 0, 1), ("espcn L3", 128, 32, 58, 58, 48, 3, 0, 1),
     ("vdsr in", 64, 3, 128, 128, 64, 3, 1, 1), ("vdsr body", 64, 64, 128, 128, 64, 3, 1, 1), ("vdsr out", 64, 64, 128, 128, 3, 3, 1, 1),
     ("edsr64 body", 32, 64, 32, 32, 64, 3, 1, 1), ("edsr64 up1", 32, 64, 32, 32, 64, 3, 1, 2), ("edsr256 body", 32, 256, 32, 32, 256, 3, 1, 1),
     ("edsr256 up2", 32, 256, 64, 64, 256, 3, 1, 2), ("srcnn L1", 16, 3, 64, 64, 64, 9, 0, 1), ("srcnn L2", 16, 64, 56, 56, 32, 5, 0, 1),
     ("fsrcnn mid", 16, 12, 28, 28, 12, 3, 1, 1)]
buf = ctypes.create_string_buffer(512)
for name, N, Ci, H, W, Co, k, p, ps in L:
    prm = _lib.ConvParams(N, Ci, H, W, Co, k, k, 1, p, 0, 0, ps, 0, 0.2, _lib.MATH_AUTO)
    for pas, pn in ((0, "fprop"), (1, "dgrad"), (2, "wgrad")):
        _lib.lib.srb_conv_describe_plan(ctypes.byref(prm), pas, buf, 512)
        print("%-13s %-5s %s" % (name, pn, buf.value.decode()))
