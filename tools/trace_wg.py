"""Debug: time srb_conv_wgrad for a few layers under the debug flags (2: stages TMA-loaded once, 4: no MMAs)."""
import ctypes, sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-super-resolution-model-collection_b200"))
import srb200
from srb200 import _lib
from srb200._lib import lib, ConvParams, t4

def run(name, N, Ci, H, W, Co, k, p, bias=True):
    dev = torch.device("cuda:0")
    x = torch.randn(N, Ci, H, W, device=dev)
    if Ci % 4 == 0 and Ci >= 8:
        x = x.contiguous(memory_format=torch.channels_last)
    Ho, Wo = H + 2 * p - k + 1, W + 2 * p - k + 1
    dz = torch.randn(N, Co, Ho, Wo, device=dev).contiguous(memory_format=torch.channels_last)
    dw = torch.empty(Co, Ci, k, k, device=dev)
    db = torch.empty(Co, device=dev) if bias else None
    prm = ConvParams(N, Ci, H, W, Co, k, k, 1, p, 0, 0, 1, 0, 0.2, _lib.MATH_AUTO)
    ws = torch.empty(int(lib.srb_conv_workspace_bytes(ctypes.byref(prm), 2)) + 1024, dtype=torch.uint8, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    tx, tz = t4(x), t4(dz)
    def call():
        rc = lib.srb_conv_wgrad(ctypes.byref(prm), ctypes.byref(tx), ctypes.byref(tz), ctypes.c_void_p(dw.data_ptr()),
                                ctypes.c_void_p(db.data_ptr()) if bias else None, ctypes.c_float(1.0), 0,
                                ctypes.c_void_p(ws.data_ptr()), ws.numel(), st)
        assert rc == 0, lib.srb_last_error()
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        call()
    e1.record()
    torch.cuda.synchronize()
    buf = ctypes.create_string_buffer(512)
    lib.srb_conv_describe_plan(ctypes.byref(prm), 2, buf, 512)
    print("%-12s %7.1f us/call   %s" % (name, e0.elapsed_time(e1) * 100, buf.value.decode()[:230]))

if __name__ == "__main__":
    dbg = lib.srb_debug_set_flags
    dbg.argtypes = [ctypes.c_int]
    dbg.restype = None
    for flags in (0, 2, 4, 6, 8, 32):
        dbg(flags)
        print("#### debug flags %d (2: TMA once, 4: no MMA, 8: no db sums, 16: no smem zeroing, 32: no partial dump)" % flags)
        only = sys.argv[1] if len(sys.argv) > 1 else ""
        if not only:
            run("espcn L1", 128, 3, 64, 64, 64, 5, 0)
            run("espcn L2", 128, 64, 60, 60, 32, 3, 0)
            run("espcn L3", 128, 32, 58, 58, 48, 3, 0)
            run("vdsr body", 64, 64, 128, 128, 64, 3, 1, bias=False)
            run("edsr64 body", 32, 64, 32, 32, 64, 3, 1)
        run("edsr256 body", 32, 256, 32, 32, 256, 3, 1)
        run("edsr256 up2", 32, 256, 64, 64, 1024, 3, 1)
    dbg(0)
