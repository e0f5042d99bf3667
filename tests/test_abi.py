"""CPU suite: the C-ABI library loads without a GPU and exports every symbol include/srb200.h declares;
host-only entry points (shape arithmetic, argument validation, error strings) behave."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
import srb200
from srb200 import _lib


def _declared():
    text = open(os.path.join(ROOT, "include", "srb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(srb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported():
    names = _declared()
    assert len(names) >= 13
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libsrb200.so does not export %s" % n
    assert sorted(_lib.SYMBOLS) == names, "python binding table out of sync with the header"


def test_version_and_error_string():
    assert _lib.lib.srb_version() == 100
    p = _lib.ConvParams(1, 3, 8, 8, 4, 3, 3, 0, 1, 0, 0, 1, 0, 0.2, 0)  # stride 0
    ho, wo = ctypes.c_int32(), ctypes.c_int32()
    rc = _lib.lib.srb_conv_out_hw(ctypes.byref(p), ctypes.byref(ho), ctypes.byref(wo))
    assert rc == -1
    assert b"stride" in _lib.lib.srb_last_error()
    with pytest.raises(srb200.SrbError):
        _lib.check(rc)


@pytest.mark.parametrize("H,W,k,s,p,tr,op,exp", [
    (64, 64, 9, 1, 0, 0, 0, (56, 56)), (60, 60, 3, 1, 0, 0, 0, (58, 58)), (128, 128, 3, 2, 1, 0, 0, (64, 64)),
    (28, 28, 9, 4, 3, 1, 1, (112, 112)), (5, 6, 4, 2, 1, 1, 0, (10, 12)), (32, 32, 3, 1, 1, 0, 0, (32, 32)),
])
def test_out_hw(H, W, k, s, p, tr, op, exp):
    prm = _lib.ConvParams(1, 4, H, W, 4, k, k, s, p, op, tr, 1, 0, 0.2, 2)
    ho, wo = ctypes.c_int32(), ctypes.c_int32()
    assert _lib.lib.srb_conv_out_hw(ctypes.byref(prm), ctypes.byref(ho), ctypes.byref(wo)) == 0
    assert (ho.value, wo.value) == exp


def test_rejects_bad_requests():
    ho, wo = ctypes.c_int32(), ctypes.c_int32()
    bad = [
        _lib.ConvParams(1, 4, 2, 2, 4, 5, 5, 1, 0, 0, 0, 1, 0, 0.2, 0),   # kernel larger than input
        _lib.ConvParams(1, 4, 8, 8, 4, 3, 3, 1, 1, 0, 0, 1, 7, 0.2, 0),   # unknown activation
        _lib.ConvParams(1, 4, 8, 8, 4, 3, 3, 2, 1, 2, 1, 1, 0, 0.2, 0),   # out_pad >= stride
        _lib.ConvParams(1, 4, 8, 8, 4, 3, 3, 1, 1, 1, 0, 1, 0, 0.2, 0),   # out_pad on a plain conv
    ]
    for p in bad:
        assert _lib.lib.srb_conv_out_hw(ctypes.byref(p), ctypes.byref(ho), ctypes.byref(wo)) == -1
    p = _lib.ConvParams(1, 4, 8, 8, 4, 3, 3, 2, 1, 0, 1, 2, 0, 0.2, 0)        # PixelShuffle + transposed
    assert _lib.lib.srb_conv_out_hw(ctypes.byref(p), ctypes.byref(ho), ctypes.byref(wo)) == -2


def test_workspace_query_is_host_only():
    p = _lib.ConvParams(16, 64, 32, 32, 64, 3, 3, 1, 1, 0, 0, 1, 1, 0.2, 2)
    for ps in (0, 1, 2):
        assert _lib.lib.srb_conv_workspace_bytes(ctypes.byref(p), ps) >= 256
