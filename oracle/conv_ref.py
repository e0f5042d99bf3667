"""TEST INFRASTRUCTURE ONLY -- ctypes loader for oracle/conv_ref.c (plain-C conv arithmetic)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libconvref.so")


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def _lib():
    if not os.path.isfile(_SO):
        build()
    return ctypes.CDLL(_SO)


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def conv2d_fwd(x, w, b, st, pad):
    N, C, H, W = x.shape
    O, _, kh, kw = w.shape
    Ho, Wo = (H + 2 * pad - kh) // st + 1, (W + 2 * pad - kw) // st + 1
    y = np.empty((N, O, Ho, Wo), np.float32)
    x, px = _f(x); w, pw = _f(w)
    pb = None
    if b is not None:
        b, pb = _f(b)
    _lib().ref_conv2d_fwd(px, pw, pb, y.ctypes.data_as(ctypes.c_void_p), N, C, H, W, O, kh, kw, st, pad)
    return y


def conv2d_bwd_data(dy, w, xshape, st, pad):
    N, C, H, W = xshape
    O, _, kh, kw = w.shape
    dx = np.empty(xshape, np.float32)
    dy, pdy = _f(dy); w, pw = _f(w)
    _lib().ref_conv2d_bwd_data(pdy, pw, dx.ctypes.data_as(ctypes.c_void_p), N, C, H, W, O, kh, kw, st, pad,
                               dy.shape[2], dy.shape[3])
    return dx


def conv2d_bwd_weight(x, dy, wshape, st, pad, bias=True):
    N, C, H, W = x.shape
    O, _, kh, kw = wshape
    dw = np.empty(wshape, np.float32)
    db = np.empty((O,), np.float32) if bias else None
    x, px = _f(x); dy, pdy = _f(dy)
    _lib().ref_conv2d_bwd_weight(px, pdy, dw.ctypes.data_as(ctypes.c_void_p),
                                 db.ctypes.data_as(ctypes.c_void_p) if bias else None,
                                 N, C, H, W, O, kh, kw, st, pad)
    return dw, db


def conv_transpose2d_fwd(x, w, b, st, pad, out_pad):
    N, Ci, H, W = x.shape
    _, Co, kh, kw = w.shape
    Ho, Wo = (H - 1) * st - 2 * pad + kh + out_pad, (W - 1) * st - 2 * pad + kw + out_pad
    y = np.empty((N, Co, Ho, Wo), np.float32)
    x, px = _f(x); w, pw = _f(w)
    pb = None
    if b is not None:
        b, pb = _f(b)
    _lib().ref_conv_transpose2d_fwd(px, pw, pb, y.ctypes.data_as(ctypes.c_void_p), N, Ci, H, W, Co, kh, kw, st, pad,
                                    out_pad)
    return y


def pixel_shuffle(t, r):
    N, Crr, H, W = t.shape
    C = Crr // (r * r)
    out = np.empty((N, C, H * r, W * r), np.float32)
    t, pt = _f(t)
    _lib().ref_pixel_shuffle(pt, out.ctypes.data_as(ctypes.c_void_p), N, C, H, W, r)
    return out


def act_fwd(x, act, a=0.0):
    x, px = _f(x)
    y = np.empty_like(x)
    _lib().ref_act_fwd(px, y.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(x.size), act, ctypes.c_float(a))
    return y


def act_bwd(x, dy, act, a=0.0):
    x, px = _f(x); dy, pdy = _f(dy)
    dx = np.empty_like(x)
    fn = _lib().ref_act_bwd
    fn.restype = ctypes.c_double
    da = fn(px, pdy, dx.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(x.size), act, ctypes.c_float(a))
    return dx, da
