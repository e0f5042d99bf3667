"""ctypes binding of libsrb200.so (C-ABI declared in include/srb200.h).

The library is the product: there is no Python/torch fallback.  Importing this module on a
machine without the built .so raises immediately (build with `python __graft_entry__.py` or
`make -C pytorch-super-resolution-model-collection_b200/csrc`).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsrb200.so")

ACT_NONE, ACT_RELU, ACT_PRELU, ACT_LRELU = 0, 1, 2, 3
MATH_FP32, MATH_TF32, MATH_AUTO, MATH_EXACT, MATH_BF16 = 0, 1, 2, 3, 4
F32, BF16, U8 = 0, 1, 2
PASS_FPROP, PASS_DGRAD, PASS_WGRAD = 0, 1, 2


class Tensor4(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("sn", ctypes.c_int64), ("sc", ctypes.c_int64),
                ("sh", ctypes.c_int64), ("sw", ctypes.c_int64), ("dtype", ctypes.c_int32)]


class ConvParams(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int32), ("Cin", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32),
                ("Cout", ctypes.c_int32), ("kh", ctypes.c_int32), ("kw", ctypes.c_int32),
                ("stride", ctypes.c_int32), ("pad", ctypes.c_int32), ("out_pad", ctypes.c_int32),
                ("transposed", ctypes.c_int32), ("ps", ctypes.c_int32), ("act", ctypes.c_int32),
                ("slope", ctypes.c_float), ("math", ctypes.c_int32)]


# every symbol include/srb200.h declares: name -> (restype, argtypes)
_P = ctypes.POINTER
_vp = ctypes.c_void_p
SYMBOLS = {
    "srb_version": (ctypes.c_int, []),
    "srb_last_error": (ctypes.c_char_p, []),
    "srb_conv_out_hw": (ctypes.c_int, [_P(ConvParams), _P(ctypes.c_int32), _P(ctypes.c_int32)]),
    "srb_conv_uses_tensor_path": (ctypes.c_int, [_P(ConvParams), ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "srb_conv_backward_folds_ps": (ctypes.c_int, [_P(ConvParams), ctypes.c_int, ctypes.c_int]),
    "srb_conv_workspace_bytes": (ctypes.c_size_t, [_P(ConvParams), ctypes.c_int]),
    "srb_conv_describe_plan": (ctypes.c_int, [_P(ConvParams), ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]),
    "srb_conv_fprop": (ctypes.c_int, [_P(ConvParams), _P(Tensor4), _vp, _vp, _vp, _P(Tensor4), _P(Tensor4),
                                      _P(Tensor4), _vp, _vp, ctypes.c_size_t, _vp]),
    "srb_conv_fprop_loss": (ctypes.c_int, [_P(ConvParams), _P(Tensor4), _vp, _vp, _P(Tensor4), ctypes.c_int, _P(Tensor4),
                                           _P(Tensor4), ctypes.c_int, _vp, _vp, ctypes.c_size_t, _vp]),
    "srb_scale_by_scalar": (ctypes.c_int, [_vp, ctypes.c_int64, _vp, ctypes.c_int, _vp]),
    "srb_weight_cache_enable": (ctypes.c_int, [ctypes.c_int]),
    "srb_weight_cache_repack": (ctypes.c_int, [_vp]),
    "srb_weight_cache_entries": (ctypes.c_int, []),
    "srb_adam_step_flat": (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_int64, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                          ctypes.c_float, _vp, _vp]),
    "srb_act_bwd": (ctypes.c_int, [_P(ConvParams), _P(Tensor4), _P(Tensor4), _vp, _P(Tensor4), _vp, _vp]),
    "srb_conv_dgrad": (ctypes.c_int, [_P(ConvParams), _P(Tensor4), _vp, _P(Tensor4), _vp, _P(Tensor4), _vp, ctypes.c_size_t,
                                      _vp]),
    "srb_conv_dgrad_add": (ctypes.c_int, [_P(ConvParams), _P(Tensor4), _vp, _P(Tensor4), _vp, _P(Tensor4), _P(Tensor4), _vp,
                                          ctypes.c_size_t, _vp]),
    "srb_conv_wgrad": (ctypes.c_int, [_P(ConvParams), _P(Tensor4), _P(Tensor4), _vp, _vp, ctypes.c_float,
                                      ctypes.c_int, _vp, ctypes.c_size_t, _vp]),
    "srb_pixel_unshuffle": (ctypes.c_int, [_P(ConvParams), _P(Tensor4), _P(Tensor4), _vp]),
    "srb_prelu_fwd": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int64, _vp]),
    "srb_prelu_bwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, ctypes.c_int64, _vp]),
    "srb_image_to_tensor": (ctypes.c_int, [_vp, _vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                           ctypes.c_float, _vp]),
    "srb_img_interp_bicubic": (ctypes.c_int, [_vp, _vp] + [ctypes.c_int32] * 6 + [_vp, _vp, _vp, _vp, ctypes.c_int32, ctypes.c_int32, _vp]),
    "srb_round_tf32": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, _vp]),
    "srb_loss_workspace_bytes": (ctypes.c_size_t, []),
    "srb_loss_fwd": (ctypes.c_int, [ctypes.c_int, _vp, _vp, ctypes.c_int64, _vp, _vp, ctypes.c_size_t, _vp]),
    "srb_loss_bwd": (ctypes.c_int, [ctypes.c_int, _vp, _vp, ctypes.c_int64, _vp, _vp, _vp]),
    "srb_bn_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int32]),
    "srb_bn_fwd": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int32, _vp, _vp, _vp, _vp, ctypes.c_int, ctypes.c_float,
                                  ctypes.c_float, _vp, _vp, ctypes.c_int, ctypes.c_float, _vp, _vp, ctypes.c_int, _vp,
                                  ctypes.c_size_t, _vp]),
    "srb_bn_bwd": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int64, ctypes.c_int32, _vp, _vp, _vp, _vp, ctypes.c_int, ctypes.c_float,
                                  _vp, _vp, _vp, _vp, ctypes.c_float, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_size_t, _vp]),
    "srb_linear_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int32, ctypes.c_int32]),
    "srb_linear_fwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _vp, ctypes.c_size_t, _vp]),
    "srb_linear_bwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_float,
                                      ctypes.c_int, _vp, ctypes.c_size_t, _vp]),
    "srb_maxpool2_fwd": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _vp]),
    "srb_maxpool2_bwd": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _vp]),
    "srb_bce_fwd": (ctypes.c_int, [_vp, _vp, ctypes.c_int32, _vp, _vp]),
    "srb_bce_bwd": (ctypes.c_int, [_vp, _vp, ctypes.c_int32, _vp, _vp, _vp]),
    "srb_allreduce_inplace": (ctypes.c_int, [_vp, _vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int64, _vp, _vp]),
    "srb_launch_count": (ctypes.c_int64, []),
}


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "libsrb200.so not found at %s -- the CUDA engine is not built; there is no fallback path. "
            "Run `python __graft_entry__.py` (or make -C .../csrc)." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class SrbError(RuntimeError):
    pass


EUNSUPPORTED = -2  # SRB_EUNSUPPORTED (include/srb200.h)


def check(rc):
    if rc != 0:
        raise SrbError("libsrb200 error %d: %s" % (rc, lib.srb_last_error().decode("utf-8", "replace")))


def t4(t):
    """torch.Tensor (4-D, fp32 or bf16, CUDA) -> Tensor4 (logical NCHW + element strides + dtype)."""
    s = t.stride()
    import torch
    dt = BF16 if t.dtype == torch.bfloat16 else (U8 if t.dtype == torch.uint8 else F32)
    return Tensor4(t.data_ptr(), s[0], s[1], s[2], s[3], dt)


def _torch_bf16():
    import torch
    return torch.bfloat16


def launch_count():
    return int(lib.srb_launch_count())
