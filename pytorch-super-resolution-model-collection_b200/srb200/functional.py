"""autograd.Functions over the libsrb200 C-ABI: the fused conv (+bias +act +residual +PixelShuffle),
the transposed conv, and the stand-alone PReLU.  Replaces the ATen ops behind base_networks.py's
blocks (Conv2d :42, ConvTranspose2d :77, PixelShuffle :157, ReLU/PReLU/LeakyReLU :51-56, torch.add :149).

No torch compute op is used on the data path: torch only allocates the tensors and provides the
stream.  A CUDA tensor is required; CPU tensors raise (there is no CPU fallback by design).
"""
import ctypes
import torch

from . import _lib
from ._lib import lib, check, t4, ConvParams

_ACT_CODES = {None: _lib.ACT_NONE, "relu": _lib.ACT_RELU, "prelu": _lib.ACT_PRELU, "lrelu": _lib.ACT_LRELU}

_state = {"math": _lib.MATH_AUTO, "grad_scale": 1.0}


def set_math(mode):
    """'auto' (default): tcgen05 TF32 tensor path where a layer qualifies; 'fp32': CUDA-core fp32 everywhere;
    'exact': fp32-accurate on the tensor cores by hi/lo operand splitting (3xTF32), full-fp32 activations;
    'bf16': bf16 STORAGE (BASELINE cfg4): activations and their gradients are bf16 channels_last tensors feeding
    tcgen05 kind::f16, parameters and their gradients stay fp32, the 3-channel network edges stay fp32 / TF32."""
    _state["math"] = {"auto": _lib.MATH_AUTO, "tf32": _lib.MATH_TF32, "fp32": _lib.MATH_FP32,
                      "exact": _lib.MATH_EXACT, "bf16": _lib.MATH_BF16}[mode]


def get_math():
    return {_lib.MATH_AUTO: "auto", _lib.MATH_TF32: "tf32", _lib.MATH_FP32: "fp32", _lib.MATH_EXACT: "exact",
            _lib.MATH_BF16: "bf16"}[_state["math"]]


def set_grad_scale(s):
    """Factor folded by the wgrad kernels into gradients they write straight into a GradBucket slot
    (1/world_size under data parallel).  Gradients returned through autograd are never scaled here."""
    _state["grad_scale"] = float(s)


_state["fuse_relu_bwd"] = True
_state["act_recorder"] = None


def record_activation_masks(recorder):
    """Test instrumentation: while `recorder` is a list, every ReLU / PReLU / LeakyReLU evaluated by the engine appends
    the boolean pattern (pre-activation > 0) of its output, in execution order (None switches it off).  The parity tests
    replay the CPU oracle's backward on exactly these patterns (activation gradients are discontinuous in z)."""
    _state["act_recorder"] = recorder


def _record(y, residual=None):
    rec = _state["act_recorder"]
    if rec is not None:
        with torch.no_grad():
            rec.append(((y - residual) if residual is not None else y) > 0)


def enable_weight_cache(on=True):
    """Opt-in packed-weight cache of libsrb200 (include/srb200.h: srb_weight_cache_enable): conv calls stop re-packing their
    filters; call `repack_weights()` after EVERY weight update (optimizer.step(), load_state_dict, manual edits).
    `srb200.TrainStepGraphs(..., weight_cache=True)` does both."""
    check(lib.srb_weight_cache_enable(1 if on else 0))


def repack_weights(device=None):
    """Refresh every cached packed filter from the current weights: one launch on the current stream."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    check(lib.srb_weight_cache_repack(ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))


def weight_cache_entries():
    return int(lib.srb_weight_cache_entries())


def set_fuse_relu_backward(on):
    """Fold a ReLU layer's threshold_backward into the dgrad epilogue of its consumer conv (default on).

    With the fusion on, the tensor autograd carries between the two layers is dL/dz (masked), not dL/dy.  Parameter and
    input gradients are bit-identical either way; only an observer of the intermediate activation's gradient could tell.
    `retain_grad()` / `register_hook()` on the activation (called before the consumer runs) switch the fusion off for
    that tensor; `torch.autograd.grad(loss, activation)` cannot be detected -- call set_fuse_relu_backward(False) first."""
    _state["fuse_relu_bwd"] = bool(on)


class _ReluToken:
    """Handshake between a conv whose output is y = ReLU(z) and the conv that consumes y.

    The consumer's dgrad can apply the producer's ReLU mask (y > 0, y being its own saved input) in its epilogue, so
    the producer's backward may skip srb_act_bwd.  Safe by construction: ReLU masking is idempotent, so whenever the
    gradient that reaches the producer is not exactly the tensor the consumer wrote (several consumers, autograd
    accumulation, hooks), the producer simply masks again.  `premasked` identifies that tensor."""
    __slots__ = ("consumers", "premasked", "bits", "observed")

    def __init__(self):
        self.consumers = 0
        self.observed = False  # x.grad is watched (retain_grad / hooks): the consumer must not pre-mask it
        self.premasked = None
        self.bits = None  # packed sign pattern of y written by the producer's fprop epilogue (int16, 16 channels per word)


_tc_cache = {}


def _uses_tensor_path(p, pas, x_cl, y_cl):
    """Cached srb_conv_uses_tensor_path (host-only planner query)."""
    key = (p.N, p.Cin, p.H, p.W, p.Cout, p.kh, p.kw, p.stride, p.pad, p.out_pad, p.transposed, p.ps, p.math, pas,
           bool(x_cl), bool(y_cl))
    r = _tc_cache.get(key)
    if r is None:
        r = bool(lib.srb_conv_uses_tensor_path(ctypes.byref(p), pas, 1 if x_cl else 0, 1 if y_cl else 0))
        _tc_cache[key] = r
    return r


_fold_cache = {}


def _folds_ps(p, x_cl, dz_cl):
    """Cached srb_conv_backward_folds_ps (host-only planner query)."""
    key = (p.N, p.Cin, p.H, p.W, p.Cout, p.kh, p.kw, p.stride, p.pad, p.ps, p.math, bool(x_cl), bool(dz_cl))
    r = _fold_cache.get(key)
    if r is None:
        r = bool(lib.srb_conv_backward_folds_ps(ctypes.byref(p), 1 if x_cl else 0, 1 if dz_cl else 0))
        _fold_cache[key] = r
    return r


_workspaces = {}


def _workspace(device, nbytes):
    """Per-(device, stream) scratch owned by torch's caching allocator, grown geometrically."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes * 1.25), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


_ws_cache = {}


def _ws_bytes(p, pas):
    """Cached srb_conv_workspace_bytes (the planner enumeration costs ~10 us of host time per query)."""
    key = (p.N, p.Cin, p.H, p.W, p.Cout, p.kh, p.kw, p.stride, p.pad, p.out_pad, p.transposed, p.ps, p.math, pas)
    r = _ws_cache.get(key)
    if r is None:
        r = int(lib.srb_conv_workspace_bytes(ctypes.byref(p), pas))
        _ws_cache[key] = r
    return r


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _dense(t):
    """Return t as NCHW-contiguous or channels_last-contiguous memory (whichever it already is, else NCHW)."""
    if t.is_contiguous() or t.is_contiguous(memory_format=torch.channels_last):
        return t
    return t.contiguous()


def _is_cl(t):
    return t.shape[1] > 1 and t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous()


def _out_format(channels):
    # activations that can feed the tensor-core kernels live in NHWC; skinny (3-channel) edges stay NCHW
    return torch.channels_last if (channels % 4 == 0 and channels >= 8) else torch.contiguous_format


def _require_cuda(*ts):
    for t in ts:
        if t is not None and (not t.is_cuda or t.dtype != torch.float32):
            raise RuntimeError("srb200 kernels need CUDA float32 tensors (got %s %s); there is no CPU path"
                               % (t.device, t.dtype))


def _bf16_mode():
    return _state["math"] == _lib.MATH_BF16


def _act_dtype(channels):
    """Storage type of an activation with `channels` channels: bf16 in bf16 mode when its pixel rows are 16-byte
    multiples (C % 8 == 0), fp32 otherwise (all other modes; the 3-channel network edges)."""
    return torch.bfloat16 if (_bf16_mode() and channels % 8 == 0 and channels >= 8) else torch.float32


def _as_act(t):
    """Layout/dtype plumbing at the engine boundary (no arithmetic): dense memory, and in bf16 mode the storage type."""
    t = _dense(t)
    want = _act_dtype(t.shape[1])
    if t.dtype != want:
        t = t.to(want)
    if want == torch.bfloat16 and not _is_cl(t):
        t = t.contiguous(memory_format=torch.channels_last)
    return t


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _params(x, weight, stride, pad, out_pad, transposed, ps, act, slope):
    N, Cin, H, W = x.shape
    if transposed:
        assert weight.shape[0] == Cin, "ConvTranspose2d weight is (Cin, Cout, kh, kw)"
        Cout = weight.shape[1]
    else:
        assert weight.shape[1] == Cin, "Conv2d weight is (Cout*ps*ps, Cin, kh, kw)"
        assert weight.shape[0] % (ps * ps) == 0
        Cout = weight.shape[0] // (ps * ps)
    return ConvParams(N, Cin, H, W, Cout, weight.shape[2], weight.shape[3], stride, pad, out_pad,
                      1 if transposed else 0, ps, _ACT_CODES[act], float(slope), _state["math"])


class SkipGrad(object):
    """Hand-over of the skip connection's gradient inside a residual block (`torch.add(out, residual)`, base_networks.py:149):
    x feeds conv1 AND the skip, so autograd would sum dL/dx = dgrad_conv1(dh) + dL/dy with an extra add pass over dx.  The
    block passes one SkipGrad to both convs: conv2's backward parks dL/dy here instead of returning it as the residual's
    gradient, conv1's backward (which always runs after it) folds it into its dgrad epilogue (srb_conv_dgrad_add)."""
    __slots__ = ("grad", "armed")

    def __init__(self):
        self.grad = None     # dL/dy parked by the block's last conv during backward
        self.armed = False   # set by conv1's forward when it will compute dL/dx


class _FusedConv(torch.autograd.Function):
    """y = PixelShuffle_ps(act(conv(x, w) + b)) + residual   (one kernel forward; act_bwd + wgrad + dgrad backward)."""

    @staticmethod
    def forward(ctx, x, weight, bias, alpha, residual, stride, pad, out_pad, transposed, ps, act, slope, in_token,
                out_token, skip_sink=None, skip_src=None):
        _require_cuda(weight, bias, alpha)
        if not x.is_cuda:
            raise RuntimeError("srb200 kernels need CUDA tensors (got %s); there is no CPU path" % x.device)
        x = _as_act(x)
        weight = weight.contiguous()
        p = _params(x, weight, stride, pad, out_pad, transposed, ps, act, slope)
        ho, wo = ctypes.c_int32(), ctypes.c_int32()
        check(lib.srb_conv_out_hw(ctypes.byref(p), ctypes.byref(ho), ctypes.byref(wo)))
        oshape = (p.N, p.Cout, ho.value * ps, wo.value * ps)
        fmt = _out_format(p.Cout)
        y = torch.empty(oshape, dtype=_act_dtype(p.Cout), device=x.device, memory_format=fmt)
        # PReLU backward needs z itself; relu/lrelu can use sign(y) unless a residual was added on top
        need_preact = act == "prelu" or (act is not None and residual is not None)
        preact = torch.empty_like(y) if need_preact else None
        if residual is not None:
            assert tuple(residual.shape) == oshape, "residual must match the block output"
            residual = _as_act(residual)
        bits = None
        if out_token is not None and ps == 1 and p.Cout % 16 == 0 and _is_cl(y) and \
                _uses_tensor_path(p, _lib.PASS_FPROP, _is_cl(x), True):
            bits = torch.empty((p.N, ho.value, wo.value, p.Cout // 16), dtype=torch.int16, device=x.device)
            out_token.bits = bits
        ws = _workspace(x.device, _ws_bytes(p, _lib.PASS_FPROP))
        tx, ty = t4(x), t4(y)
        tr = t4(residual) if residual is not None else None
        tp = t4(preact) if preact is not None else None
        check(lib.srb_conv_fprop(ctypes.byref(p), ctypes.byref(tx), _ptr(weight), _ptr(bias), _ptr(alpha),
                                 ctypes.byref(tr) if tr is not None else None, ctypes.byref(ty),
                                 ctypes.byref(tp) if tp is not None else None, _ptr(bits),
                                 _ptr(ws), ws.numel(), _stream(x.device)))
        ctx.p = p
        ctx.act = act
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        ctx.params = (weight, bias)  # GradBucket direct-write targets (ddp.py)
        ctx.in_token, ctx.out_token = in_token, out_token
        ctx.skip_sink, ctx.skip_src = skip_sink, skip_src
        if skip_sink is not None:  # this conv's backward will compute dL/dx: the skip's gradient can ride in its epilogue
            skip_sink.armed = bool(ctx.needs_input_grad[0])
            skip_sink.grad = None
        ctx.save_for_backward(x, weight, alpha, preact if need_preact else (y if act is not None else None))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, alpha, ref = ctx.saved_tensors
        p = ctx.p
        dev = x.device
        st = _stream(dev)
        tok = ctx.out_token
        premasked = tok is not None and tok.premasked == (dy.data_ptr(), dy._version, tuple(dy.shape), tuple(dy.stride()))
        if tok is not None:
            tok.premasked = None
        dy = _as_act(dy)
        dres = dy if (ctx.has_res and ctx.needs_input_grad[4]) else None
        if dres is not None and ctx.skip_src is not None and ctx.skip_src.armed:
            ctx.skip_src.grad = dres  # the block's first conv adds it to its dgrad (SkipGrad)
            dres = None
        dalpha = None
        if premasked:
            dz = dy  # the consumer's dgrad epilogue already applied this layer's ReLU mask (and the tf32 rounding)
        elif ctx.act is not None:
            dz = torch.empty(dy.shape, dtype=dy.dtype, device=dev, memory_format=_out_format(p.Cout))
            if ctx.act == "prelu":
                dalpha = torch.zeros_like(alpha)
            tdy, tref, tdz = t4(dy), t4(ref), t4(dz)
            check(lib.srb_act_bwd(ctypes.byref(p), ctypes.byref(tdy), ctypes.byref(tref), _ptr(alpha),
                                  ctypes.byref(tdz), _ptr(dalpha), st))
        else:
            dz = dy
        if p.ps > 1 and _folds_ps(p, _is_cl(x), _is_cl(dz)):
            pass  # dgrad / wgrad read dz in y's layout: the un-shuffle is their TMA traversal (no extra pass over dz)
        elif p.ps > 1 and p.math != _lib.MATH_FP32 and x.shape[1] % 4 == 0 and x.shape[1] >= 8 and _is_cl(x):
            # PixelShuffle layer on the tensor path: undo the shuffle once (NHWC, tf32) and run dgrad/wgrad as a
            # plain conv with Cout*r*r output channels
            r = p.ps
            dzu = torch.empty((p.N, p.Cout * r * r, dz.shape[2] // r, dz.shape[3] // r), dtype=dz.dtype,
                              device=dev, memory_format=torch.channels_last)
            tdz0, tdzu = t4(dz), t4(dzu)
            check(lib.srb_pixel_unshuffle(ctypes.byref(p), ctypes.byref(tdz0), ctypes.byref(tdzu), st))
            dzu = _as_act(dzu)  # bf16 mode with an fp32 (3-channel) shuffled output: the un-shuffled gradient is stored bf16
            p = ConvParams(p.N, p.Cin, p.H, p.W, p.Cout * r * r, p.kh, p.kw, p.stride, p.pad, 0, 0, 1, p.act,
                           p.slope, p.math)
            dz = dzu
        tdz, tx = t4(dz), t4(x)
        dw = db = dx = None
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            wparam, bparam = ctx.params
            direct = getattr(wparam, "_srb_direct", False) and wparam.grad is not None and \
                (bparam is None or (getattr(bparam, "_srb_direct", False) and bparam.grad is not None))
            accumulate = 0
            if direct:
                # the wgrad kernel writes the parameter's slot of the flat gradient bucket, pre-scaled: the first write of a
                # step overwrites (no zero_grad pass), later ones (a module applied twice before one backward, e.g.
                # srgan.py:275-286, or micro-batch accumulation without begin_step) add -- GradBucket tracks `_srb_written`
                dw_t, db_t, scale = wparam.grad, (bparam.grad if bparam is not None else None), _state["grad_scale"]
                accumulate = 1 if getattr(wparam, "_srb_written", False) else 0
                wparam._srb_written = True
                if bparam is not None:
                    bparam._srb_written = True
            else:
                dw_t = dw = torch.empty_like(weight)
                db_t = db = torch.empty(weight.shape[1] if p.transposed else weight.shape[0], dtype=torch.float32,
                                        device=dev) if ctx.has_bias else None
                scale = 1.0
            ws = _workspace(dev, _ws_bytes(p, _lib.PASS_WGRAD))
            check(lib.srb_conv_wgrad(ctypes.byref(p), ctypes.byref(tx), ctypes.byref(tdz), _ptr(dw_t), _ptr(db_t),
                                     ctypes.c_float(scale), accumulate, _ptr(ws), ws.numel(), st))
        if ctx.needs_input_grad[0]:
            dx = torch.empty(x.shape, dtype=x.dtype, device=dev,
                             memory_format=torch.channels_last if _is_cl(x) else torch.contiguous_format)
            tdx = t4(dx)
            itok = ctx.in_token
            fuse = itok is not None and itok.consumers == 1 and not itok.observed and _state["fuse_relu_bwd"]
            # x = ReLU(z_prev): its sign pattern is the producer's mask -- in packed form when the producer's fprop wrote
            # it and this dgrad runs on the tensor path, else read from x itself
            bits = None
            if fuse and itok.bits is not None and p.Cin % 16 == 0 and \
                    _uses_tensor_path(p, _lib.PASS_DGRAD, _is_cl(dx), _is_cl(dz)):
                bits = itok.bits
            if fuse and bits is None and p.math == _lib.MATH_BF16:
                fuse = False  # bf16 storage mode folds the mask only in its packed form
            tmask = t4(x) if (fuse and bits is None) else None
            ws = _workspace(dev, _ws_bytes(p, _lib.PASS_DGRAD))
            skip = None
            if ctx.skip_sink is not None and ctx.skip_sink.grad is not None:
                skip, ctx.skip_sink.grad = ctx.skip_sink.grad, None
            if skip is not None and tuple(skip.shape) == tuple(dx.shape) and skip.dtype == dx.dtype:
                tadd = t4(skip)
                rc = lib.srb_conv_dgrad_add(ctypes.byref(p), ctypes.byref(tdz), _ptr(weight),
                                            ctypes.byref(tmask) if tmask is not None else None, _ptr(bits),
                                            ctypes.byref(tadd), ctypes.byref(tdx), _ptr(ws), ws.numel(), st)
                if rc == _lib.EUNSUPPORTED:  # nothing was launched: plain dgrad (no mask folded in), then the sum
                    fuse = False
                    check(lib.srb_conv_dgrad(ctypes.byref(p), ctypes.byref(tdz), _ptr(weight), None, None, ctypes.byref(tdx),
                                             _ptr(ws), ws.numel(), st))
                    dx = dx + skip
                else:
                    check(rc)
            else:
                if skip is not None:
                    fuse, tmask, bits = False, None, None
                check(lib.srb_conv_dgrad(ctypes.byref(p), ctypes.byref(tdz), _ptr(weight),
                                         ctypes.byref(tmask) if tmask is not None else None, _ptr(bits), ctypes.byref(tdx),
                                         _ptr(ws), ws.numel(), st))
                if skip is not None:
                    dx = dx + skip.to(dx.dtype)
            if fuse:
                itok.premasked = (dx.data_ptr(), dx._version, tuple(dx.shape), tuple(dx.stride()))
        elif ctx.skip_sink is not None and ctx.skip_sink.grad is not None:
            ctx.skip_sink.grad = None  # unreachable by construction (armed == needs_input_grad[0]); never leak a stale gradient
        return dx, dw, db, dalpha, dres, None, None, None, None, None, None, None, None, None, None, None


def _apply(x, weight, bias, alpha, residual, stride, pad, out_pad, transposed, ps, act, slope, skip_sink=None, skip_src=None):
    """_FusedConv.apply plus the ReLU-backward handshake (see _ReluToken)."""
    in_token = getattr(x, "_srb_relu", None) if x.requires_grad else None
    if in_token is not None:
        in_token.consumers += 1
        # the fused dgrad hands the producer dL/dz (already masked) in place of dL/dy: anything that observes x.grad
        # (retain_grad, tensor hooks registered before this call) must see the un-masked gradient, so do not fuse
        if x.retains_grad or getattr(x, "_backward_hooks", None):
            in_token.observed = True
    # y = ReLU(z) exactly (no residual on top); only worth it when a gradient will flow back into this layer
    out_token = _ReluToken() if (act == "relu" and residual is None and torch.is_grad_enabled()) else None
    y = _FusedConv.apply(x, weight, bias, alpha, residual, stride, pad, out_pad, transposed, ps, act, slope, in_token,
                         out_token, skip_sink, skip_src)
    if out_token is not None and y.requires_grad:
        y._srb_relu = out_token
    if act is not None and _state["act_recorder"] is not None:
        _record(y, residual)
    return y


def conv2d(x, weight, bias=None, stride=1, padding=0, activation=None, alpha=None, slope=0.2, residual=None,
           pixel_shuffle=1, skip_sink=None, skip_src=None):
    """Fused Conv2d -> +bias -> act -> PixelShuffle(r) -> +residual.

    activation: None | 'relu' | 'prelu' (alpha = the nn.PReLU weight, shape (1,)) | 'lrelu' (slope).
    skip_sink / skip_src: one shared SkipGrad for the first (sink: its input x is also the block's skip) and the last (src: its
    `residual` is that x) conv of a residual block -- the skip's gradient is then added inside the first conv's dgrad kernel.
    """
    if activation == "prelu":
        assert alpha is not None and alpha.numel() == 1, "base_networks.py uses nn.PReLU() with one shared slope"
    return _apply(x, weight, bias, alpha if activation == "prelu" else None, residual,
                  int(stride), int(padding), 0, False, int(pixel_shuffle), activation, slope, skip_sink, skip_src)


def conv_transpose2d(x, weight, bias=None, stride=1, padding=0, output_padding=0, activation=None, alpha=None,
                     slope=0.2, residual=None):
    """Fused ConvTranspose2d -> +bias -> act -> +residual (weight is (Cin, Cout, kh, kw) like torch)."""
    return _apply(x, weight, bias, alpha if activation == "prelu" else None, residual,
                  int(stride), int(padding), int(output_padding), True, 1, activation, slope)


class _PReLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, alpha):
        _require_cuda(x, alpha)
        x = _dense(x)
        y = torch.empty_like(x)
        check(lib.srb_prelu_fwd(_ptr(x), _ptr(alpha), _ptr(y), x.numel(), _stream(x.device)))
        ctx.save_for_backward(x, alpha)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, alpha = ctx.saved_tensors
        if _is_cl(x):
            dy = dy.contiguous(memory_format=torch.channels_last)
        else:
            dy = dy.contiguous()
        dx = torch.empty_like(x)
        dalpha = torch.zeros_like(alpha)
        check(lib.srb_prelu_bwd(_ptr(x), _ptr(dy), _ptr(alpha), _ptr(dx), _ptr(dalpha), x.numel(),
                                _stream(x.device)))
        return dx, dalpha


def prelu(x, alpha):
    """Stand-alone PReLU with one shared slope (fsrcnn.py:26)."""
    assert alpha.numel() == 1
    y = _PReLU.apply(x, alpha)
    if _state["act_recorder"] is not None:
        _record(y)
    return y


class _Loss(torch.autograd.Function):
    """mean((y-t)^2) / mean(|y-t|) with a one-pass backward (srb_loss_fwd / srb_loss_bwd)."""

    @staticmethod
    def forward(ctx, y, t, kind):
        _require_cuda(y, t)
        assert y.shape == t.shape, "loss operands must have the same shape"
        y = _dense(y)
        if t.stride() != y.stride():  # same dense layout on both sides (layout plumbing, not arithmetic)
            t = t.contiguous(memory_format=torch.channels_last) if (y.dim() == 4 and _is_cl(y)) else t.contiguous()
            if t.stride() != y.stride():
                y = y.contiguous()
                t = t.contiguous()
        loss = torch.empty((), dtype=torch.float32, device=y.device)
        ws = _workspace(y.device, int(lib.srb_loss_workspace_bytes()))
        check(lib.srb_loss_fwd(kind, _ptr(y), _ptr(t), y.numel(), _ptr(loss), _ptr(ws), ws.numel(), _stream(y.device)))
        ctx.kind = kind
        ctx.save_for_backward(y, t)
        return loss

    @staticmethod
    def backward(ctx, g):
        y, t = ctx.saved_tensors
        g = g.contiguous()
        dy = torch.empty_like(y)
        check(lib.srb_loss_bwd(ctx.kind, _ptr(y), _ptr(t), y.numel(), _ptr(g), _ptr(dy), _stream(y.device)))
        return dy, None, None


def mse_loss(y, t):
    """nn.MSELoss() (mean) as two fused kernels forward and one backward (srcnn.py:84, espcn.py:84, vdsr.py:96)."""
    return _Loss.apply(y, t, 0)


def l1_loss(y, t):
    """nn.L1Loss() (mean) (edsr.py:98)."""
    return _Loss.apply(y, t, 1)


def image_to_tensor(img_u8, out=None):
    """torchvision ToTensor on the device (dataset.py:90): uint8 (N,H,W,C) image batch in [0,255] -> fp32 (N,C,H,W) in [0,1].
    The host ships 1-byte pixels; `out` (a resident fp32 NCHW buffer, e.g. a CUDA-graph input slot) is filled in place."""
    if not img_u8.is_cuda or img_u8.dtype != torch.uint8 or img_u8.dim() != 4:
        raise RuntimeError("image_to_tensor needs a CUDA uint8 (N,H,W,C) tensor")
    img_u8 = img_u8.contiguous()
    n, h, w, c = img_u8.shape
    if out is None:
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=img_u8.device)
    assert out.is_contiguous() and tuple(out.shape) == (n, c, h, w) and out.dtype == torch.float32
    check(lib.srb_image_to_tensor(_ptr(img_u8), _ptr(out), n, h, w, c, ctypes.c_float(1.0 / 255.0), _stream(img_u8.device)))
    return out


class _FusedConvLoss(torch.autograd.Function):
    """loss = mean((y-t)^2) or mean(|y-t|) with y = PixelShuffle_ps(conv(x, w) + b), in ONE kernel (srb_conv_fprop_loss): the
    epilogue of the last conv compares with the target, accumulates the loss and emits d loss / d y -- for PixelShuffle(4)
    layers already un-shuffled into the conv's own NHWC layout.  Backward = wgrad + dgrad on that saved gradient."""

    @staticmethod
    def forward(ctx, x, weight, bias, target, kind, stride, pad, ps, need_output):
        _require_cuda(weight, bias, None if target.dtype == torch.uint8 else target)
        if not target.is_cuda:
            raise RuntimeError("srb200 kernels need CUDA tensors (target on %s); there is no CPU path" % target.device)
        x = _as_act(x)
        weight = weight.contiguous()
        target = target.contiguous()
        p = _params(x, weight, stride, pad, 0, False, ps, None, 0.2)
        ho, wo = ctypes.c_int32(), ctypes.c_int32()
        check(lib.srb_conv_out_hw(ctypes.byref(p), ctypes.byref(ho), ctypes.byref(wo)))
        oshape = (p.N, p.Cout, ho.value * ps, wo.value * ps)
        if target.dtype == torch.uint8:
            # the decoded image batch itself, (N,H,W,C) bytes: the epilogue reads t = byte/255 (ToTensor) through this NCHW view
            assert tuple(target.shape) == (oshape[0], oshape[2], oshape[3], oshape[1]), "uint8 target must be the (N,H,W,C) image"
            target = target.permute(0, 3, 1, 2)
        assert tuple(target.shape) == oshape, "target must have the shape of the network output"
        dev = x.device
        y = torch.empty(oshape, dtype=torch.float32, device=dev) if need_output else None
        unshuf = ps > 1
        if unshuf:
            dz = torch.empty((p.N, p.Cout * ps * ps, ho.value, wo.value), dtype=torch.float32, device=dev,
                             memory_format=torch.channels_last)
        else:
            dz = torch.empty(oshape, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        ws = _workspace(dev, _ws_bytes(p, _lib.PASS_FPROP))
        tx, tt, tdz = t4(x), t4(target), t4(dz)
        ty = t4(y) if y is not None else None
        check(lib.srb_conv_fprop_loss(ctypes.byref(p), ctypes.byref(tx), _ptr(weight), _ptr(bias), ctypes.byref(tt), kind,
                                      ctypes.byref(ty) if ty is not None else None, ctypes.byref(tdz), 1 if unshuf else 0,
                                      _ptr(loss), _ptr(ws), ws.numel(), _stream(dev)))
        # geometry the backward kernels see: the plain conv with Cout*ps*ps output channels
        ctx.pb = ConvParams(p.N, p.Cin, p.H, p.W, p.Cout * ps * ps, p.kh, p.kw, p.stride, p.pad, 0, 0, 1, _lib.ACT_NONE,
                            p.slope, p.math) if unshuf else p
        ctx.has_bias = bias is not None
        ctx.params = (weight, bias)
        ctx.in_token = getattr(x, "_srb_relu", None) if x.requires_grad else None
        if ctx.in_token is not None:
            ctx.in_token.consumers += 1
        ctx.save_for_backward(x, weight, dz)
        ctx.set_materialize_grads(False)
        if y is None:
            return loss
        ctx.mark_non_differentiable(y)
        return loss, y

    @staticmethod
    def backward(ctx, gloss, *unused):
        x, weight, dz = ctx.saved_tensors
        p = ctx.pb
        dev = x.device
        st = _stream(dev)
        if gloss is None:
            return (None,) * 9
        # d loss_total / d z = gloss * saved gradient (gloss == 1 for `loss.backward()`: the kernel then returns at once)
        check(lib.srb_scale_by_scalar(_ptr(dz), dz.numel(), _ptr(gloss.contiguous().float()),
                                      1 if p.math in (_lib.MATH_AUTO, _lib.MATH_TF32) else 0, st))
        if p.math == _lib.MATH_BF16:
            dz = _as_act(dz)
        tdz, tx = t4(dz), t4(x)
        dw = db = dx = None
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            wparam, bparam = ctx.params
            direct = getattr(wparam, "_srb_direct", False) and wparam.grad is not None and \
                (bparam is None or (getattr(bparam, "_srb_direct", False) and bparam.grad is not None))
            accumulate = 0
            if direct:
                dw_t, db_t, scale = wparam.grad, (bparam.grad if bparam is not None else None), _state["grad_scale"]
                accumulate = 1 if getattr(wparam, "_srb_written", False) else 0
                wparam._srb_written = True
                if bparam is not None:
                    bparam._srb_written = True
            else:
                dw_t = dw = torch.empty_like(weight)
                db_t = db = torch.empty(weight.shape[0], dtype=torch.float32, device=dev) if ctx.has_bias else None
                scale = 1.0
            ws = _workspace(dev, _ws_bytes(p, _lib.PASS_WGRAD))
            check(lib.srb_conv_wgrad(ctypes.byref(p), ctypes.byref(tx), ctypes.byref(tdz), _ptr(dw_t), _ptr(db_t),
                                     ctypes.c_float(scale), accumulate, _ptr(ws), ws.numel(), st))
        if ctx.needs_input_grad[0]:
            dx = torch.empty(x.shape, dtype=x.dtype, device=dev,
                             memory_format=torch.channels_last if _is_cl(x) else torch.contiguous_format)
            tdx = t4(dx)
            itok = ctx.in_token
            fuse = itok is not None and itok.consumers == 1 and not itok.observed and _state["fuse_relu_bwd"]
            bits = None
            if fuse and itok.bits is not None and p.Cin % 16 == 0 and \
                    _uses_tensor_path(p, _lib.PASS_DGRAD, _is_cl(dx), _is_cl(dz)):
                bits = itok.bits
            if fuse and bits is None and p.math == _lib.MATH_BF16:
                fuse = False
            tmask = t4(x) if (fuse and bits is None) else None
            ws = _workspace(dev, _ws_bytes(p, _lib.PASS_DGRAD))
            check(lib.srb_conv_dgrad(ctypes.byref(p), ctypes.byref(tdz), _ptr(weight),
                                     ctypes.byref(tmask) if tmask is not None else None, _ptr(bits), ctypes.byref(tdx),
                                     _ptr(ws), ws.numel(), st))
            if fuse:
                itok.premasked = (dx.data_ptr(), dx._version, tuple(dx.shape), tuple(dx.stride()))
        return dx, dw, db, None, None, None, None, None, None


_LOSS_KINDS = {"mse": 0, "l2": 0, "l1": 1}


def loss_fusable(x_shape, x_cl, weight_shape, stride, padding, pixel_shuffle):
    """Can conv2d_loss run this last layer?  (Conv2d on the tensor path, stride 1; PixelShuffle 1 or 4.)"""
    if stride != 1 or pixel_shuffle not in (1, 4) or _state["math"] in (_lib.MATH_FP32, _lib.MATH_EXACT):
        return False
    if pixel_shuffle == 4 and weight_shape[0] % 16 != 0:
        return False
    N, Cin, H, W = x_shape
    p = ConvParams(N, Cin, H, W, weight_shape[0] // (pixel_shuffle ** 2), weight_shape[2], weight_shape[3], stride, padding, 0,
                   0, pixel_shuffle, _lib.ACT_NONE, 0.2, _state["math"])
    return _uses_tensor_path(p, _lib.PASS_FPROP, x_cl, False)


def conv2d_loss(x, weight, bias, target, kind="mse", stride=1, padding=0, pixel_shuffle=1, need_output=False):
    """Last conv of a network fused with nn.MSELoss / nn.L1Loss (mean): returns loss, or (loss, y) with need_output=True
    (y is then a plain non-differentiable tensor: the gradient flows through the loss)."""
    return _FusedConvLoss.apply(x, weight, bias, target, _LOSS_KINDS[kind], int(stride), int(padding), int(pixel_shuffle),
                                bool(need_output))
