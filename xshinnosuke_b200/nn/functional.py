"""Functional ops (mirror of the hot-path subset of xs/nn/functional.py).

Same signatures, defaults and side effects as the reference: each op (i) computes -- here by one or
two libxsb200 kernel launches on the current CUDA stream, (ii) wires the dynamic graph
(`add_in_bounds` / `add_out_bounds`) when `out.is_dynamic`, (iii) stashes what its backward needs in
`out.cache`, (iv) sets `out.grad_fn` and gives grad-requiring operands a zeroed grad
(`initialize_ops_grad`). `out=` is honoured (the static-graph front end passes preallocated buffers).
Ops outside the hot path named in SURVEY.md section 8a are not provided.
"""
import ctypes

import numpy as np

from ..core import state as GLOBAL
from ..core.device import DevArray, rt
from ..core.tensor import Parameter, Tensor
from ..utils.common import initialize_ops_grad, pair
from .._lib import lib
from . import x3
from .grad_fn import (AddBackward, AddmmBackward, Avgpool2DBackward, BatchNormBackward, Conv2DBackward,
                      CrossEntropyLossBackward, Dropout2DBackward, FlattenBackward, GroupNormBackward, LayerNormBackward,
                      Maxpool2DBackward, MeanBackward, MMBackward, MSELossBackward, Pad2DBackward, ReLUBackward,
                      SigmoidBackward, SumBackward, TanhBackward, ViewBackward)


def _mark_inputs(t):
    if GLOBAL.INPUTS is None:
        GLOBAL.INPUTS = t


def _out_arr(out, shape, dtype="float32"):
    """Resolve the destination DevArray: a fresh one, or the (active prefix of the) preallocated `out`."""
    if out is None or out._arr is None:
        arr = DevArray.empty(shape, dtype)
        if out is None:
            return Tensor(arr), arr
        out._arr = arr
        return out, arr
    arr = out.arr
    if tuple(arr.shape) != tuple(shape):
        raise ValueError(f"out has shape {arr.shape}, op produces {tuple(shape)}")
    return out, arr


def _wire(out, operands, requires_grad, grad_fn, cache=None, grad_ops=None):
    out.requires_grad = bool(requires_grad)
    if out.is_dynamic:
        out.add_in_bounds(*operands)
        for o in operands:
            if o is not None:
                o.add_out_bounds(out)
    else:
        # static (preallocated) output: its closure must see THIS step's operands -- the backward
        # kernels read the live input tensor where the reference reads a cached copy (col / xhat)
        out._in_bounds = list(operands)
    if GLOBAL.COMPUTE_GRAD and out.requires_grad:
        if cache:
            out.cache.update(cache)
        out.grad_fn = grad_fn
        ops = operands if grad_ops is None else grad_ops
        initialize_ops_grad(*ops)
        for o in ops:
            if isinstance(o, Parameter) and o.requires_grad:
                o.pending_uses += 1
    return out


class DeviceArgmax:
    """`cache['pool_argmax']`: the max-pool arg-max kept on the device as one byte per output
    (k*k <= 255). np.asarray(...) yields the reference's representation -- int64, flat, (n,oh,ow,c)
    order (xs/nn/functional.py:475)."""

    def __init__(self, dev, count):
        self.dev, self.count = dev, count

    def __array__(self, dtype=None, copy=None):
        a = self.dev[: self.count].cpu().numpy().astype(np.int64)
        return a if dtype is None else a.astype(dtype)

    @property
    def size(self):
        return self.count


# ---------------------------------------------------------------------------------------------- math
def add(lhs: Tensor, rhs: Tensor, out: Tensor = None) -> Tensor:
    """xs/nn/functional.py:6-21 (same-shape operands: the residual add)."""
    _mark_inputs(lhs)
    a, b = lhs.arr, rhs.arr
    if tuple(a.shape) != tuple(b.shape):
        return _add_rowvector(lhs, rhs, out)
    if a.cl != b.cl:
        b = b.as_cl() if a.cl else b.logical_rowmajor()
    out, y = _out_arr(out, a.shape)
    y.cl = a.cl
    ybuf, n = y.buf, y.size
    y.wptr  # noqa: B018 -- clears stale pending state of a reused (static) buffer

    def run(mask=None, _a=a, _b=b):
        # deferred: a following ReLU(inplace=True) turns this into ONE add+ReLU(+mask) pass
        if mask is None:
            lib.xsb_add(_a.ptr, _b.ptr, ybuf.data_ptr(), n, rt.stream())
        else:
            if rt.fuse_bn_add_relu and _bn_add_relu(_a, _b, ybuf, mask):
                return
            lib.xsb_add_relu(_a.ptr, _b.ptr, ybuf.data_ptr(), mask.data_ptr(), n, rt.stream())
        rt.launches += 1
    run.fusable_relu = n % 8 == 0
    y.pending_op = run
    return _wire(out, (lhs, rhs), lhs.requires_grad or rhs.requires_grad, AddBackward)


def _bn_add_relu(a, b, ybuf, mask):
    """The tail of a residual block: an operand of the add is a BatchNorm output whose apply pass is still parked (F.batch_norm
    defers it by one op) -- z = max(0, bn(x) + other) and the mask leave from ONE kernel that reads the BatchNorm INPUT; both
    operands may be such outputs (down-sampling shortcut). The parked apply passes stay parked: whoever reads a BatchNorm
    output later still gets it from the plain kernel. -> True when the fused kernel ran."""
    pa, pb = getattr(a.pending_op, "bn", None), getattr(b.pending_op, "bn", None)
    if pa is None and pb is None:
        return False
    if pa is None:                       # (fp32 addition commutes exactly)
        a, b, pa, pb = b, a, pb, None
    R, C = pa["R"], pa["C"]
    if pa["x"].cl != a.cl or a.cl != b.cl or a.size != R * C:
        return False
    if pb is not None and (pb["R"] != R or pb["C"] != C or pb["x"].cl != b.cl):
        pb = None
    bsrc = pb["x"].ptr if pb is not None else b.ptr
    rc = lib.xsb_bn_add_relu(pa["x"].ptr, pa["mean"].peek_ptr(), pa["invstd"].peek_ptr(), pa["gamma"].ptr, pa["beta"].ptr, bsrc,
                             pb["mean"].peek_ptr() if pb is not None else None,
                             pb["invstd"].peek_ptr() if pb is not None else None,
                             pb["gamma"].ptr if pb is not None else None, pb["beta"].ptr if pb is not None else None,
                             ybuf.data_ptr(), mask.data_ptr(), R, C, rt.stream())
    if rc != 0:
        return False
    rt.launches += 1
    return True


def _add_rowvector(lhs, rhs, out):
    """lhs [..., N] + rhs broadcast along the leading axes (rhs = [1, N] / [N] / [1, C, H, W]: numpy's `add(lhs, rhs,
    out=empty_like(lhs))`, xs/nn/functional.py:6-21). The result has lhs's shape, in logical row-major order."""
    a = lhs.arr.logical_rowmajor()
    b = rhs.arr.logical_rowmajor()
    tail = tuple(s for s in b.shape)
    while tail and tail[0] == 1:
        tail = tail[1:]
    if not tail or tuple(a.shape[len(a.shape) - len(tail):]) != tail:
        raise ValueError(f"add: operands could not be broadcast together with shapes {a.shape} {b.shape}")
    cols = b.size
    rows = a.size // cols
    tmp = DevArray.empty(a.shape)
    tmp.cl = False
    lib.xsb_broadcast_axis(b.ptr, tmp.wptr, 1, rows, cols, 1.0, 0, rt.stream())
    out, y = _out_arr(out, a.shape)
    ycl = y.cl
    y.cl = False
    lib.xsb_add(a.ptr, tmp.ptr, y.wptr, a.size, rt.stream())
    rt.launches += 2
    if ycl and len(a.shape) == 4:        # a preallocated 4-D buffer keeps channels-last storage
        conv = y.as_cl()
        if conv is not y:
            y.buf[: y.size].copy_(conv.buf[: y.size])
        y.cl = True
    return _wire(out, (lhs, rhs), lhs.requires_grad or rhs.requires_grad, AddBackward, cache={"broadcast": (rows, cols)})


def _gemm_fwd(x, w, bias, y):
    """y = x . w (+ bias), deferred by one op: a ReLU that consumes y (Dense(..., activation='relu'), xs/layers/linear.py:
    43-60) asks the producer for the fused bias+ReLU epilogue -- y AND max(0, y) leave from one kernel."""
    m, k = x.shape
    n = w.shape[-1]
    ybuf = y.buf
    y.wptr  # noqa: B018 -- clears stale pending state of a reused (static) buffer

    def run(mask=None, act_out=None, _x=x, _w=w, _b=bias):
        bptr = _b.ptr if _b is not None else None
        split = rt.math_mode == x3.MATH_FP32X3 and x3.gemm_ok(k, n)
        if act_out is not None and not split:
            lib.xsb_gemm_bias_relu(_x.ptr, _w.ptr, ybuf.data_ptr(), bptr, act_out, m, n, k, 0, 0, x3.c_mode(False),
                                   rt.ws_big.data_ptr(), rt.WS_BIG_FLOATS, rt.stream())
            rt.launches += 1
            return True
        if split:
            x3.gemm(_x.ptr, m * k, _w.ptr, k * n, ybuf.data_ptr(), bptr, m, n, k, 0, 0, 0)
        else:
            lib.xsb_gemm(_x.ptr, _w.ptr, ybuf.data_ptr(), bptr, m, n, k, 0, 0, 0, x3.c_mode(False), rt.ws_big.data_ptr(),
                         rt.WS_BIG_FLOATS, rt.stream())
            rt.launches += 1
        return False
    run.fusable_relu = False
    run.fusable_act = True
    y.pending_op = run


def _as_matrix(t):
    a = t.arr
    if a.ndim != 2:
        raise NotImplementedError(f"matrix operand must be 2-D, got {a.shape}")
    return a


def mm(lhs: Tensor, rhs: Tensor, out: Tensor = None) -> Tensor:
    """xs/nn/functional.py:81-96: out = lhs . rhs."""
    _mark_inputs(lhs)
    x, w = _as_matrix(lhs), _as_matrix(rhs)
    out, y = _out_arr(out, (x.shape[0], w.shape[-1]))
    _gemm_fwd(x, w, None, y)
    return _wire(out, (lhs, rhs), lhs.requires_grad or rhs.requires_grad, MMBackward)


def addmm(b: Tensor, x1: Tensor, x2: Tensor, out: Tensor = None) -> Tensor:
    """xs/nn/functional.py:99-120: out = x1 . x2 + b (Dense; x2 is [in, out], b is [1, out])."""
    _mark_inputs(x1)
    x, w = _as_matrix(x1), _as_matrix(x2)
    if b.arr.size != w.shape[-1]:
        raise ValueError("bias size mismatch")
    out, y = _out_arr(out, (x.shape[0], w.shape[-1]))
    _gemm_fwd(x, w, b.arr, y)
    return _wire(out, (x1, x2, b), x1.requires_grad or x2.requires_grad or b.requires_grad, AddmmBackward,
                 grad_ops=(b, x1, x2))


def _reduce(x, axis, keepdims, op):
    _mark_inputs(x)
    a = x.arr.logical_rowmajor()
    shape = a.shape
    if axis is None:
        outer, red, inner = 1, a.size, 1
        oshape = (1,) * len(shape) if keepdims else (1,)
    else:
        axis = axis % len(shape)
        outer = int(np.prod(shape[:axis], dtype=np.int64))
        red = shape[axis]
        inner = int(np.prod(shape[axis + 1:], dtype=np.int64))
        oshape = shape[:axis] + ((1,) if keepdims else ()) + shape[axis + 1:]
        if not oshape:
            oshape = (1,)
    y = DevArray.empty(oshape)
    y.cl = False
    lib.xsb_reduce_axis(a.ptr, y.wptr, outer, red, inner, 1.0 / red if op == "mean" else 1.0, rt.stream())
    rt.launches += 1
    out = Tensor(y)
    return _wire(out, (x,), x.requires_grad, MeanBackward if op == "mean" else SumBackward,
                 cache={"axis": axis, "ori": (outer, red, inner), "op": op})


def mean(x: Tensor, axis: int = None, keepdims: bool = False, out: Tensor = None) -> Tensor:
    """xs/nn/functional.py:244-260."""
    return _reduce(x, axis, keepdims, "mean")


def sum(x: Tensor, axis: int = None, keepdims: bool = False, out: Tensor = None) -> Tensor:
    """xs/nn/functional.py:224-241."""
    return _reduce(x, axis, keepdims, "sum")


# ---------------------------------------------------------------------------------------------- activations
def relu(inputs: Tensor, inplace: bool = False, out: Tensor = None, _mask=None) -> Tensor:
    """xs/nn/functional.py:297-313. `_mask` (private): device byte buffer that receives the x<0 bit
    mask in the same pass -- used by layers.ReLU(inplace=True)."""
    _mark_inputs(inputs)
    x = inputs.arr
    mptr = _mask.data_ptr() if _mask is not None else None
    if inplace:
        lib.xsb_relu_fwd(x.ptr, x.peek_ptr(), mptr, x.size, rt.stream())
        rt.launches += 1
        return inputs
    out, y = _out_arr(out, x.shape)
    y.cl = x.cl
    producer = x.pending_op
    fused = False
    if producer is not None and getattr(producer, "fusable_act", False) and _mask is None:
        # the producer (Dense / Conv2D with activation='relu') has not run yet: its epilogue writes x and max(0, x)
        x.pending_op = None
        fused = producer(act_out=y.wptr)
    if not fused:
        lib.xsb_relu_fwd(x.ptr, y.wptr, mptr, x.size, rt.stream())
        rt.launches += 1
    return _wire(out, (inputs,), inputs.requires_grad, ReLUBackward)


def softmax(inputs: Tensor, out: Tensor = None) -> Tensor:
    """xs/nn/functional.py:348-357 (host-side helper; the training path uses the fused cross_entropy)."""
    z = inputs.eval.astype(np.float64)
    e = np.exp(z - z.max(axis=-1, keepdims=True))
    res = Tensor((e / e.sum(axis=-1, keepdims=True)).astype(np.float32))
    res.requires_grad = inputs.requires_grad
    return res


# ---------------------------------------------------------------------------------------------- shape
def flatten(inputs: Tensor, start: int = 1, out: Tensor = None) -> Tensor:
    """xs/nn/functional.py:370-384. The result is in the reference's NCHW feature order: the
    channels-last activation is transposed back once here (zero-copy when H*W == 1)."""
    _mark_inputs(inputs)
    a = inputs.arr.logical_rowmajor()
    shape = tuple(a.shape[:start]) + (int(np.prod(a.shape[start:], dtype=np.int64)),)
    flat = a.reshaped(shape)
    if out is None:
        out = Tensor(flat)
    else:
        dst = out.arr
        if dst is None:
            out._arr = flat
        else:
            flat.flush()
            dst.buf[: flat.size].copy_(flat.buf[: flat.size])
            dst.pending_zero = False
    return _wire(out, (inputs,), inputs.requires_grad, FlattenBackward)


def view(inputs: Tensor, shape, out: Tensor = None) -> Tensor:
    """xs/nn/functional.py:814-828."""
    _mark_inputs(inputs)
    a = inputs.arr.logical_rowmajor()
    shape = tuple(int(s) for s in shape)
    if -1 in shape:
        known = -int(np.prod(shape, dtype=np.int64))
        shape = tuple(a.size // known if s == -1 else s for s in shape)
    v = a.reshaped(shape)
    if out is None:
        out = Tensor(v)
    else:
        dst = out.arr
        if dst is None:
            out._arr = v
        else:
            v.flush()
            dst.buf[: v.size].copy_(v.as_cl().buf[: v.size] if dst.cl else v.buf[: v.size])
            dst.pending_zero = False
    return _wire(out, (inputs,), inputs.requires_grad, ViewBackward)


# ---------------------------------------------------------------------------------------------- conv / pool
def conv2d(inputs: Tensor, weight: Tensor, bias: Tensor = None, stride=(1, 1), padding: int = 0,
           out: Tensor = None) -> Tensor:
    """xs/nn/functional.py:405-442. Implicit GEMM over channels-last data (xsb_conv2d_fwd);
    OH = (H - kh + 2p)//s + 1. Bias is fused into the kernel epilogue."""
    _mark_inputs(inputs)
    x = inputs.arr.as_cl()
    w = weight.arr.as_cl()
    n, c, h, wd = x.shape
    oc, wc, kh, kw = w.shape
    if wc != c:
        raise ValueError(f"conv2d: input has {c} channels, weight expects {wc}")
    sh, sw = pair(stride)
    oh = (h - kh + 2 * padding) // sh + 1
    ow = (wd - kw + 2 * padding) // sw + 1
    out, y = _out_arr(out, (n, oc, oh, ow))
    y.cl = True
    ybuf = y.buf
    y.wptr  # noqa: B018 -- clears stale pending state of a reused (static) buffer
    barr = bias.arr if bias is not None else None

    def run(mask=None, stats=False, act_out=None, _x=x, _w=w):
        # deferred by one op: a BatchNorm that consumes y asks for `stats` and the conv epilogue gathers the
        # per-channel (count, mean, M2) partials on the way out -> no statistics pass over y (returns #partials);
        # a ReLU layer that consumes y (Conv2D(..., activation='relu')) asks for `act_out`: y and max(0, y) in one call
        split = rt.math_mode == x3.MATH_FP32X3 and x3.conv_ok(c, oc)
        if act_out is not None and not split:
            lib.xsb_conv2d_fwd_relu(_x.ptr, _w.ptr, barr.ptr if barr is not None else None, ybuf.data_ptr(), act_out, n, h, wd,
                                    c, oc, kh, kw, sh, sw, padding, x3.c_mode(False), rt.ws_big.data_ptr(), rt.WS_BIG_FLOATS,
                                    rt.stream())
            rt.launches += 1
            return True
        if split:
            # 3xTF32: three accumulating tcgen05 passes; the x split is kept for the weight gradient
            xparts = x3.split(_x.ptr, _x.size)
            if GLOBAL.COMPUTE_GRAD and weight.requires_grad:
                out.cache["x3_x"] = xparts
            x3.conv_fwd(xparts, x3.split(_w.ptr, _w.size), barr.ptr if barr is not None else None, ybuf.data_ptr(),
                        (n, h, wd, c, oc, kh, kw, sh, sw, padding))
            return 0                    # (statistics of a partial sum are useless: the BatchNorm runs its own pass)
        args = (_x.ptr, _w.ptr, barr.ptr if barr is not None else None, ybuf.data_ptr(), n, h, wd, c, oc, kh, kw, sh, sw,
                padding, x3.c_mode(False), rt.ws_big.data_ptr(), rt.WS_BIG_FLOATS)
        rt.launches += 1
        if not stats:
            lib.xsb_conv2d_fwd(*args, rt.stream())
            return 0
        nparts = ctypes.c_int(0)
        lib.xsb_conv2d_fwd_stats(*args, rt.ws_stats.data_ptr(), rt.WS_STATS_FLOATS, ctypes.byref(nparts), rt.stream())
        return nparts.value
    run.fusable_relu = False
    run.fusable_act = True
    run.conv_stats = True
    y.pending_op = run
    rg = inputs.requires_grad or weight.requires_grad or bool(bias is not None and bias.requires_grad)
    return _wire(out, (inputs, weight, bias), rg, Conv2DBackward, cache={"stride": (sh, sw), "padding": padding})


def max_pool2d(inputs: Tensor, kernel_size: int = 2, stride=None, padding: int = 0, out: Tensor = None) -> Tensor:
    """xs/nn/functional.py:445-491. Always the reference's im2col/argmax semantics (Q1): zero padding,
    first max wins, `cache['pool_argmax']` in (n,oh,ow,c) order."""
    _mark_inputs(inputs)
    if stride is None:
        stride = (kernel_size, kernel_size)
    sh, sw = pair(stride)
    k = int(kernel_size)
    x = inputs.arr.as_cl()
    n, c, h, w = x.shape
    oh = (h + 2 * padding - k) // sh + 1
    ow = (w + 2 * padding - k) // sw + 1
    out, y = _out_arr(out, (n, c, oh, ow))
    y.cl = True
    need_grad = GLOBAL.COMPUTE_GRAD and inputs.requires_grad
    amax = rt.empty(max(y.size, 1), "uint8") if need_grad else None
    aptr = amax.data_ptr() if amax is not None else None
    # the input is a not-yet-materialised BatchNorm -> in-place ReLU result: pool straight from the conv output
    # (BN apply + ReLU + mask + pool in one pass, the rectified tensor is never written unless someone else reads it)
    fused = False
    br = getattr(x.pending_op, "bn_relu", None) if x is inputs.arr else None
    if br is not None and br["x"].cl and (rt.fuse_bn_relu_pool == "all" or (k, sh, sw, padding) == (3, 2, 2, 1)):
        rc = lib.xsb_bn_relu_maxpool_fwd(br["x"].ptr, br["mean"].peek_ptr(), br["invstd"].peek_ptr(), br["gamma"].ptr,
                                         br["beta"].ptr, y.wptr, aptr, br["mask"].data_ptr(), n, h, w, c, k, sh, sw, padding,
                                         rt.stream())
        fused = rc == 0
    if not fused:
        lib.xsb_maxpool2d_fwd(x.ptr, y.wptr, aptr, n, h, w, c, k, sh, sw, padding, rt.stream())
    rt.launches += 1
    cache = {"mode": "im2col", "stride": (sh, sw), "padding": padding, "kernel_size": k}
    if need_grad:
        cache["pool_argmax"] = DeviceArgmax(amax, y.size)
    return _wire(out, (inputs,), inputs.requires_grad, Maxpool2DBackward, cache=cache)


def avg_pool2d(inputs: Tensor, kernel_size: int, stride=None, padding: int = 0, out: Tensor = None) -> Tensor:
    """xs/nn/functional.py:539-583 (padding must be 0: the reference's padded branch double-counts it)."""
    _mark_inputs(inputs)
    if padding != 0:
        raise NotImplementedError("avg_pool2d with padding != 0 is broken in the reference (functional.py:551-553)")
    if stride is None:
        stride = (kernel_size, kernel_size)
    sh, sw = pair(stride)
    k = int(kernel_size)
    x = inputs.arr.as_cl()
    n, c, h, w = x.shape
    oh = (h - k) // sh + 1
    ow = (w - k) // sw + 1
    if oh <= 0 or ow <= 0:
        raise ValueError(f"avg_pool2d: {h}x{w} input is smaller than the {k}x{k} window")
    out, y = _out_arr(out, (n, c, oh, ow))
    y.cl = True
    lib.xsb_avgpool2d_fwd(x.ptr, y.wptr, n, h, w, c, k, sh, sw, rt.stream())
    rt.launches += 1
    return _wire(out, (inputs,), inputs.requires_grad, Avgpool2DBackward,
                 cache={"mode": "im2col", "stride": (sh, sw), "padding": 0, "kernel_size": k})


# ---------------------------------------------------------------------------------------------- normalisation
def batch_norm(inputs: Tensor, gamma: Tensor, beta: Tensor, moving_mean: Tensor, moving_variance: Tensor,
               axis: int = 1, training: bool = True, epsilon: float = 1e-5, momentum: float = 0.9,
               out: Tensor = None) -> Tensor:
    """xs/nn/functional.py:676-740, training branch (the only one the reference ever reaches, Q5)."""
    _mark_inputs(inputs)
    x = inputs.arr
    nparts = 0
    producer = x.pending_op
    if training and producer is not None and getattr(producer, "conv_stats", False) and x.ndim == 4 and axis == 1:
        x.pending_op = None
        nparts = producer(None, stats=True)   # the producing conv runs now and leaves statistics partials
    if x.ndim == 4 and axis == 1:
        x = x.as_cl()
    elif x.ndim == 2 and axis in (1, -1):
        pass
    else:
        raise NotImplementedError(f"batch_norm over axis {axis} of a {x.ndim}-D tensor is outside the hot path")
    C = x.shape[1]
    R = x.size // C
    out, y = _out_arr(out, x.shape)
    y.cl = x.cl
    if not training:
        # inference branch (functional.py:708-725): normalise with the running statistics; no backward
        lib.xsb_bn_eval_fwd(x.ptr, gamma.arr.ptr, beta.arr.ptr, moving_mean.arr.ptr, moving_variance.arr.ptr, y.wptr, R,
                            C, float(epsilon), rt.ws_small.data_ptr(), rt.stream())
        rt.launches += 2
        out.requires_grad = False
        if out.is_dynamic:
            out.add_in_bounds(inputs, gamma, beta)
            inputs.add_out_bounds(out)
        return out
    save_mean, save_invstd = DevArray.empty((C,)), DevArray.empty((C,))
    if nparts > 0:
        lib.xsb_bn_finalize(rt.ws_stats.data_ptr(), nparts, R, C, moving_mean.arr.ptr, moving_variance.arr.ptr,
                            save_mean.wptr, save_invstd.wptr, float(epsilon), float(momentum), rt.stream())
    else:
        lib.xsb_bn_stats(x.ptr, moving_mean.arr.ptr, moving_variance.arr.ptr, save_mean.wptr, save_invstd.wptr, R, C,
                         float(epsilon), float(momentum), rt.ws_small.data_ptr(), rt.stream())
    rt.launches += 1
    ybuf = y.buf
    y.wptr  # noqa: B018 -- clears stale pending state of a reused (static) buffer

    def run(mask=None, _x=x, _g=gamma.arr, _b=beta.arr):
        # deferred apply pass: a following ReLU(inplace=True) fuses into it (y = max(0, bn(x)) + bit mask)
        lib.xsb_bn_apply(_x.ptr, ybuf.data_ptr(), save_mean.peek_ptr(), save_invstd.peek_ptr(), _g.ptr, _b.ptr, R, C,
                         0 if mask is None else 1, None if mask is None else mask.data_ptr(), rt.stream())
        rt.launches += 1
    run.fusable_relu = C % 4 == 0 and (R * (C // 4)) % 2 == 0 and x.cl == y.cl
    run.bn = dict(x=x, mean=save_mean, invstd=save_invstd, gamma=gamma.arr, beta=beta.arr, R=R, C=C)
    y.pending_op = run
    rg = inputs.requires_grad or gamma.requires_grad or beta.requires_grad
    axis_field = tuple(i for i in range(x.ndim) if i != (axis % x.ndim))
    return _wire(out, (inputs, gamma, beta), rg, BatchNormBackward,
                 cache={"axis_field": axis_field, "save_mean": save_mean, "save_invstd": save_invstd})


# ---------------------------------------------------------------------------------------------- SURVEY 8f-4 ops
def _unary(inputs, out, kind, grad_fn):
    _mark_inputs(inputs)
    x = inputs.arr
    out, y = _out_arr(out, x.shape)
    y.cl = x.cl
    lib.xsb_unary_fwd(kind, x.ptr, y.wptr, x.size, rt.stream())
    rt.launches += 1
    return _wire(out, (inputs,), inputs.requires_grad, grad_fn)


def sigmoid(inputs: Tensor, out: Tensor = None) -> Tensor:
    """xs/nn/functional.py:316-329: y = 1 / (1 + exp(-x))."""
    return _unary(inputs, out, 0, SigmoidBackward)


def tanh(inputs: Tensor, out: Tensor = None) -> Tensor:
    """xs/nn/functional.py:332-345."""
    return _unary(inputs, out, 1, TanhBackward)


class DeviceBitMask:
    """`cache['mask']` of dropout2d: one bit per element on the device (bit set = kept); np.asarray(...) yields the
    reference's 0/1 array in the tensor's PHYSICAL element order."""

    def __init__(self, dev, count):
        self.dev, self.count = dev, count

    def __array__(self, dtype=None, copy=None):
        bits = np.unpackbits(self.dev.cpu().numpy(), bitorder="little")[: self.count]
        return bits if dtype is None else bits.astype(dtype)


_DROPOUT_CALLS = [0]


def dropout2d(inputs: Tensor, keep_prob: float = 0.5, out: Tensor = None, seed: int = None) -> Tensor:
    """xs/nn/functional.py:650-674. Training: y = x * mask / keep_prob with mask ~ Bernoulli(keep_prob) per element, drawn
    by a stateless device hash of (seed, call counter, element index) -- not NumPy's binomial stream, so parity is defined
    GIVEN the mask (`out.cache['mask']`). Not training: the input passes through; without grad tracking it is returned."""
    _mark_inputs(inputs)
    if not GLOBAL.COMPUTE_GRAD:
        return inputs
    x = inputs.arr
    out, y = _out_arr(out, x.shape)
    y.cl = x.cl
    mask = None
    if GLOBAL.TRAINING:
        mask = rt.empty((x.size + 7) // 8, "uint8")
        _DROPOUT_CALLS[0] += 1
        base = int(np.random.get_state()[1][0]) if seed is None else int(seed)     # follows np.random.seed(...)
        lib.xsb_dropout_fwd(x.ptr, y.wptr, mask.data_ptr(), x.size, float(keep_prob),
                            (base * 1000003 + _DROPOUT_CALLS[0]) & 0x7FFFFFFFFFFFFFFF, None, rt.stream())
        rt.launches += 1
    else:
        x.flush()
        y.wptr  # noqa: B018
        y.buf[: max(x.size, 1)].copy_(x.buf[: max(x.size, 1)])
    cache = {"mask": DeviceBitMask(mask, x.size) if mask is not None else None, "keep_prob": keep_prob}
    return _wire(out, (inputs,), inputs.requires_grad, Dropout2DBackward, cache=cache)


def pad2d(inputs: Tensor, padding: tuple, out: Tensor = None) -> Tensor:
    """xs/nn/functional.py:629-647: zero padding of the two spatial axes by (padding[0], padding[1]) on both sides."""
    _mark_inputs(inputs)
    x = inputs.arr.as_cl()
    n, c, h, w = x.shape
    ph, pw = int(padding[0]), int(padding[1])
    out, y = _out_arr(out, (n, c, h + 2 * ph, w + 2 * pw))
    y.cl = True
    lib.xsb_pad2d_fwd(x.ptr, y.wptr, n, h, w, c, ph, pw, rt.stream())
    rt.launches += 1
    return _wire(out, (inputs,), inputs.requires_grad, Pad2DBackward, cache={"padding": (ph, pw)})


def layer_norm(inputs: Tensor, gamma: Tensor, beta: Tensor, epsilon: float = 1e-10, out: Tensor = None) -> Tensor:
    """xs/nn/functional.py:743-775: statistics over every axis but 0 (biased variance), per-element gamma / beta of the
    sample's shape. Computed on the logical (NCHW row-major) order, where a sample is one contiguous row."""
    _mark_inputs(inputs)
    x = inputs.arr.logical_rowmajor()
    n = x.shape[0]
    f = x.size // n
    if gamma.arr.size != f or beta.arr.size != f:
        raise ValueError(f"layer_norm: gamma / beta must have the sample's {f} elements")
    out, y = _out_arr(out, x.shape)
    want_cl = y.cl and len(x.shape) == 4 and out is not None and out.is_static
    y.cl = False
    mean, rstd = DevArray.empty((n,)), DevArray.empty((n,))
    lib.xsb_layernorm_fwd(x.ptr, gamma.arr.logical_rowmajor().ptr, beta.arr.logical_rowmajor().ptr, y.wptr, mean.wptr, rstd.wptr,
                          n, f, float(epsilon), rt.stream())
    rt.launches += 1
    if want_cl:
        conv = y.as_cl()
        if conv is not y:
            y.buf[: y.size].copy_(conv.buf[: y.size])
        y.cl = True
    rg = inputs.requires_grad or gamma.requires_grad or beta.requires_grad
    return _wire(out, (inputs, gamma, beta), rg, LayerNormBackward, cache={"mean": mean, "rstd": rstd})


def group_norm(inputs: Tensor, gamma: Tensor, beta: Tensor, epsilon: float = 1e-5, groups: int = 16,
               out: Tensor = None) -> Tensor:
    """xs/nn/functional.py:778-812: statistics per (sample, group of C/groups channels) over (c, h, w), per-channel affine."""
    _mark_inputs(inputs)
    x = inputs.arr.as_cl()
    n, c, h, w = x.shape
    if c % groups != 0:
        raise ValueError(f"group_norm: {c} channels are not divisible into {groups} groups")
    out, y = _out_arr(out, x.shape)
    y.cl = True
    mean, rstd = DevArray.empty((n * groups,)), DevArray.empty((n * groups,))
    lib.xsb_groupnorm_fwd(x.ptr, gamma.arr.ptr, beta.arr.ptr, y.wptr, mean.wptr, rstd.wptr, n, h * w, c, int(groups),
                          float(epsilon), rt.stream())
    rt.launches += 1
    rg = inputs.requires_grad or gamma.requires_grad or beta.requires_grad
    return _wire(out, (inputs, gamma, beta), rg, GroupNormBackward, cache={"mean": mean, "rstd": rstd, "groups": int(groups)})


# ---------------------------------------------------------------------------------------------- losses
def _loss_out(out, reduction):
    if reduction not in ("mean", "sum"):
        raise TypeError("unknown type for reduction" if reduction != "none" else "reduction='none' is outside the hot path")
    return _out_arr(out, (1,))


def mse_loss(x1: Tensor, x2: Tensor, reduction: str = "mean", out: Tensor = None) -> Tensor:
    """xs/nn/functional.py:843-873: 'mean' = sum((a-b)^2) / 2 / numel (Q9)."""
    assert x1.shape == x2.shape
    _mark_inputs(x1)
    a = x1.arr
    b = x2.arr
    if a.cl != b.cl:
        b = b.as_cl() if a.cl else b.logical_rowmajor()
    out, y = _loss_out(out, reduction)
    lib.xsb_mse_fwd(a.ptr, b.ptr, y.wptr, a.size, 0 if reduction == "mean" else 1, rt.stream())
    rt.launches += 1
    return _wire(out, (x1, x2), x1.requires_grad or x2.requires_grad, MSELossBackward,
                 cache={"reduction": reduction})


def cross_entropy(x1: Tensor, x2: Tensor, reduction: str = "mean", out: Tensor = None) -> Tensor:
    """xs/nn/functional.py:984-1010: softmax -> log -> NLL, fused in one kernel; the softmax is kept
    in `x1.cache['softmax']` for the backward; the loss is marked leaf so `.item()` survives backward (Q14)."""
    _mark_inputs(x1)
    out, y = _loss_out(out, reduction)
    x2.long_()
    z = x1.arr.logical_rowmajor()
    c = z.shape[-1]
    rows = z.size // c
    if x2.arr.size != rows:
        raise ValueError(f"cross_entropy: {rows} rows of logits but {x2.arr.size} labels")
    sm = x1.cache.get("softmax")
    if sm is None or sm._arr is None or tuple(sm.arr.shape) != tuple(z.shape):
        sm = Tensor(DevArray.empty(z.shape))
        sm._arr.cl = False
    x1.cache["softmax"] = sm
    lib.xsb_softmax_xent_fwd(z.ptr, x2.arr.ptr, sm.arr.wptr, y.wptr, rows, c, 0 if reduction == "mean" else 1,
                             rt.label_flag.data_ptr(), rt.stream())
    rt.launches += 1
    out.cache["label_check"] = True     # a NaN loss is looked up against the kernel's bad-label flag on its first read
    if rows != z.shape[0] and reduction == "mean":
        raise NotImplementedError("cross_entropy 'mean' over >2-D logits")
    out.requires_grad = x1.requires_grad or x2.requires_grad
    if out.is_dynamic:
        out.add_in_bounds(x1, x2)
        x1.add_out_bounds(out)
        x2.add_out_bounds(out)
        out.is_leaf = True
    else:
        out._in_bounds = [x1, x2]   # static loss buffer: this step's logits and labels
    if GLOBAL.COMPUTE_GRAD and out.requires_grad:
        out.cache["reduction"] = reduction
        out.grad_fn = CrossEntropyLossBackward
        initialize_ops_grad(x1, x2)
    return out
