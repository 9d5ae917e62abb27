"""Backward closures `XxxBackward(outputs) -> None` (mirror of xs/nn/grad_fn.py).

Contract kept from the reference: a closure reads `outputs.grad`, `outputs.in_bounds` and
`outputs.cache`, and ACCUMULATES into `in_bound.grad` for every operand that requires grad.
Each one is a thin launcher of libxsb200 kernels; "accumulate" is passed to the kernel as a flag
(0 the first time a freshly zeroed gradient is written, 1 afterwards), so no zero-fill and no
separate `+=` pass is ever run.
"""
import ctypes

from ..core.device import DevArray, fill_floats, rt
from ..core.tensor import Parameter, Tensor
from .._lib import lib


def _done(t):
    """Bookkeeping for the data-parallel engine: one pending use of a Parameter has produced its gradient."""
    if isinstance(t, Parameter) and t.pending_uses > 0:
        t.pending_uses -= 1


def _grad_arr(t):
    if t.grad is None:
        t.zero_grad()
    return t.grad.arr


def _same_layout(src, like):
    """src DevArray re-laid to match `like`'s physical order (both have equal logical shapes)."""
    if src.cl == like.cl:
        return src
    return src.as_cl() if like.cl else src.logical_rowmajor()


def accumulate_into(t, src):
    """t.grad += src (src: DevArray of t's logical shape)."""
    g = _grad_arr(t)
    src.flush()                          # (a postponed add would otherwise hide behind pending_zero)
    src = _same_layout(src, g)
    if g.take_acc():
        lib.xsb_axpy(g.peek_ptr(), src.ptr, 1.0, g.size, rt.stream())
        rt.launches += 1
    elif src.pending_zero:
        fill_floats(g.buf, 0.0, g.size)
    else:
        src.flush()
        g.buf[: g.size].copy_(src.buf[: g.size])   # first write: a plain device-to-device copy
    _done(t)


def accumulate_rowmajor_into(t, src_rm):
    """t.grad += src where src is in LOGICAL row-major order and t may be channels-last."""
    g = _grad_arr(t)
    if g.cl:
        n, c, h, w = g.shape
        if c > 1 and h * w > 1:
            lib.xsb_nchw_to_nhwc(src_rm.ptr, g.peek_ptr(), n, c, h, w, g.take_acc(), rt.stream())
            rt.launches += 1
            return
    accumulate_into(t, DevArray(src_rm.buf, g.shape, g.dtype, cl=g.cl, _share=src_rm))


# ------------------------------------------------------------------ elementwise
def AddBackward(outputs: Tensor, relu_mask=None):
    """xs/nn/grad_fn.py:6-16 (same-shape operands). With `relu_mask` (handed over by the in-place ReLUBackward
    that precedes it) the mask is applied while the gradient is written into each operand: one pass per
    operand instead of mask-in-place + copies."""
    dy = outputs.grad.arr
    targets = [t for t in outputs.in_bounds if t.requires_grad]
    bcast = outputs.cache.get("broadcast")
    if bcast is not None:
        # a row-vector operand ([M, N] + [1, N] or [N]): its gradient is dY summed over the first mismatching axis
        # (xs/nn/grad_fn.py:9-12); the full-shape operand takes dY itself
        if relu_mask is not None:
            lib.xsb_relu_bwd_mask(dy.ptr, relu_mask.data_ptr(), dy.peek_ptr(), dy.size, 0, rt.stream())
            rt.launches += 1
        rows, cols = bcast
        for t in targets:
            if t.arr.size == dy.size:
                accumulate_into(t, dy)
            else:
                g = _grad_arr(t)
                src = dy.logical_rowmajor()
                tmp = DevArray.empty((cols,))
                lib.xsb_reduce_axis(src.ptr, tmp.wptr, 1, rows, cols, 1.0, rt.stream())
                rt.launches += 1
                accumulate_into(t, DevArray(tmp.buf, g.shape, g.dtype, cl=g.cl))
        return
    if relu_mask is not None:
        if all(_grad_arr(t).cl == dy.cl for t in targets):
            aliases, plain = [], []
            for t in targets:
                g = _grad_arr(t)
                if (t.grad_fn is BatchNormBackward and t.is_dynamic and len(t.out_bounds) == 1 and g.pending_zero
                        and not t.cache.get("inplace") and dy.lazy_mask is None and tuple(g.shape) == tuple(dy.shape)):
                    # the branch ends in a BatchNorm whose only consumer is this add: hand it dZ itself plus the
                    # mask -- its backward kernels mask on the fly, no masked copy is written and read back
                    alias = DevArray(dy.buf, g.shape, g.dtype, cl=g.cl)
                    dy.ptr  # noqa: B018 -- dZ itself must be materialised before it is aliased
                    alias.lazy_mask = relu_mask
                    t.grad._arr = alias
                    aliases.append(alias)
                    continue
                plain.append((t, g))
            for t, g in plain:
                dy_ptr = dy.ptr

                def masked_copy(acc, _g=g, _ptr=dy_ptr, _keep=(dy.buf, relu_mask)):
                    lib.xsb_relu_bwd_mask(_ptr, relu_mask.data_ptr(), _g.peek_ptr(), _g.size, acc, rt.stream())
                    rt.launches += 1
                # identity shortcut next to a BatchNorm branch: that BatchNorm's backward reads the same masked dZ -- it
                # writes this copy on the way (xsb_bn_bwd_dxsum_side). Parked like a postponed add: a reader that comes
                # first, or a BatchNorm kernel without the side output, still gets it from relu_bwd_mask
                if (rt.fuse_add_bwd_side and len(aliases) == 1 and aliases[0].side_out is None and t.is_dynamic
                        and not isinstance(t, Parameter)      # (a parameter's gradient may be all-reduced as soon as _done)
                        and tuple(g.shape) == tuple(dy.shape) and g.defer_add(masked_copy)):
                    aliases[0].side_out = (g, masked_copy)
                else:
                    masked_copy(g.take_acc())
                _done(t)
            return
        lib.xsb_relu_bwd_mask(dy.ptr, relu_mask.data_ptr(), dy.peek_ptr(), dy.size, 0, rt.stream())
        rt.launches += 1
    for t in targets:
        accumulate_into(t, dy)


def ReLUBackward(outputs: Tensor):
    """xs/nn/grad_fn.py:146-157. In-place form: mask outputs.grad in place with the bit mask
    recorded by the forward, pop the producer's closure and run it; otherwise dX += dY*[x>=0]."""
    if outputs.cache.get("inplace"):
        mask = outputs.cache["mask"]
        fn = outputs.cache["grad_fn"].pop()
        outputs.grad_fn = fn
        if (fn is BatchNormBackward or fn is AddBackward) and outputs.is_dynamic:
            fn(outputs, relu_mask=mask)      # the producer's backward kernels apply the mask on the fly
            return
        # (static graph: the producer LAYER runs the restored closure once more after this one -- reference behaviour,
        # xs/layers/base.py:126-134 -- so the gradient itself must carry the mask, as the reference's in-place masking does)
        g = outputs.grad.arr
        lib.xsb_relu_bwd_mask(g.ptr, mask.data_ptr(), g.peek_ptr(), g.size, 0, rt.stream())
        rt.launches += 1
        if fn is not None:
            fn(outputs)
    else:
        inputs, = outputs.in_bounds
        if inputs.requires_grad:
            dy = outputs.grad.arr
            g = _grad_arr(inputs)
            x = _same_layout(inputs.arr, g)
            dy = _same_layout(dy, g)
            lib.xsb_relu_bwd(dy.ptr, x.ptr, g.peek_ptr(), g.size, g.take_acc(), rt.stream())
            rt.launches += 1


# ------------------------------------------------------------------ shape ops
def FlattenBackward(outputs: Tensor):
    """xs/nn/grad_fn.py:173-176: dX += reshape(dY) (dY is in NCHW feature order)."""
    inputs, = outputs.in_bounds
    if inputs.requires_grad:
        accumulate_rowmajor_into(inputs, outputs.grad.arr.logical_rowmajor())


ViewBackward = FlattenBackward  # xs/nn/grad_fn.py:497-500


def MeanBackward(outputs: Tensor):
    """xs/nn/grad_fn.py:128-137: dX += broadcast(dY) / (#reduced)."""
    inputs, = outputs.in_bounds
    if inputs.requires_grad:
        outer, red, inner = outputs.cache["ori"]
        scale = 1.0 / red if outputs.cache["op"] == "mean" else 1.0
        tmp = DevArray.empty(inputs.shape)
        tmp.cl = False
        lib.xsb_broadcast_axis(outputs.grad.arr.logical_rowmajor().ptr, tmp.wptr, outer, red, inner, scale, 0,
                               rt.stream())
        rt.launches += 1
        accumulate_rowmajor_into(inputs, tmp)


SumBackward = MeanBackward


# ------------------------------------------------------------------ Dense
def _gemm(a, b, c_arr, bias, M, N, K, ta, tb, acc):
    from . import x3
    if rt.math_mode == x3.MATH_FP32X3 and x3.gemm_ok(M if ta else K, K if tb else N, N):
        x3.gemm(a, M * K, b, K * N, c_arr, bias, M, N, K, ta, tb, acc)
        return
    lib.xsb_gemm(a, b, c_arr, bias, M, N, K, ta, tb, acc, x3.c_mode(False), rt.ws_big.data_ptr(), rt.WS_BIG_FLOATS,
                 rt.stream())
    rt.launches += 1


def MMBackward(outputs: Tensor):
    """xs/nn/grad_fn.py:73-83: dX += dY.W^T ; dW += X^T.dY."""
    x, w = outputs.in_bounds
    dy = outputs.grad.arr
    m, n = dy.shape
    k = x.shape[-1]
    if x.requires_grad:
        g = _grad_arr(x)
        _gemm(dy.ptr, w.arr.ptr, g.peek_ptr(), None, m, k, n, 0, 1, g.take_acc())
    if w.requires_grad:
        g = _grad_arr(w)
        _gemm(x.arr.ptr, dy.ptr, g.peek_ptr(), None, k, n, m, 1, 0, g.take_acc())
        _done(w)


def AddmmBackward(outputs: Tensor):
    """xs/nn/grad_fn.py:86-99: db += sum_rows(dY) (dY itself when N == 1, Q12), then MMBackward."""
    x1, x2, b = outputs.in_bounds
    dy = outputs.grad.arr
    m, n = dy.shape
    if b.requires_grad:
        g = _grad_arr(b)
        lib.xsb_colsum(dy.ptr, g.peek_ptr(), m, n, g.take_acc(), rt.ws_small.data_ptr(), rt.stream())
        rt.launches += 1
        _done(b)
    k = x1.shape[-1]
    if x1.requires_grad:
        g = _grad_arr(x1)
        _gemm(dy.ptr, x2.arr.ptr, g.peek_ptr(), None, m, k, n, 0, 1, g.take_acc())
    if x2.requires_grad:
        g = _grad_arr(x2)
        _gemm(x1.arr.ptr, dy.ptr, g.peek_ptr(), None, k, n, m, 1, 0, g.take_acc())
        _done(x2)


# ------------------------------------------------------------------ conv / pool / bn
def Conv2DBackward(outputs: Tensor):
    """xs/nn/grad_fn.py:186-203: dW += dY^T.col ; db += sum dY ; dX += col2im(dY.W) -- as three implicit
    GEMM / reduction kernels over the channels-last tensors (no col buffer, no col2im scatter). The
    reference's unconditional dcol GEMM (line 199) is skipped when the input needs no gradient."""
    inputs, weight, bias = outputs.in_bounds
    x = inputs.arr.as_cl()
    dy = outputs.grad.arr.as_cl()
    n, c, h, w = x.shape
    oc, _, kh, kw = weight.shape
    (sh, sw), pad = outputs.cache["stride"], outputs.cache["padding"]
    ws, wsn, st = rt.ws_big.data_ptr(), rt.WS_BIG_FLOATS, rt.stream()
    from . import x3
    geom = (n, h, w, c, oc, kh, kw, sh, sw, pad)
    split_mode = rt.math_mode == x3.MATH_FP32X3
    mode = x3.c_mode(False)
    xparts = outputs.cache.pop("x3_x", None)
    dyparts = None
    if weight.requires_grad:
        g = _grad_arr(weight)
        if split_mode and x3.conv_ok(c, oc):
            dyparts = x3.split(dy.ptr, dy.size)
            x3.conv_wgrad(dyparts, xparts if xparts is not None else x3.split(x.ptr, x.size), g.peek_ptr(), geom, g.take_acc())
        else:
            lib.xsb_conv2d_wgrad(dy.ptr, x.ptr, g.peek_ptr(), *geom, g.take_acc(), mode, ws, wsn, st)
            rt.launches += 1
        del xparts
        _done(weight)
    if bias is not None and bias.requires_grad:
        if outputs.cache.pop("bias_grad_done", False):
            pass   # the BatchNorm backward that produced dY summed its columns into bias.grad on the way
        else:
            g = _grad_arr(bias)
            lib.xsb_colsum(dy.ptr, g.peek_ptr(), dy.size // oc, oc, g.take_acc(), rt.ws_small.data_ptr(), st)
            rt.launches += 1
        _done(bias)
    if inputs.requires_grad:
        g = _grad_arr(inputs)
        assert g.cl, "conv input gradient must be channels-last"
        dy_ptr, w_ptr = dy.ptr, weight.arr.ptr
        if split_mode and x3.dgrad_ok(c, oc, kh, kw, sh, sw):
            dyparts = dyparts if dyparts is not None else x3.split(dy_ptr, dy.size)
            wparts = x3.split(w_ptr, weight.arr.size)

            def dgrad(acc, _keep=(dy.buf, weight.arr.buf)):
                x3.conv_dgrad(dyparts, wparts, g.peek_ptr(), geom, acc)
        else:
            def dgrad(acc, _keep=(dy.buf, weight.arr.buf)):      # the buffers stay alive while the kernel is postponed
                lib.xsb_conv2d_dgrad(dy_ptr, w_ptr, g.peek_ptr(), *geom, acc, mode, ws, wsn, st)
                rt.launches += 1

        # a filter smaller than its stride (the 1x1/2 shortcut) reaches only some parity classes of dX: postponed until
        # dX is read, so that the other branch of the block writes "=" and this one only adds (see DevArray.defer_add)
        if (kh < sh or kw < sw) and g.defer_add(dgrad):
            return
        dgrad(g.take_acc())


def Maxpool2DBackward(outputs: Tensor):
    """xs/nn/grad_fn.py:206-235 (im2col branch; the 'reshape' branch is dead code, Q1)."""
    inputs, = outputs.in_bounds
    if inputs.requires_grad:
        k = outputs.cache["kernel_size"]
        sh, sw = outputs.cache["stride"]
        pad = outputs.cache["padding"]
        amax = outputs.cache["pool_argmax"].dev
        n, c, h, w = inputs.shape
        g = _grad_arr(inputs)
        dy = outputs.grad.arr.as_cl()
        dy_ptr = dy.ptr

        def scatter(acc, _keep=(dy.buf, amax)):
            lib.xsb_maxpool2d_bwd(dy_ptr, amax.data_ptr(), g.peek_ptr(), n, h, w, c, k, sh, sw, pad, acc, rt.stream())
            rt.launches += 1

        # BatchNorm -> in-place ReLU -> this 3x3/2 pool (the ResNet stem): the BatchNorm backward can gather the pool's
        # gradient on the fly (xsb_bn_bwd_pool_dxsum) -- 4x less to read than the scattered copy, which is then never
        # written. Parked like a postponed add: anyone else who reads the gradient gets the scattered copy first.
        stack = inputs.cache.get("grad_fn")
        if (rt.fuse_pool_bn_bwd and k == 3 and (sh, sw) == (2, 2) and pad == 1 and inputs.cache.get("inplace") and stack
                and stack[-1] is BatchNormBackward and inputs.is_dynamic and len(inputs.out_bounds) == 1 and g.cl):
            scatter.pool = (dy_ptr, amax.data_ptr(), (n, h, w, c))
            if g.defer_add(scatter):
                return
        scatter(g.take_acc())


def Avgpool2DBackward(outputs: Tensor):
    """xs/nn/grad_fn.py:284-309: dX += col2im(repeat(dY)/k^2)."""
    inputs, = outputs.in_bounds
    if inputs.requires_grad:
        k = outputs.cache["kernel_size"]
        sh, sw = outputs.cache["stride"]
        n, c, h, w = inputs.shape
        g = _grad_arr(inputs)
        dy = outputs.grad.arr.as_cl()
        lib.xsb_avgpool2d_bwd(dy.ptr, g.peek_ptr(), n, h, w, c, k, sh, sw, g.take_acc(), rt.stream())
        rt.launches += 1


def BatchNormBackward(outputs: Tensor, relu_mask=None):
    """xs/nn/grad_fn.py:379-411. xhat is recomputed from the retained input and the saved
    mean / invstd (the reference caches xhat and sqrtvar and destroys them here). `relu_mask`: the bit mask of
    an in-place ReLU applied to this BN's output; dY is masked inside the kernels (fused ReLUBackward)."""
    inputs, gamma, beta = outputs.in_bounds
    dy = outputs.grad.arr
    x = inputs.arr
    # dy is the still-parked gradient of a 3x3/2 max-pool (Maxpool2DBackward): gather it inside the BatchNorm kernels
    parked = dy._pz[2] if dy.pending_zero else None
    pool = getattr(parked, "pool", None)
    if pool is not None and inputs.requires_grad and x.ndim == 4 and x.cl and dy.cl and dy.lazy_mask is None:
        dpool_ptr, amax_ptr, (pn, ph, pw, pc) = pool
        gx = _grad_arr(inputs)
        gg = _grad_arr(gamma) if gamma.requires_grad else None
        gb = _grad_arr(beta) if beta.requires_grad else None
        cb = None
        if (inputs.grad_fn is Conv2DBackward and inputs.is_dynamic and len(inputs.out_bounds) == 1 and gx.pending_zero):
            conv_bias = inputs.in_bounds[2]
            if conv_bias is not None and conv_bias.requires_grad:
                cb = _grad_arr(conv_bias)
        rc = lib.xsb_bn_bwd_pool_dxsum(
            dpool_ptr, amax_ptr, x.ptr, gamma.arr.ptr, outputs.cache["save_mean"].ptr, outputs.cache["save_invstd"].ptr,
            gx.peek_ptr(), gg.peek_ptr() if gg is not None else None, gb.peek_ptr() if gb is not None else None,
            pn, ph, pw, pc, 0 if gx.pending_zero else 1, 0 if (gg is None or gg.pending_zero) else 1,
            0 if (gb is None or gb.pending_zero) else 1, relu_mask.data_ptr() if relu_mask is not None else None,
            rt.ws_small.data_ptr(), cb.peek_ptr() if cb is not None else None,
            0 if (cb is None or cb.pending_zero) else 1, rt.stream())
        if rc == 0:
            dy._pz[2] = None           # consumed: the scattered copy is never made (dy stays logically zero)
            for a in (gx, gg, gb, cb):
                if a is not None:
                    a.pending_zero = False
            if cb is not None:
                inputs.cache["bias_grad_done"] = True
            rt.launches += 2
            _done(gamma)
            _done(beta)
            return
    if dy.lazy_mask is not None and relu_mask is None and (x.ndim != 4 or (x.cl and dy.cl)):
        relu_mask, dy_ptr = dy.lazy_mask, dy.buf.data_ptr()     # masked alias of the add's dZ (see AddBackward)
    else:
        dy_ptr = None
    if x.ndim == 4:
        x, dy = x.as_cl(), dy.as_cl()
    if dy_ptr is None:
        dy_ptr = dy.ptr
    C = gamma.arr.size
    R = x.size // C
    gx = _grad_arr(inputs) if inputs.requires_grad else None
    gg = _grad_arr(gamma) if gamma.requires_grad else None
    gb = _grad_arr(beta) if beta.requires_grad else None
    # x is the output of a convolution with a trainable bias and this BN is its only consumer: the bias gradient
    # (column sums of dX, xs/nn/grad_fn.py:196-197) is gathered while dX is written instead of by a second pass
    cb = None
    if (gx is not None and x.ndim == 4 and inputs.grad_fn is Conv2DBackward and inputs.is_dynamic
            and len(inputs.out_bounds) == 1 and gx.pending_zero):
        conv_bias = inputs.in_bounds[2]
        if conv_bias is not None and conv_bias.requires_grad:
            cb = _grad_arr(conv_bias)
    # the identity shortcut of the residual block wants the same masked dZ written out (AddBackward parked that copy): the
    # apply pass writes it on the way, if it is still parked and the gradient it belongs to is still untouched
    side = None
    want = getattr(outputs.grad.arr, "side_out", None)
    if want is not None and relu_mask is not None and gx is not None and x.ndim == 4:
        sg, parked = want
        if sg._pz[2] is parked and sg.pending_zero and sg.lazy_mask is None and sg.size == R * C:
            side = sg
    if cb is not None or side is not None:
        was_zero = cb.pending_zero if cb is not None else False
        done, sdone = ctypes.c_int(0), ctypes.c_int(0)
        args = (dy_ptr, x.ptr, gamma.arr.ptr, outputs.cache["save_mean"].ptr, outputs.cache["save_invstd"].ptr,
                gx.peek_ptr(), gg.peek_ptr() if gg is not None else None,
                gb.peek_ptr() if gb is not None else None, R, C, gx.take_acc(),
                gg.take_acc() if gg is not None else 0, gb.take_acc() if gb is not None else 0,
                relu_mask.data_ptr() if relu_mask is not None else None, rt.ws_small.data_ptr(),
                cb.peek_ptr() if cb is not None else None, cb.take_acc() if cb is not None else 0, ctypes.byref(done))
        if side is not None:
            lib.xsb_bn_bwd_dxsum_side(*args, side.buf.data_ptr(), ctypes.byref(sdone), rt.stream())
            if sdone.value:                 # written as the first "=" writer: the parked relu_bwd_mask is void
                side._pz[2] = None
                side.pending_zero = False
        else:
            lib.xsb_bn_bwd_dxsum(*args, rt.stream())
        if cb is not None:
            if done.value:
                inputs.cache["bias_grad_done"] = True
            else:
                cb.pending_zero = was_zero
        rt.launches += 2
        _done(gamma)
        _done(beta)
        return
    lib.xsb_bn_bwd(dy_ptr, x.ptr, gamma.arr.ptr, outputs.cache["save_mean"].ptr, outputs.cache["save_invstd"].ptr,
                   gx.peek_ptr() if gx is not None else None, gg.peek_ptr() if gg is not None else None,
                   gb.peek_ptr() if gb is not None else None, R, C,
                   gx.take_acc() if gx is not None else 0, gg.take_acc() if gg is not None else 0,
                   gb.take_acc() if gb is not None else 0, relu_mask.data_ptr() if relu_mask is not None else None,
                   rt.ws_small.data_ptr(), rt.stream())
    rt.launches += 2 if gx is not None else 1
    _done(gamma)
    _done(beta)


# ------------------------------------------------------------------ losses
def CrossEntropyLossBackward(outputs: Tensor):
    """xs/nn/grad_fn.py:604-620: d logits += (softmax - onehot)/N * dLoss."""
    logits, labels = outputs.in_bounds
    if logits.requires_grad:
        probs = logits.cache["softmax"].arr
        n = logits.shape[0]
        c = logits.shape[-1]
        rows = probs.size // c
        g = _grad_arr(logits)
        red = 0 if outputs.cache["reduction"] == "mean" else 1
        # 'mean' divides by shape[0] (grad_fn.py:614-616) even when logits have more leading axes
        lib.xsb_softmax_xent_bwd(probs.ptr, labels.arr.ptr, outputs.grad.arr.ptr, g.peek_ptr(), rows, c, red,
                                 g.take_acc(), rt.stream())
        rt.launches += 1
        if rows != n and red == 0:
            raise NotImplementedError("cross_entropy 'mean' over >2-D logits")


def MSELossBackward(outputs: Tensor):
    """xs/nn/grad_fn.py:503-520: d pred += (pred-true)*dLoss (/numel for 'mean'); d true -= the same."""
    y_pred, y_true = outputs.in_bounds
    a, b = y_pred.arr, _same_layout(y_true.arr, y_pred.arr)
    ga = _grad_arr(y_pred) if y_pred.requires_grad else None
    gb = _grad_arr(y_true) if y_true.requires_grad else None
    if ga is None and gb is None:
        return
    red = 0 if outputs.cache["reduction"] == "mean" else 1
    lib.xsb_mse_bwd(a.ptr, b.ptr, outputs.grad.arr.ptr, ga.peek_ptr() if ga is not None else None,
                    gb.peek_ptr() if gb is not None else None, a.size, red,
                    ga.take_acc() if ga is not None else 0, gb.take_acc() if gb is not None else 0, rt.stream())
    rt.launches += 1


# ------------------------------------------------------------------ SURVEY 8f-4 ops around the hot path
def _unary_backward(kind):
    def closure(outputs: Tensor):
        inputs, = outputs.in_bounds
        if inputs.requires_grad:
            g = _grad_arr(inputs)
            y, dy = _same_layout(outputs.arr, g), _same_layout(outputs.grad.arr, g)
            lib.xsb_unary_bwd(kind, dy.ptr, y.ptr, g.peek_ptr(), g.size, g.take_acc(), rt.stream())
            rt.launches += 1
    return closure


SigmoidBackward = _unary_backward(0)      # xs/nn/grad_fn.py:161-164: dX += dY * y * (1 - y)
SigmoidBackward.__name__ = "SigmoidBackward"
TanhBackward = _unary_backward(1)         # xs/nn/grad_fn.py:167-170: dX += dY * (1 - y^2)
TanhBackward.__name__ = "TanhBackward"


def Dropout2DBackward(outputs: Tensor):
    """xs/nn/grad_fn.py:319-325: dX += mask * keep_prob * dY (the reference multiplies by keep_prob here; kept)."""
    inputs, = outputs.in_bounds
    if inputs.requires_grad:
        g = _grad_arr(inputs)
        dy = _same_layout(outputs.grad.arr, outputs.arr)
        if g.cl != outputs.arr.cl:
            raise NotImplementedError("dropout backward across layouts")
        lib.xsb_dropout_bwd(dy.ptr, outputs.cache["mask"].dev.data_ptr(), g.peek_ptr(), g.size, float(outputs.cache["keep_prob"]),
                            g.take_acc(), rt.stream())
        rt.launches += 1


def Pad2DBackward(outputs: Tensor):
    """xs/nn/grad_fn.py:312-316: dX += the un-padded interior of dY."""
    inputs, = outputs.in_bounds
    if inputs.requires_grad:
        ph, pw = outputs.cache["padding"]
        n, c, h, w = inputs.shape
        g = _grad_arr(inputs)
        assert g.cl
        lib.xsb_pad2d_bwd(outputs.grad.arr.as_cl().ptr, g.peek_ptr(), n, h, w, c, ph, pw, g.take_acc(), rt.stream())
        rt.launches += 1


def LayerNormBackward(outputs: Tensor):
    """xs/nn/grad_fn.py:414-447: dgamma += sum_n dY*xhat, dbeta += sum_n dY, dX = closed form of the dvar / dmean chain.
    (As shipped the reference reads `inputs.cache['normalized_x']` at line 418 and raises KeyError whenever gamma is
    trainable; the formulas are its own.)"""
    inputs, gamma, beta = outputs.in_bounds
    x = inputs.arr.logical_rowmajor()
    dy = outputs.grad.arr.logical_rowmajor()
    n = x.shape[0]
    f = x.size // n
    gx = DevArray.empty(x.shape) if inputs.requires_grad else None
    if gx is not None:
        gx.cl = False
    gg = _grad_arr(gamma) if gamma.requires_grad else None
    gb = _grad_arr(beta) if beta.requires_grad else None
    lib.xsb_layernorm_bwd(dy.ptr, x.ptr, gamma.arr.logical_rowmajor().ptr, outputs.cache["mean"].ptr, outputs.cache["rstd"].ptr,
                          gx.wptr if gx is not None else None, gg.peek_ptr() if gg is not None else None,
                          gb.peek_ptr() if gb is not None else None, n, f, 0, gg.take_acc() if gg is not None else 0,
                          gb.take_acc() if gb is not None else 0, rt.stream())
    rt.launches += 2
    if gx is not None:
        accumulate_rowmajor_into(inputs, gx)
    _done(gamma)
    _done(beta)


def GroupNormBackward(outputs: Tensor):
    """xs/nn/grad_fn.py:450-490 (per-channel dgamma / dbeta over (n, h, w); the shipped code adds a keepdims [1,C,1,1] sum
    into a (C,) gradient and fails to broadcast -- the sums are its own)."""
    inputs, gamma, beta = outputs.in_bounds
    x = inputs.arr.as_cl()
    dy = outputs.grad.arr.as_cl()
    n, c, h, w = x.shape
    gx = _grad_arr(inputs) if inputs.requires_grad else None
    gg = _grad_arr(gamma) if gamma.requires_grad else None
    gb = _grad_arr(beta) if beta.requires_grad else None
    lib.xsb_groupnorm_bwd(dy.ptr, x.ptr, gamma.arr.ptr, outputs.cache["mean"].ptr, outputs.cache["rstd"].ptr,
                          gx.peek_ptr() if gx is not None else None, gg.peek_ptr() if gg is not None else None,
                          gb.peek_ptr() if gb is not None else None, n, h * w, c, outputs.cache["groups"],
                          gx.take_acc() if gx is not None else 0, gg.take_acc() if gg is not None else 0,
                          gb.take_acc() if gb is not None else 0, rt.stream())
    rt.launches += 2
    _done(gamma)
    _done(beta)
