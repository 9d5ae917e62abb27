"""NumPy restatement of the reference hot-path arithmetic (TEST INFRASTRUCTURE ONLY).

Plain functions on ndarrays, no autograd objects: each forward returns what the
reference stashes in ``out.cache`` and each backward takes exactly that. Arithmetic
follows the reference step by step (materialised float64 ``col`` buffer, ``np.dot``
GEMMs, overlap-add ``col2im``) so that (a) results agree with the reference to the
last few ulps and (b) timing this module is a fair stand-in for timing the reference's
NumPy path on a box where ``/root/reference`` does not exist.

All citations are relative to the reference repository root.
"""
import numpy as np

__all__ = [
    "unfold", "fold", "conv2d_fwd", "conv2d_bwd", "max_pool2d_fwd", "max_pool2d_bwd",
    "avg_pool2d_fwd", "avg_pool2d_bwd", "relu_fwd", "relu_bwd", "batch_norm_fwd",
    "batch_norm_bwd", "addmm_fwd", "addmm_bwd", "mm_bwd", "softmax", "cross_entropy_fwd",
    "cross_entropy_bwd", "mse_loss_fwd", "mse_loss_bwd", "sgd_step", "adam_step",
    "momentum_step", "rmsprop_step", "adagrad_step", "adadelta_step", "conv_out_size", "pool_out_size", "rel_err",
    "sigmoid_fwd", "sigmoid_bwd", "tanh_fwd", "tanh_bwd", "pad2d_fwd", "pad2d_bwd", "dropout_fwd", "dropout_bwd",
    "layer_norm_fwd", "layer_norm_bwd", "group_norm_fwd", "group_norm_bwd", "add_broadcast_bwd",
]


def _pair(v):
    return (int(v), int(v)) if isinstance(v, (int, float)) else (int(v[0]), int(v[1]))


# Checker-side model of the tf32 tensor-core mode (no reference counterpart): with TF32_OPERANDS = True every GEMM operand
# (col / W / dY / x) is first rounded to fp32 and cut to tf32's 10 mantissa bits, as tcgen05 kind::tf32 reads fp32 data
# (tools/tf32_probe.py: truncation); products and sums stay float64. "The reference's algorithm on tf32-truncated
# operands" is what the GPU's tf32 mode must reproduce -- it separates kernel error from the amplification of the tf32
# rounding itself by a deep, randomly initialised network (tools/parity_probe.py).
TF32_OPERANDS = False


def _gemm_in(a):
    if not TF32_OPERANDS:
        return a
    f = np.ascontiguousarray(a, dtype=np.float32)
    return (f.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32).astype(np.float64)


def conv_out_size(h, k, s, p):
    """xs/nn/functional.py:415-416 -- (H - k + 2p) // s + 1."""
    return (h - k + 2 * p) // s + 1


def pool_out_size(h, k, s, p):
    """xs/nn/functional.py:457-458 -- on the already padded extent: (H + 2p - k) // s + 1."""
    return (h + 2 * p - k) // s + 1


def rel_err(a, b):
    """SURVEY.md 8c error metric: max|a-b| / max|b| (b = oracle)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b)) if b.size else 0.0
    num = np.max(np.abs(a - b)) if b.size else 0.0
    return float(num / den) if den > 0 else float(num)


# --------------------------------------------------------------------------------------
# im2col / col2im  (xs/utils/common.py:35-46 and 49-65)
# --------------------------------------------------------------------------------------
def unfold(xp, oh, ow, kh, kw, stride):
    """Materialise the patch matrix. ``xp`` is the ALREADY padded input [N,C,Hp,Wp].

    Returns float64 ``col`` of shape [N*OH*OW, C*kh*kw]; rows ordered (n, oh, ow),
    columns ordered (c, ky, kx) -- xs/utils/common.py:35-46. The buffer is float64
    whatever the input dtype is (np.zeros default), which is why the reference's conv
    and pooling math runs in double precision.
    """
    sh, sw = stride
    n, c = xp.shape[:2]
    patches = np.zeros((n, c, kh, kw, oh, ow))
    for ky in range(kh):
        for kx in range(kw):
            patches[:, :, ky, kx] = xp[:, :, ky:ky + sh * oh:sh, kx:kx + sw * ow:sw]
    return patches.transpose(0, 4, 5, 1, 2, 3).reshape(n * oh * ow, c * kh * kw)


def fold(x_shape, pad, kh, kw, stride, dcol):
    """Overlap-add the patch-matrix gradient back to the input -- xs/utils/common.py:49-65.

    The scratch canvas is [N,C,H+2p+sh-1,W+2p+sw-1]; the result is the crop
    ``[p:H+p, p:W+p]``.
    """
    sh, sw = stride
    n, c, h, w = x_shape
    oh = (h + 2 * pad - kh) // sh + 1
    ow = (w + 2 * pad - kw) // sw + 1
    d6 = dcol.reshape(n, oh, ow, c, kh, kw).transpose(0, 3, 4, 5, 1, 2)
    canvas = np.zeros((n, c, h + 2 * pad + sh - 1, w + 2 * pad + sw - 1))
    for ky in range(kh):
        for kx in range(kw):
            canvas[:, :, ky:ky + sh * oh:sh, kx:kx + sw * ow:sw] += d6[:, :, ky, kx]
    return canvas[:, :, pad:h + pad, pad:w + pad]


def _zero_pad(x, p):
    if p == 0:
        return x
    return np.pad(x, ((0, 0), (0, 0), (p, p), (p, p)), "constant")


# --------------------------------------------------------------------------------------
# conv2d  (xs/nn/functional.py:405-442, xs/nn/grad_fn.py:186-203)
# --------------------------------------------------------------------------------------
def conv2d_fwd(x, w, b=None, stride=(1, 1), padding=0):
    """y = im2col(pad0(x)) . W^T (+ b) reshaped to NCHW; returns (y, col).

    x [N,C,H,W], w [OC,C,kh,kw], b [1,OC] or None. ``y`` takes x's dtype
    (functional.py:421), the product itself is float64 (col is float64).
    """
    stride = _pair(stride)
    n, _, h, wd = x.shape
    oc, _, kh, kw = w.shape
    oh = conv_out_size(h, kh, stride[0], padding)
    ow = conv_out_size(wd, kw, stride[1], padding)
    col = unfold(_zero_pad(x, padding), oh, ow, kh, kw, stride)
    flat = np.dot(_gemm_in(col), _gemm_in(w.reshape(oc, -1)).T)
    if b is not None:
        flat = flat + b
    y = np.empty((n, oc, oh, ow), dtype=x.dtype)
    y[...] = flat.reshape(n, oh, ow, oc).transpose(0, 3, 1, 2)
    return y, col


def conv2d_bwd(dy, x_shape, w, col, stride=(1, 1), padding=0, need_dx=True):
    """Returns (dx, dw, db) -- the three increments Conv2DBackward accumulates.

    dw = (dY^T . col) reshaped (grad_fn.py:191-194), db = sum dY over (0,2,3) (196-197),
    dx = col2im(dY . W) (199-203).
    """
    stride = _pair(stride)
    oc, c, kh, kw = w.shape
    g = dy.transpose(1, 0, 2, 3).reshape(oc, -1)
    db = np.sum(dy, axis=(0, 2, 3))
    g = _gemm_in(g)
    dw = np.dot(g, _gemm_in(col)).reshape(oc, c, kh, kw)
    # the dcol GEMM runs even when the input needs no gradient (grad_fn.py:199 sits
    # outside the `if`), only the col2im is skipped -- kept so CPU timings stay faithful
    dcol = g.T.dot(_gemm_in(w.reshape(oc, -1)))
    dx = fold(x_shape, padding, kh, kw, stride, dcol) if need_dx else None
    return dx, dw, db


# --------------------------------------------------------------------------------------
# max_pool2d  (xs/nn/functional.py:445-491, xs/nn/grad_fn.py:206-235)
# --------------------------------------------------------------------------------------
def max_pool2d_fwd(x, kernel_size=2, stride=None, padding=0):
    """Returns (y, argmax). The reference ALWAYS takes the im2col branch:
    zero padding (not -inf, Q2), windows flattened to rows ordered (n, oh, ow, c) with
    k*k entries ordered (ky, kx); ``argmax`` is int64 flat [N*OH*OW*C], first max wins.
    """
    k = int(kernel_size)
    stride = (k, k) if stride is None else _pair(stride)
    xp = _zero_pad(x, padding)
    n, c, hp, wp = xp.shape
    oh = (hp - k) // stride[0] + 1
    ow = (wp - k) // stride[1] + 1
    rows = unfold(xp, oh, ow, k, k, stride).reshape(-1, k * k)
    argmax = np.argmax(rows, axis=1)
    top = np.max(rows, axis=1)
    y = np.empty((n, c, oh, ow), dtype=x.dtype)
    y[...] = top.reshape(n, oh, ow, c).transpose(0, 3, 1, 2)
    return y, argmax


def max_pool2d_bwd(dy, argmax, x_shape, kernel_size, stride, padding):
    """dx: scatter dY to the arg-max slot of each window, then col2im (grad_fn.py:223-233)."""
    k = int(kernel_size)
    stride = (k, k) if stride is None else _pair(stride)
    g = dy.transpose(0, 2, 3, 1)
    slots = np.zeros((g.size, k * k))
    slots[np.arange(argmax.size), argmax.ravel()] = g.ravel()
    dcol = slots.reshape(g.shape[0] * g.shape[1] * g.shape[2], -1)
    return fold(x_shape, padding, k, k, stride, dcol)


# --------------------------------------------------------------------------------------
# avg_pool2d  (xs/nn/functional.py:539-583, xs/nn/grad_fn.py:284-309) -- padding == 0 only
# --------------------------------------------------------------------------------------
def avg_pool2d_fwd(x, kernel_size, stride=None, padding=0):
    assert padding == 0, "reference double-counts padding (functional.py:551-553); only p=0 is used"
    k = int(kernel_size)
    stride = (k, k) if stride is None else _pair(stride)
    n, c, h, w = x.shape
    oh = (h - k) // stride[0] + 1
    ow = (w - k) // stride[1] + 1
    rows = unfold(x, oh, ow, k, k, stride).reshape(-1, k * k)
    y = np.empty((n, c, oh, ow), dtype=x.dtype)
    y[...] = np.mean(rows, axis=1).reshape(n, oh, ow, c).transpose(0, 3, 1, 2)
    return y


def avg_pool2d_bwd(dy, x_shape, kernel_size, stride=None):
    k = int(kernel_size)
    stride = (k, k) if stride is None else _pair(stride)
    g = dy.transpose(0, 2, 3, 1)
    spread = np.repeat(g.ravel(), k * k) / (k * k)
    dcol = spread.reshape(g.shape[0] * g.shape[1] * g.shape[2], -1)
    return fold(x_shape, 0, k, k, stride, dcol)


# --------------------------------------------------------------------------------------
# relu  (xs/nn/td_functional.py:59-60, xs/nn/grad_fn.py:146-157)
# --------------------------------------------------------------------------------------
def relu_fwd(x):
    return np.maximum(0.0, x)


def relu_bwd(dy, x):
    """Gradient passes where x >= 0 (mask is x < 0, grad_fn.py:156)."""
    g = dy.copy()
    g[x < 0] = 0
    return g


# --------------------------------------------------------------------------------------
# batch_norm, training branch  (xs/nn/functional.py:676-740, xs/nn/grad_fn.py:379-411)
# --------------------------------------------------------------------------------------
def _bcast(v, ndim, axis):
    shp = [1] * ndim
    shp[axis] = -1
    return np.asarray(v).reshape(shp)


def batch_norm_fwd(x, gamma, beta, moving_mean, moving_var, axis=1, epsilon=1e-5, momentum=0.9):
    """Returns (y, xhat, sqrtvar, new_moving_mean, new_moving_var).

    Batch statistics over every axis but ``axis``; biased variance (np.var);
    running = momentum * running + (1 - momentum) * batch (functional.py:700-706, Q4).
    """
    red = tuple(i for i in range(x.ndim) if i != axis)
    mean = np.mean(x, axis=red, keepdims=True)
    var = np.var(x, axis=red, keepdims=True)
    sqrtvar = np.sqrt(var + epsilon)
    xhat = (x - mean) / sqrtvar
    y = _bcast(gamma, x.ndim, axis) * xhat + _bcast(beta, x.ndim, axis)
    new_mean = momentum * np.asarray(moving_mean) + (1 - momentum) * mean.ravel()
    new_var = momentum * np.asarray(moving_var) + (1 - momentum) * var.ravel()
    return y.astype(x.dtype, copy=False), xhat, sqrtvar, new_mean, new_var


def batch_norm_bwd(dy, xhat, sqrtvar, gamma, axis=1):
    """Returns (dx, dgamma, dbeta) -- grad_fn.py:385-411.

    dx = (M*dxhat - sum(dxhat) - xhat*sum(dxhat*xhat)) / (M*sqrtvar), dxhat = dy*gamma,
    M = number of reduced elements per channel.
    """
    red = tuple(i for i in range(dy.ndim) if i != axis)
    dgamma = np.sum(dy * xhat, axis=red)
    dbeta = np.sum(dy, axis=red)
    m = 1
    for a in red:
        m *= xhat.shape[a]
    dxhat = dy * _bcast(gamma, dy.ndim, axis)
    s1 = np.sum(dxhat, axis=red, keepdims=True)
    s2 = np.sum(dxhat * xhat, axis=red, keepdims=True)
    dx = (dxhat * m - s1 - xhat * s2) * (1.0 / (m * sqrtvar))
    return dx, dgamma, dbeta


# --------------------------------------------------------------------------------------
# Dense  (xs/nn/functional.py:81-120, xs/nn/grad_fn.py:73-99)
# --------------------------------------------------------------------------------------
def addmm_fwd(b, x, w):
    """y = x . W + b with W [in,out] (NOT transposed), b [1,out]."""
    return np.dot(_gemm_in(x), _gemm_in(w)) + b


def addmm_bwd(dy, x, w, b_shape):
    """Returns (dx, dw, db). db sums over the first mismatching axis, or is dy itself
    when N == 1 (grad_fn.py:89-95, Q12)."""
    mism = [k for k, v in enumerate(dy.shape) if v != b_shape[k]]
    db = np.sum(dy, axis=mism[0]) if mism else dy
    g = _gemm_in(dy)
    return np.dot(g, _gemm_in(w).T), np.dot(_gemm_in(x).T, g), db


def mm_bwd(dy, x, w):
    return np.dot(dy, w.T), np.dot(x.T, dy)


# --------------------------------------------------------------------------------------
# losses  (xs/nn/td_functional.py:73-78,115-128; functional.py:843-873,984-1010;
#          grad_fn.py:503-520,604-620)
# --------------------------------------------------------------------------------------
def softmax(z):
    e = np.exp(z - np.max(z, axis=-1, keepdims=True))
    return e / np.sum(e, axis=-1, keepdims=True)


def cross_entropy_fwd(logits, labels, reduction="mean"):
    """Returns (loss[1], probs). loss = |-(sum log p[i, y_i]) / N| (td_functional.py:119-127)."""
    probs = softmax(logits)
    rows = int(np.prod(logits.shape[:-1]))
    picked = np.log(probs).reshape(rows, -1)[np.arange(rows), labels.reshape(rows).astype(np.int64)]
    if reduction == "mean":
        loss = -np.sum(picked) / logits.shape[0]
    elif reduction == "sum":
        loss = -(-np.sum(picked))  # reference negates twice, then abs() (119-121,127)
    else:
        raise ValueError(reduction)
    return np.abs(np.asarray(loss, dtype=np.float64)).reshape(1), probs


def cross_entropy_bwd(probs, labels, dloss=1.0, reduction="mean"):
    """d logits = (p - onehot) / N * dloss (grad_fn.py:604-620)."""
    rows = int(np.prod(probs.shape[:-1]))
    g = probs.reshape(rows, -1).copy()
    g[np.arange(rows), labels.reshape(rows).astype(np.int64)] -= 1
    g = g.reshape(probs.shape)
    if reduction == "mean":
        g = g / probs.shape[0]
    return g * dloss


def mse_loss_fwd(a, b, reduction="mean"):
    """XS 'mean' = sum((a-b)^2) / 2 / numel (functional.py:857-860, Q9)."""
    d2 = np.power(a - b, 2)
    if reduction == "mean":
        return np.asarray(np.sum(d2) / 2 / np.prod(b.shape), dtype=np.float64).reshape(1)
    if reduction == "sum":
        return np.asarray(np.sum(d2), dtype=np.float64).reshape(1)
    return d2


def mse_loss_bwd(a, b, dloss=1.0, reduction="mean"):
    """d a = (a - b) * dloss (/ numel for 'mean'); d b = -d a (grad_fn.py:503-520)."""
    g = (a - b) * dloss
    if reduction == "mean":
        g = g / np.prod(a.shape)
    return g


# --------------------------------------------------------------------------------------
# optimizers  (xs/optim/sgd.py:5-9, xs/optim/adam.py:13-30)
# --------------------------------------------------------------------------------------
def sgd_step(params, grads, lr):
    """p -= lr * g, in place, per tensor; weight_decay is never applied (Q8)."""
    for p, g in zip(params, grads):
        p -= lr * g


def adam_step(params, grads, state, lr=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8):
    """state = {'t': int, 'v': [first moments], 'm': [second moments]} (names as in adam.py).

    v = b1 v + (1-b1) g ; m = b2 m + (1-b2) g^2 ; p -= lr * (v/(1-b1^t)) / (sqrt(m/(1-b2^t)) + eps).
    """
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    if "v" not in state:
        state["v"] = [np.zeros_like(p) for p in params]
        state["m"] = [np.zeros_like(p) for p in params]
    for i, (p, g) in enumerate(zip(params, grads)):
        v = beta1 * state["v"][i] + (1 - beta1) * g
        m = beta2 * state["m"][i] + (1 - beta2) * np.square(g)
        vc = v / (1 - pow(beta1, t))
        mc = m / (1 - pow(beta2, t))
        p -= lr * (vc / (np.sqrt(mc) + epsilon))
        state["v"][i] = v
        state["m"][i] = m


def _state(state, key, params):
    if key not in state:
        state[key] = [np.zeros_like(p) for p in params]
    return state[key]


def momentum_step(params, grads, state, lr=0.01, rho=0.9):
    """xs/optim/sgd.py:12-29: v = rho v + (1-rho) g ; p -= lr v."""
    v = _state(state, "v", params)
    for i, (p, g) in enumerate(zip(params, grads)):
        v[i] = rho * v[i] + (1 - rho) * g
        p -= lr * v[i]


def rmsprop_step(params, grads, state, lr=0.001, rho=0.9, epsilon=1e-7):
    """xs/optim/rmsprop.py:4-21: s = rho s + (1-rho) g^2 ; p -= lr g / sqrt(s + eps)."""
    s = _state(state, "s", params)
    for i, (p, g) in enumerate(zip(params, grads)):
        s[i] = rho * s[i] + (1 - rho) * np.square(g)
        p -= lr * g / np.sqrt(s[i] + epsilon)


def adagrad_step(params, grads, state, lr=0.01, epsilon=1e-7):
    """xs/optim/adagrad.py:4-18: s += g^2 ; p -= lr g / sqrt(s + eps)."""
    s = _state(state, "s", params)
    for i, (p, g) in enumerate(zip(params, grads)):
        s[i] = s[i] + np.power(g, 2)
        p -= lr * g / np.sqrt(s[i] + epsilon)


def adadelta_step(params, grads, state, rho=0.95, epsilon=1e-7):
    """xs/optim/adadelta.py:4-26 (lr is accepted by the reference but unused)."""
    s, x = _state(state, "s", params), _state(state, "x", params)
    for i, (p, g) in enumerate(zip(params, grads)):
        s[i] = rho * s[i] + (1 - rho) * np.power(g, 2)
        d = np.sqrt((x[i] + epsilon) / (s[i] + epsilon)) * g
        p -= d
        x[i] = rho * x[i] + (1 - rho) * np.power(d, 2)


# --------------------------------------------------------------------------------------
# SURVEY 8f-4 ops around the hot path
# --------------------------------------------------------------------------------------
def sigmoid_fwd(x):
    """xs/nn/td_functional.py:63-66."""
    return 1.0 / (1.0 + np.exp(-x))


def sigmoid_bwd(dy, y):
    """xs/nn/grad_fn.py:161-164: dY * y * (1 - y)."""
    return dy * y * (1 - y)


def tanh_fwd(x):
    """xs/nn/td_functional.py:69-70."""
    return np.tanh(x)


def tanh_bwd(dy, y):
    """xs/nn/grad_fn.py:167-170: dY * (1 - y^2)."""
    return dy * (1 - np.square(y))


def pad2d_fwd(x, padding):
    """xs/nn/functional.py:637-638."""
    return np.pad(x, ((0, 0), (0, 0), (padding[0], padding[0]), (padding[1], padding[1])), "constant")


def pad2d_bwd(dy, padding):
    """xs/nn/grad_fn.py:312-316 (the interior of dY; the shipped slice [p:-p] is empty for p = 0)."""
    ph, pw = padding
    return dy[:, :, ph:dy.shape[2] - ph, pw:dy.shape[3] - pw]


def dropout_fwd(x, mask, keep_prob):
    """xs/nn/functional.py:658-661, given the 0/1 mask: x * mask / keep_prob."""
    return x * mask / keep_prob


def dropout_bwd(dy, mask, keep_prob):
    """xs/nn/grad_fn.py:319-325 as written: mask * keep_prob * dY (it multiplies by keep_prob on the way back; the shipped
    in-place multiply of the int64 mask raises a casting error, so this formula is pinned by reading, not by running)."""
    return mask * keep_prob * dy


def layer_norm_fwd(x, gamma, beta, epsilon=1e-10):
    """xs/nn/functional.py:743-775 as intended (the shipped code raises at line 754: `np.subtract(x, mean, out=mean)` with a
    keepdims mean -- PARITY UNPINNED for this op): statistics over every axis but 0, biased variance, per-element affine.
    Returns (y, xhat, std)."""
    axes = tuple(range(1, x.ndim))
    xmu = x - np.mean(x, axis=axes, keepdims=True)
    std = np.sqrt(np.var(x, axis=axes, keepdims=True) + epsilon)
    xhat = xmu / std
    return gamma * xhat + beta, xhat, std


def layer_norm_bwd(dy, xhat, std, gamma):
    """xs/nn/grad_fn.py:414-447: dgamma = sum_n dY*xhat, dbeta = sum_n dY; dX by the dvar / dmean chain (written out)."""
    axes = tuple(range(1, dy.ndim))
    dgamma, dbeta = np.sum(dy * xhat, axis=0), np.sum(dy, axis=0)
    g = dy * gamma
    n = np.prod(dy.shape[1:])
    std_inv = 1.0 / std
    xmu = xhat * std
    dvar = -0.5 * np.sum(g * xmu, axis=axes, keepdims=True) * std_inv ** 3
    dmean = -np.sum(g * std_inv, axis=axes, keepdims=True) - 2.0 * dvar * np.mean(xmu, axis=axes, keepdims=True)
    return g * std_inv + (2.0 / n) * dvar * xmu + dmean / n, dgamma, dbeta


def group_norm_fwd(x, gamma, beta, epsilon=1e-5, groups=16):
    """xs/nn/functional.py:778-812 as intended (the shipped code raises at line 789 like layer_norm -- PARITY UNPINNED):
    statistics per (sample, group) over (c/groups, h, w), per-channel affine. Returns (y, xhat [N,C,H,W], std [N,G,1,1,1])."""
    n, c, h, w = x.shape
    d = x.reshape(n, groups, c // groups, h, w)
    xmu = d - np.mean(d, axis=(2, 3, 4), keepdims=True)
    std = np.sqrt(np.var(d, axis=(2, 3, 4), keepdims=True) + epsilon)
    xhat = (xmu / std).reshape(n, c, h, w)
    return gamma.reshape(1, c, 1, 1) * xhat + beta.reshape(1, c, 1, 1), xhat, std


def group_norm_bwd(dy, xhat, std, gamma, groups):
    """xs/nn/grad_fn.py:450-490: dgamma / dbeta over (n, h, w); dX by the per-group dvar / dmean chain."""
    n, c, h, w = dy.shape
    dgamma, dbeta = np.sum(dy * xhat, axis=(0, 2, 3)), np.sum(dy, axis=(0, 2, 3))
    g = (dy * gamma.reshape(1, c, 1, 1)).reshape(n, groups, c // groups, h, w)
    xh = xhat.reshape(n, groups, c // groups, h, w)
    m = c // groups * h * w
    std_inv = 1.0 / std
    xmu = xh * std
    dvar = -0.5 * std_inv ** 3 * np.sum(g * xmu, axis=(2, 3, 4), keepdims=True)
    dmean = np.sum(-g * std_inv, axis=(2, 3, 4), keepdims=True) + dvar * -2.0 / m * np.sum(xmu, axis=(2, 3, 4), keepdims=True)
    dx = g * std_inv + dvar * 2.0 / m * xmu + dmean / m
    return dx.reshape(n, c, h, w), dgamma, dbeta


def add_broadcast_bwd(dy, shape):
    """xs/nn/grad_fn.py:6-16: an operand whose shape differs gets dY summed over the FIRST mismatching axis."""
    mism = [k for k, v in enumerate(dy.shape) if k >= len(shape) or v != shape[k]]
    return np.sum(dy, axis=mism[0]) if mism else dy
