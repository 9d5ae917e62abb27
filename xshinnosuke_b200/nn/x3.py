"""fp32-accurate GEMM-shaped ops ON the tensor cores: the 3xTF32 split ("fp32x3" math mode).

The reference computes in float64 (xs/utils/common.py:37,56) and north_star holds fp32 results to 1e-5; tcgen05 has no fp32
MMA and kind::tf32 alone gives ~5e-4. Every operand is split x ~ hi + lo with hi = x rounded to tf32 and lo = (x - hi)
rounded to tf32 (xsb_split_tf32); a product becomes

    x . w  =  lo . hi  +  hi . lo  +  hi . hi        (lo . lo and the rounding of lo: ~2^-22 relative, unbiased, dropped)

three accumulating tcgen05 launches of the ordinary tf32 kernels (small terms first, bias with the last). Both parts are
exact in tf32, so the tensor core's operand truncation changes nothing and every product is exact in fp32; what remains
is the fp32 accumulation in TMEM (measured 4e-7 at K = 4608, tools/tf32_probe.py).
Shapes outside the tensor-core envelope run the fp32 FFMA kernel once instead.
"""
from .._lib import MATH_FP32, MATH_TF32, lib
from ..core.device import rt

MATH_FP32X3 = 2


def conv_ok(c, oc):
    """Envelope of the tcgen05 convolution kernels (conv_tc.cu: fprop_ok / stem_ok)."""
    return oc % 64 == 0 and (c % 32 == 0 or c <= 4)


def dgrad_ok(c, oc, kh, kw, sh, sw):
    return oc % 32 == 0 and c % 64 == 0 and kh * kw <= 64 and sh <= 8 and sw <= 8


def gemm_ok(*pitches):
    """TMA sources need row pitches that are multiples of 16 bytes (gemm_tc.cuh)."""
    return all(p % 4 == 0 for p in pitches)


def split(ptr, n):
    """-> (hi, lo) device buffers of n floats."""
    hi, lo = rt.empty(n), rt.empty(n)
    lib.xsb_split_tf32(ptr, hi.data_ptr(), lo.data_ptr(), n, rt.stream())
    rt.launches += 1
    return hi, lo


def _three(call, a, b, acc0):
    """call(a_ptr, b_ptr, accumulate, last) for the three products of (a_hi + a_lo) . (b_hi + b_lo)."""
    (a_hi, a_lo), (b_hi, b_lo) = a, b
    call(a_lo.data_ptr(), b_hi.data_ptr(), acc0, False)
    call(a_hi.data_ptr(), b_lo.data_ptr(), 1, False)
    call(a_hi.data_ptr(), b_hi.data_ptr(), 1, True)
    rt.launches += 3


def _ws():
    return rt.ws_big.data_ptr(), rt.WS_BIG_FLOATS


def conv_fwd(xs_, ws_, bias_ptr, y_ptr, geom, accumulate=0):
    ws, wsn = _ws()
    _three(lambda a, b, acc, last: lib.xsb_conv2d_fwd_acc(a, b, bias_ptr if last else None, y_ptr, *geom, acc, MATH_TF32, ws, wsn,
                                                          rt.stream()), xs_, ws_, accumulate)


def conv_dgrad(dys, ws_, dx_ptr, geom, accumulate):
    ws, wsn = _ws()
    _three(lambda a, b, acc, last: lib.xsb_conv2d_dgrad(a, b, dx_ptr, *geom, acc, MATH_TF32, ws, wsn, rt.stream()), dys, ws_,
           accumulate)


def conv_wgrad(dys, xs_, dw_ptr, geom, accumulate):
    ws, wsn = _ws()
    _three(lambda a, b, acc, last: lib.xsb_conv2d_wgrad(a, b, dw_ptr, *geom, acc, MATH_TF32, ws, wsn, rt.stream()), dys, xs_,
           accumulate)


def gemm(a_ptr, a_n, b_ptr, b_n, c_ptr, bias_ptr, M, N, K, ta, tb, accumulate):
    ws, wsn = _ws()
    a, b = split(a_ptr, a_n), split(b_ptr, b_n)
    _three(lambda p, q, acc, last: lib.xsb_gemm(p, q, c_ptr, bias_ptr if last else None, M, N, K, ta, tb, acc, MATH_TF32, ws,
                                                wsn, rt.stream()), a, b, accumulate)


def c_mode(ok):
    """The math_mode a single C-ABI call gets while the host is in fp32x3 mode and the shape is NOT split."""
    return MATH_FP32 if rt.math_mode == MATH_FP32X3 else rt.math_mode
