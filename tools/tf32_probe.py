"""How does tcgen05.mma.kind::tf32 reduce fp32 operands to tf32 -- truncation or rounding? (decides how the 3xTF32
split of XSB_MATH_FP32X3 forms its `hi` parts). Runs the tcgen05 Dense kernel on x.W with raw / truncated / RNE-rounded
operands and reports which pre-processing leaves the result bit-identical to the raw one."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xshinnosuke_b200._lib import MATH_TF32, lib  # noqa: E402
from xshinnosuke_b200.core.device import DevArray, rt  # noqa: E402


def trunc(a):
    return (a.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def rne(a):
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0xFFF + ((u >> 13) & 1)) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def gemm(x, w):
    m, k = x.shape
    n = w.shape[1]
    dx, dw, out = DevArray.from_numpy(x), DevArray.from_numpy(w), DevArray.empty((m, n))
    lib.xsb_gemm(dx.ptr, dw.ptr, out.wptr, None, m, n, k, 0, 0, 0, MATH_TF32, rt.ws_big.data_ptr(), rt.WS_BIG_FLOATS, rt.stream())
    return out.to_numpy()


rt.ensure()
rs = np.random.RandomState(0)
x, w = rs.randn(128, 256).astype(np.float32), rs.randn(256, 128).astype(np.float32)
y = gemm(x, w)
yt, yr = gemm(trunc(x), trunc(w)), gemm(rne(x), rne(w))
ref = x.astype(np.float64) @ w.astype(np.float64)
print("raw == truncated operands bitwise:", np.array_equal(y, yt), " raw == RNE operands bitwise:", np.array_equal(y, yr))
print("rel err raw %.3e  trunc %.3e  rne %.3e" % tuple(np.abs(v - ref).max() / np.abs(ref).max() for v in (y, yt, yr)))
# exactness of the fp32 accumulation: products of tf32-exact operands, long K
xh, wh = trunc(rs.randn(128, 4608).astype(np.float32)), trunc(rs.randn(4608, 128).astype(np.float32))
ya = gemm(xh, wh)
refa = xh.astype(np.float64) @ wh.astype(np.float64)
print("tf32-exact operands, K=4608: rel err %.3e (fp32 accumulation only)" % (np.abs(ya - refa).max() / np.abs(refa).max()))
