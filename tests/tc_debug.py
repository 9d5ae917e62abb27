"""GPU debug/diagnostic script for the tcgen05 conv kernels: prints rel. errors vs the oracle per shape.
Run under `timeout` (a broken mbarrier pipeline hangs instead of failing)."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), HERE]

import xshinnosuke_b200 as xs  # noqa: E402
from oracle import xs_numpy as X  # noqa: E402
from xshinnosuke_b200 import nn  # noqa: E402
from xshinnosuke_b200._lib import lib  # noqa: E402
from xshinnosuke_b200.nn import functional as F  # noqa: E402
import ctypes  # noqa: E402


def stats():
    buf = (ctypes.c_uint64 * 2)()
    lib.xsb_tc_stats(ctypes.addressof(buf))
    return buf[0], buf[1]


def one(n, c, h, oc, k, s, p, which="all"):
    rs = np.random.RandomState(n * 100 + c + h)
    x = rs.randn(n, c, h, h).astype(np.float32)
    w = (rs.randn(oc, c, k, k) * 0.1).astype(np.float32)
    b = rs.randn(1, oc).astype(np.float32)
    yr, col = X.conv2d_fwd(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), (s, s), p)
    g = rs.randn(*yr.shape).astype(np.float32)
    dxr, dwr, dbr = X.conv2d_bwd(g.astype(np.float64), x.shape, w.astype(np.float64), col, (s, s), p)
    out = {}
    for mode in ("fp32", "tf32"):
        xs.set_math_mode(mode)
        s0 = stats()
        tx, tw, tb = xs.tensor(x, requires_grad=True), nn.Parameter(w), nn.Parameter(b)
        y = F.conv2d(tx, tw, tb, (s, s), p)
        ey = X.rel_err(y.eval, yr)
        y.backward(gradients=g)
        s1 = stats()
        out[mode] = (ey, X.rel_err(tx.grad.eval, dxr), X.rel_err(tw.grad.eval, dwr), s1[0] - s0[0], s1[1] - s0[1])
    print(f"N{n} C{c} H{h} OC{oc} k{k} s{s} p{p}: " + " | ".join(
        f"{m}: y {v[0]:.2e} dx {v[1]:.2e} dw {v[2]:.2e} (tc {v[3]}, ffma {v[4]})" for m, v in out.items()), flush=True)
    return out["tf32"]


if __name__ == "__main__":
    xs.runtime.ensure()
    t0 = time.time()
    shapes = [(2, 32, 8, 64, 3, 1, 1), (4, 64, 14, 64, 3, 1, 1), (3, 64, 9, 128, 3, 1, 1), (8, 64, 28, 128, 3, 2, 1),
              (8, 64, 28, 128, 1, 2, 0), (4, 128, 14, 256, 3, 1, 1), (2, 256, 7, 512, 3, 1, 1), (5, 96, 11, 192, 1, 1, 0),
              (2, 64, 12, 64, 5, 1, 2), (16, 64, 56, 64, 3, 1, 1)]
    if len(sys.argv) > 1:
        shapes = shapes[: int(sys.argv[1])]
    worst = 0.0
    for sh in shapes:
        r = one(*sh)
        worst = max(worst, r[0], r[1], r[2])
    print(f"worst tf32 rel err {worst:.3e}  ({time.time() - t0:.1f}s)")
