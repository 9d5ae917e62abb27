"""GPU (B200): parity of the CUDA path with the reference, through the public API -> ctypes -> C ABI.

fp32 math mode: outputs and gradients within 1e-5 (max|a-b|/max|b|) of the reference's NumPy path;
max-pool values, arg-max indices and all shapes bit-exact (BASELINE.json north_star)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(autouse=True)
def _fp32_mode():
    import xshinnosuke_b200 as xs
    import xshinnosuke_b200.core.state as G
    from xshinnosuke_b200._lib import lib
    lib.load()                      # fail loudly if the CUDA library is missing
    xs.runtime.ensure()
    xs.set_math_mode("fp32")
    G.INPUTS = G.OUTPUTS = G.GRAPH = None
    yield
    G.INPUTS = G.OUTPUTS = G.GRAPH = None


def test_library_is_the_cuda_one():
    import xshinnosuke_b200 as xs
    from xshinnosuke_b200._lib import LIB_PATH, lib
    import os
    assert os.path.exists(LIB_PATH)
    assert type(lib).__name__ == "_Lib"
    assert xs.runtime.device.type == "cuda"
    assert lib.xsb_device_sm_count() >= 100


def test_conv(golden):
    import parity_suite as S
    S.check_conv(golden, TOL)


def test_pools(golden):
    import parity_suite as S
    S.check_maxpool(golden, TOL)
    S.check_avgpool(golden, TOL)


def test_relu_add(golden):
    import parity_suite as S
    S.check_relu_add(golden, TOL)
    S.check_inplace_relu_layer(TOL)


def test_postponed_shortcut_dgrad():
    """fp32 FFMA kernels, then the tcgen05 path (parity-class dgrad: the postponed 1x1/2 skips its zero-fill launches)."""
    import parity_suite as S
    import xshinnosuke_b200 as xs
    S.check_postponed_shortcut_dgrad(TOL)
    xs.set_math_mode("tf32")
    try:
        S.check_postponed_shortcut_dgrad(2e-3, c=64, oc=128, hw=16)
    finally:
        xs.set_math_mode("fp32")


def test_bn(golden):
    import parity_suite as S
    S.check_bn(golden, TOL)
    S.check_bn_eval(TOL)


def test_dense_losses(golden):
    import parity_suite as S
    S.check_dense_losses(golden, TOL)


def test_optimizers(golden):
    import parity_suite as S
    S.check_optimizers(golden, TOL)
    S.check_weight_decay(golden, TOL)


def test_gradient_check_scripts(golden):
    import parity_suite as S
    S.check_gradient_check_scripts(golden, TOL)


def test_trajectories(golden):
    import parity_suite as S
    S.check_test_fc(golden, TOL)
    S.check_test_conv(golden, TOL, steps=50)


@pytest.mark.parametrize("tag,n,hw", [("n2_33", 2, 33), ("n4_56", 4, 56), ("n32_56", 32, 56)])
def test_resnet18(golden, tag, n, hw):
    import parity_suite as S
    S.check_resnet18(golden, tag, n, hw, tol_logits=5e-5, tol_loss=2e-5, tail_rtol=0.1 if n == 32 else None)


def test_fit_static_graph(golden):
    import parity_suite as S
    S.check_fit_configs(golden, TOL)


# ------------------------------------------------------------------ sizes beyond the golden file: vs the oracle
@pytest.mark.parametrize("n,c,h,oc,k,s,p", [(8, 64, 28, 64, 3, 1, 1), (8, 64, 28, 128, 3, 2, 1), (8, 64, 28, 128, 1, 2, 0),
                                           (4, 3, 56, 64, 7, 2, 3), (16, 256, 7, 512, 3, 1, 1), (5, 20, 9, 12, 3, 1, 1)])
def test_conv_vs_oracle(n, c, h, oc, k, s, p):
    import xshinnosuke_b200 as xs
    from oracle import xs_numpy as X
    from xshinnosuke_b200 import nn
    from xshinnosuke_b200.nn import functional as F
    rs = np.random.RandomState(n * 1000 + c)
    x = rs.randn(n, c, h, h).astype(np.float32)
    w = (rs.randn(oc, c, k, k) * 0.1).astype(np.float32)
    b = rs.randn(1, oc).astype(np.float32)
    tx, tw, tb = xs.tensor(x, requires_grad=True), nn.Parameter(w), nn.Parameter(b)
    y = F.conv2d(tx, tw, tb, (s, s), p)
    yr, col = X.conv2d_fwd(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), (s, s), p)
    assert X.rel_err(y.eval, yr) <= TOL
    g = rs.randn(*yr.shape).astype(np.float32)
    y.backward(gradients=g)
    dx, dw, db = X.conv2d_bwd(g.astype(np.float64), x.shape, w.astype(np.float64), col, (s, s), p)
    assert X.rel_err(tx.grad.eval, dx) <= TOL
    assert X.rel_err(tw.grad.eval, dw) <= TOL
    assert X.rel_err(tb.grad.eval.reshape(-1), db) <= TOL


@pytest.mark.parametrize("n,c,h,k,s,p", [(32, 64, 28, 3, 2, 1), (256, 8, 27, 2, 2, 0), (4, 64, 112, 3, 2, 1)])
def test_maxpool_vs_oracle(n, c, h, k, s, p):
    import xshinnosuke_b200 as xs
    from oracle import xs_numpy as X
    from xshinnosuke_b200.nn import functional as F
    rs = np.random.RandomState(n + c + h)
    x = np.maximum(rs.randn(n, c, h, h), 0).astype(np.float32)   # post-ReLU: ~50 % exact zeros -> ties
    tx = xs.tensor(x, requires_grad=True)
    y = F.max_pool2d(tx, k, s, p)
    yr, am = X.max_pool2d_fwd(x, k, s, p)
    assert np.array_equal(y.eval, yr)
    assert np.array_equal(np.asarray(y.cache["pool_argmax"]), am)
    g = rs.randn(*yr.shape).astype(np.float32)
    y.backward(gradients=g)
    assert X.rel_err(tx.grad.eval, X.max_pool2d_bwd(g, am, x.shape, k, s, p)) <= 1e-6


def test_bn_large_vs_oracle():
    import xshinnosuke_b200 as xs
    from oracle import xs_numpy as X
    from xshinnosuke_b200 import nn
    from xshinnosuke_b200.nn import functional as F
    rs = np.random.RandomState(3)
    x = (rs.randn(32, 64, 28, 28) * 2 + 1).astype(np.float32)
    gm, bt = (rs.rand(64) + 0.5).astype(np.float32), rs.randn(64).astype(np.float32)
    tx, tg, tb = xs.tensor(x, requires_grad=True), nn.Parameter(gm), nn.Parameter(bt)
    mm, mv = xs.zeros(64), xs.ones(64)
    y = F.batch_norm(tx, tg, tb, mm, mv, 1, True, 1e-6, 0.99)
    yr, xhat, sv, mm1, mv1 = X.batch_norm_fwd(x.astype(np.float64), gm, bt, np.zeros(64), np.ones(64), 1, 1e-6, 0.99)
    assert X.rel_err(y.eval, yr) <= TOL
    assert X.rel_err(mm.eval, mm1) <= TOL and X.rel_err(mv.eval, mv1) <= TOL
    g = rs.randn(*x.shape).astype(np.float32)
    y.backward(gradients=g)
    dx, dg, db = X.batch_norm_bwd(g.astype(np.float64), xhat, sv, gm, 1)
    assert X.rel_err(tx.grad.eval, dx) <= TOL
    assert X.rel_err(tg.grad.eval, dg) <= TOL and X.rel_err(tb.grad.eval, db) <= TOL


def test_no_silent_fallback_for_tensor_core_mode():
    """Asking for a math mode whose kernel is not built must raise, never fall back."""
    import xshinnosuke_b200 as xs
    from xshinnosuke_b200._lib import lib
    from xshinnosuke_b200.core.device import rt
    a = xs.tensor(np.ones((4, 4), dtype=np.float32))
    try:
        lib.xsb_gemm(a.arr.ptr, a.arr.ptr, a.arr.ptr, None, 4, 4, 4, 0, 0, 0, 7, None, 0, rt.stream())
    except RuntimeError:
        return
    # mode 7 does not exist: reaching here means a silent fallback
    raise AssertionError("unknown math mode was accepted")


# ------------------------------------------------------------------ tf32 tensor-core mode (tcgen05): <= 2e-3
TOL_TC = 2e-3


def _tc_stats():
    import ctypes
    from xshinnosuke_b200._lib import lib
    buf = (ctypes.c_uint64 * 2)()
    lib.xsb_tc_stats(ctypes.addressof(buf))
    return buf[0], buf[1]


@pytest.mark.parametrize("n,c,h,oc,k,s,p,n_tc", [
    (4, 64, 14, 64, 3, 1, 1, 3), (8, 64, 28, 128, 3, 2, 1, 3), (8, 64, 28, 128, 1, 2, 0, 3), (4, 64, 15, 128, 3, 2, 1, 3), (4, 128, 14, 256, 3, 1, 1, 3),
    (2, 256, 7, 512, 3, 1, 1, 3), (2, 64, 12, 64, 5, 1, 2, 3), (16, 64, 56, 64, 3, 1, 1, 3), (3, 32, 9, 64, 3, 1, 1, 2),
    (2, 3, 20, 64, 7, 2, 3, 2), (4, 3, 56, 64, 7, 2, 3, 2), (2, 1, 28, 64, 2, 1, 0, 2), (3, 3, 33, 128, 3, 1, 1, 2),
    (2, 4, 19, 64, 5, 2, 2, 2), (2, 3, 20, 8, 3, 1, 1, 0),
    # strided dgrad = one stride-1 conv per parity class (5x5/2; trailing row without gradient; classes no tap reaches)
    (2, 64, 16, 64, 5, 2, 2, 3), (2, 64, 10, 64, 3, 2, 0, 3), (2, 64, 8, 64, 2, 3, 0, 3), (2, 64, 17, 128, 7, 2, 3, 3),
    # filter-resident halo kernel (fwd and stride-1 dgrad), incl. partial tiles and zero padding via TMA OOB fill
    (2, 64, 32, 64, 3, 1, 1, 3), (2, 64, 30, 64, 3, 1, 2, 3), (3, 32, 34, 128, 3, 1, 0, 2), (150, 64, 16, 64, 3, 1, 1, 3),
    # wave-aware tiling: fewer tiles than SMs -> column parts (incl. unequal 192/160/160); full wave + split tail wave
    (24, 256, 7, 512, 3, 1, 1, 3), (100, 128, 14, 256, 3, 1, 1, 3)])
def test_conv_tf32_tensor_cores(n, c, h, oc, k, s, p, n_tc):
    """fwd / dgrad / wgrad on tcgen05 (kind::tf32) against the float64 oracle; `n_tc` = how many of the three
    must have run on the tensor cores for this shape (C <= 4 stems: fwd + wgrad via the padded-4-channel kernels,
    their dgrad and channel counts off the 32/64 grid run the fp32 FFMA kernel)."""
    import xshinnosuke_b200 as xs
    from oracle import xs_numpy as X
    from xshinnosuke_b200 import nn
    from xshinnosuke_b200.nn import functional as F
    xs.set_math_mode("tf32")
    try:
        rs = np.random.RandomState(n * 1000 + c + h)
        x = rs.randn(n, c, h, h).astype(np.float32)
        w = (rs.randn(oc, c, k, k) * 0.1).astype(np.float32)
        b = rs.randn(1, oc).astype(np.float32)
        s0 = _tc_stats()
        tx, tw, tb = xs.tensor(x, requires_grad=True), nn.Parameter(w), nn.Parameter(b)
        y = F.conv2d(tx, tw, tb, (s, s), p)
        yr, col = X.conv2d_fwd(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), (s, s), p)
        assert X.rel_err(y.eval, yr) <= TOL_TC
        g = rs.randn(*yr.shape).astype(np.float32)
        y.backward(gradients=g)
        dx, dw, db = X.conv2d_bwd(g.astype(np.float64), x.shape, w.astype(np.float64), col, (s, s), p)
        assert X.rel_err(tx.grad.eval, dx) <= TOL_TC
        s1 = _tc_stats()       # after dX was read: the dgrad of a filter smaller than its stride is postponed until then
        assert X.rel_err(tw.grad.eval, dw) <= TOL_TC
        assert X.rel_err(tb.grad.eval.reshape(-1), db) <= 1e-5
        assert s1[0] - s0[0] == n_tc and (s1[0] - s0[0]) + (s1[1] - s0[1]) == 3
    finally:
        xs.set_math_mode("fp32")


@pytest.mark.parametrize("m,n,k,on_tc", [
    (128, 500, 784, True),      # BASELINE configs[1]: Dense(784 -> 500), batch 128 (K tail of 16, N tail of 116)
    (128, 100, 4608, True),     # ResNet-18 head (split-K)
    (128, 10, 500, False),      # Dense(10): 40-byte rows cannot be a TMA source -> FFMA kernel
    (77, 36, 40, True), (300, 260, 200, True), (1, 16, 8, True), (130, 132, 36, True)])
def test_dense_gemm_tf32_tensor_cores(m, n, k, on_tc):
    """xsb_gemm in tf32 mode = the tcgen05 kernel for all three Dense products (y = x.W + b, dX = dY.W^T, dW = x^T.dY,
    each with and without "+="), ragged M / N / K served by TMA zero fill and clipped stores; float64 NumPy is the
    oracle (xs/nn/functional.py:81-120, xs/nn/grad_fn.py:73-99)."""
    from xshinnosuke_b200._lib import MATH_TF32, lib
    from xshinnosuke_b200.core.device import DevArray, rt
    from oracle import xs_numpy as X
    rs = np.random.RandomState(m + n + k)
    x, w = rs.randn(m, k).astype(np.float32), (rs.randn(k, n) * 0.1).astype(np.float32)
    b, dy = rs.randn(n).astype(np.float32), rs.randn(m, n).astype(np.float32)
    c0y, c0x, c0w = rs.randn(m, n).astype(np.float32), rs.randn(m, k).astype(np.float32), rs.randn(k, n).astype(np.float32)
    dx_, dw_, db_, ddy = DevArray.from_numpy(x), DevArray.from_numpy(w), DevArray.from_numpy(b), DevArray.from_numpy(dy)
    x64, w64, dy64 = x.astype(np.float64), w.astype(np.float64), dy.astype(np.float64)

    def run(a, bm, c_init, bias, M, N, K, ta, tb, acc):
        out = DevArray.from_numpy(c_init.copy())
        lib.xsb_gemm(a.ptr, bm.ptr, out.wptr, bias.ptr if bias is not None else None, M, N, K, ta, tb, acc, MATH_TF32,
                     rt.ws_big.data_ptr(), rt.WS_BIG_FLOATS, rt.stream())
        return out.to_numpy()

    s0 = _tc_stats()
    for acc in (0, 1):
        y = run(dx_, dw_, c0y, db_, m, n, k, 0, 0, acc)
        assert X.rel_err(y, x64 @ w64 + b + acc * c0y) <= TOL_TC
        gx = run(ddy, dw_, c0x, None, m, k, n, 0, 1, acc)
        assert X.rel_err(gx, dy64 @ w64.T + acc * c0x) <= TOL_TC
        gw = run(dx_, ddy, c0w, None, k, n, m, 1, 0, acc)
        assert X.rel_err(gw, x64.T @ dy64 + acc * c0w) <= TOL_TC
    s1 = _tc_stats()
    if on_tc:
        assert s1[0] - s0[0] == 6 and s1[1] == s0[1]
    else:
        assert s1[0] - s0[0] < 6 and (s1[0] - s0[0]) + (s1[1] - s0[1]) == 6


@pytest.mark.parametrize("n,c,h,oc,k,s,p", [
    (4, 64, 14, 64, 3, 1, 1), (8, 64, 28, 128, 3, 2, 1), (2, 256, 7, 512, 3, 1, 1), (3, 3, 56, 64, 7, 2, 3),
    (5, 64, 32, 64, 3, 1, 1), (4, 64, 30, 64, 3, 1, 2), (2, 128, 9, 256, 1, 2, 0), (150, 64, 16, 64, 3, 1, 1),
    (24, 256, 7, 512, 3, 1, 1), (100, 128, 14, 256, 3, 1, 1)])
def test_conv_bn_relu_fused_statistics_tf32(n, c, h, oc, k, s, p):
    """Conv2D -> BatchNormalization -> ReLU(inplace) in tf32 mode: the batch statistics come from the partials the conv
    epilogue gathers (xsb_conv2d_fwd_stats + xsb_bn_finalize, no pass over the conv output) and, backward, the conv
    bias gradient from the column sums the BN backward gathers while writing dX (xsb_bn_bwd_dxsum). Checked against
    the float64 oracle chain; statistics additionally against the conv output the GPU itself produced (<= 1e-5)."""
    import xshinnosuke_b200 as xs
    from oracle import xs_numpy as X
    from xshinnosuke_b200 import nn
    from xshinnosuke_b200._lib import lib
    from xshinnosuke_b200.nn import functional as F
    xs.set_math_mode("tf32")
    try:
        rs = np.random.RandomState(7 * n + c + h)
        x = rs.randn(n, c, h, h).astype(np.float32)
        w = (rs.randn(oc, c, k, k) * 0.1).astype(np.float32)
        b = rs.randn(1, oc).astype(np.float32)
        gm, bt = (rs.rand(oc) + 0.5).astype(np.float32), rs.randn(oc).astype(np.float32)
        tx, tw, tb = xs.tensor(x, requires_grad=c >= 32), nn.Parameter(w), nn.Parameter(b)
        tg, tbt = nn.Parameter(gm), nn.Parameter(bt)
        mm, mv = xs.tensor(np.zeros(oc, dtype=np.float32)), xs.tensor(np.ones(oc, dtype=np.float32))
        calls = []
        orig_stats, orig_fin, orig_dxs = lib.xsb_conv2d_fwd_stats, lib.xsb_bn_finalize, lib.xsb_bn_bwd_dxsum
        orig_bst = lib.xsb_bn_stats
        lib.xsb_conv2d_fwd_stats = lambda *a: (calls.append("stats"), orig_stats(*a))[1]
        lib.xsb_bn_finalize = lambda *a: (calls.append("finalize"), orig_fin(*a))[1]
        lib.xsb_bn_bwd_dxsum = lambda *a: (calls.append("dxsum"), orig_dxs(*a))[1]
        lib.xsb_bn_stats = lambda *a: (calls.append("bn_stats"), orig_bst(*a))[1]
        try:
            yc = F.conv2d(tx, tw, tb, (s, s), p)
            yb = F.batch_norm(yc, tg, tbt, mm, mv, 1, True, 1e-6, 0.99)
            conv_out = yc.eval.astype(np.float64)       # what the GPU conv produced (tf32)
            got = yb.eval
            g = rs.randn(*got.shape).astype(np.float32)
            yb.backward(gradients=g)
        finally:
            lib.xsb_conv2d_fwd_stats, lib.xsb_bn_finalize, lib.xsb_bn_bwd_dxsum = orig_stats, orig_fin, orig_dxs
            lib.xsb_bn_stats = orig_bst
        # (a handful of tiles with a long K -- e.g. one 98-pixel tile with K = 2304 -- run split-K: partial tiles carry no
        # statistics and the BatchNorm takes them from one pass over the small output instead; the shapes that fill the
        # SMs must come through the fused path)
        assert len(calls) == 3 and calls[0] == "stats" and calls[2] == "dxsum" and calls[1] in ("finalize", "bn_stats"), calls
        if n >= 24 or c <= 4:
            assert calls[1] == "finalize", calls
        # statistics: exact w.r.t. the conv output they were gathered from
        mean_ref, var_ref = conv_out.mean(axis=(0, 2, 3)), conv_out.var(axis=(0, 2, 3))
        assert X.rel_err(mm.eval, 0.01 * mean_ref) <= 1e-5 and X.rel_err(mv.eval, 0.99 + 0.01 * var_ref) <= 1e-5
        ybn_ref, xhat, sv, _, _ = X.batch_norm_fwd(conv_out, gm.astype(np.float64), bt.astype(np.float64), np.zeros(oc),
                                                   np.ones(oc), 1, 1e-6, 0.99)
        assert X.rel_err(got, ybn_ref) <= 1e-5
        # whole chain against the float64 oracle
        yr, col = X.conv2d_fwd(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), (s, s), p)
        yb_r, xhat_r, sv_r, _, _ = X.batch_norm_fwd(yr, gm.astype(np.float64), bt.astype(np.float64), np.zeros(oc),
                                                    np.ones(oc), 1, 1e-6, 0.99)
        assert X.rel_err(got, yb_r) <= TOL_TC
        dyc, dg, dbt = X.batch_norm_bwd(g.astype(np.float64), xhat_r, sv_r, gm.astype(np.float64), 1)
        dx, dw, db = X.conv2d_bwd(dyc, x.shape, w.astype(np.float64), col, (s, s), p, need_dx=c >= 32)
        assert X.rel_err(tg.grad.eval, dg) <= TOL_TC and X.rel_err(tbt.grad.eval, dbt) <= TOL_TC
        assert X.rel_err(tw.grad.eval, dw) <= TOL_TC
        if c >= 32:
            assert X.rel_err(tx.grad.eval, dx) <= TOL_TC
        # the conv bias gradient is analytically zero (BN removes the mean): both sides are rounding noise
        assert float(np.max(np.abs(tb.grad.eval))) <= 1e-3 * float(np.max(np.abs(g))) * np.sqrt(g.size / oc)
    finally:
        xs.set_math_mode("fp32")


@pytest.mark.parametrize("n,c,h,oc,k,s,p", [(4, 128, 14, 256, 3, 1, 1), (8, 64, 28, 128, 3, 2, 1), (8, 64, 28, 128, 1, 2, 0),
                                          (4, 64, 15, 128, 3, 2, 1), (2, 64, 17, 128, 7, 2, 3), (2, 64, 8, 64, 2, 3, 0),
                                          (24, 256, 7, 512, 3, 1, 1), (3, 96, 9, 64, 5, 1, 2)])
def test_conv_dgrad_mn_major_filter_tf32(n, c, h, oc, k, s, p):
    """dgrad with the forward filter read AS STORED (MN-major B operand, xsb_tc_config(1, 1); opt-in): stride-1 flipped
    taps, every parity class of strided convolutions, column parts -- against the float64 oracle."""
    import xshinnosuke_b200 as xs
    from oracle import xs_numpy as X
    from xshinnosuke_b200 import nn
    from xshinnosuke_b200._lib import lib
    from xshinnosuke_b200.nn import functional as F
    xs.set_math_mode("tf32")
    lib.xsb_tc_config(1, 1)
    try:
        rs = np.random.RandomState(5 * n + c + h)
        x = rs.randn(n, c, h, h).astype(np.float32)
        w = (rs.randn(oc, c, k, k) * 0.1).astype(np.float32)
        tx, tw = xs.tensor(x, requires_grad=True), nn.Parameter(w)
        y = F.conv2d(tx, tw, None, (s, s), p)
        yr, col = X.conv2d_fwd(x.astype(np.float64), w.astype(np.float64), None, (s, s), p)
        g = rs.randn(*yr.shape).astype(np.float32)
        y.backward(gradients=g)
        dx = X.conv2d_bwd(g.astype(np.float64), x.shape, w.astype(np.float64), col, (s, s), p)[0]
        assert X.rel_err(tx.grad.eval, dx) <= TOL_TC
    finally:
        lib.xsb_tc_config(1, 0)
        xs.set_math_mode("fp32")


@pytest.mark.parametrize("n,c,h,oc,k,s,p", [(4, 128, 14, 256, 3, 1, 1), (24, 256, 7, 512, 3, 1, 1), (100, 128, 14, 256, 3, 1, 1),
                                          (9, 64, 28, 128, 3, 1, 1), (5, 128, 9, 256, 1, 1, 0)])
def test_conv_cta_pair_kernels_tf32(n, c, h, oc, k, s, p):
    """The cta_group::2 forward kernels (two CTAs, M = 256 MMAs issued by the leader, B split across the pair; opt-in):
    fwd, stride-1 dgrad and the fused BN statistics against the float64 oracle, incl. odd tile counts (the last pair's
    second CTA has no tile), column parts and the split tail wave."""
    import xshinnosuke_b200 as xs
    from oracle import xs_numpy as X
    from xshinnosuke_b200 import nn
    from xshinnosuke_b200._lib import lib
    from xshinnosuke_b200.nn import functional as F
    xs.set_math_mode("tf32")
    lib.xsb_tc_config(0, 1)
    try:
        rs = np.random.RandomState(3 * n + c + h)
        x = rs.randn(n, c, h, h).astype(np.float32)
        w = (rs.randn(oc, c, k, k) * 0.1).astype(np.float32)
        b = rs.randn(1, oc).astype(np.float32)
        gm, bt = (rs.rand(oc) + 0.5).astype(np.float32), rs.randn(oc).astype(np.float32)
        tx, tw, tb = xs.tensor(x, requires_grad=True), nn.Parameter(w), nn.Parameter(b)
        mm, mv = xs.tensor(np.zeros(oc, dtype=np.float32)), xs.tensor(np.ones(oc, dtype=np.float32))
        yc = F.conv2d(tx, tw, tb, (s, s), p)
        yb = F.batch_norm(yc, nn.Parameter(gm), nn.Parameter(bt), mm, mv, 1, True, 1e-6, 0.99)
        conv_out = yc.eval.astype(np.float64)
        yr, col = X.conv2d_fwd(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), (s, s), p)
        assert X.rel_err(conv_out, yr) <= TOL_TC
        assert X.rel_err(mm.eval, 0.01 * conv_out.mean(axis=(0, 2, 3))) <= 1e-5
        assert X.rel_err(mv.eval, 0.99 + 0.01 * conv_out.var(axis=(0, 2, 3))) <= 1e-5
        g = rs.randn(*yr.shape).astype(np.float32)
        y2 = F.conv2d(tx, tw, tb, (s, s), p)
        y2.backward(gradients=g)
        dx, dw, db = X.conv2d_bwd(g.astype(np.float64), x.shape, w.astype(np.float64), col, (s, s), p)
        assert X.rel_err(tx.grad.eval, dx) <= TOL_TC and X.rel_err(tw.grad.eval, dw) <= TOL_TC
    finally:
        lib.xsb_tc_config(0, 0)
        xs.set_math_mode("fp32")


@pytest.mark.parametrize("n,c,h,k,s,p", [(4, 64, 20, 3, 2, 1), (3, 8, 17, 3, 2, 1), (2, 16, 12, 2, 2, 0), (2, 64, 9, 3, 1, 1),
                                         (2, 64, 30, 3, 2, 1), (5, 32, 13, 3, 2, 1), (1, 128, 7, 3, 2, 1)])
def test_bn_relu_maxpool_fused_equals_unfused(n, c, h, k, s, p):
    """BatchNormalization -> ReLU(inplace) -> MaxPooling2D: the fused kernel (pool straight from the BN input, the
    rectified tensor never materialised) must give bit-identical pooled values, arg-max, ReLU mask effect and input
    gradient as the three separate kernels; the unfused chain itself is pinned to the oracle by the other tests."""
    import xshinnosuke_b200 as xs
    from oracle import xs_numpy as X
    from xshinnosuke_b200 import nn
    from xshinnosuke_b200._lib import lib
    from xshinnosuke_b200.layers import BatchNormalization, MaxPooling2D, ReLU
    rs = np.random.RandomState(n + c + h)
    x = rs.randn(n, c, h, h).astype(np.float32)
    g = None
    res = []
    keep = (xs.runtime.fuse_bn_relu_pool, xs.runtime.fuse_pool_bn_bwd)
    bwd_calls = []
    for fuse in (True, False):
        xs.runtime.fuse_bn_relu_pool = "all" if fuse else False     # forward: BN apply + ReLU + mask + pool in one kernel
        xs.runtime.fuse_pool_bn_bwd = fuse      # backward: the BatchNorm kernels gather the pool's gradient (3x3/2 only)
        np.random.seed(0)
        bn, relu, pool = BatchNormalization(), ReLU(True), MaxPooling2D(k, s, p)
        calls = []
        orig = lib.xsb_bn_relu_maxpool_fwd
        lib.xsb_bn_relu_maxpool_fwd = lambda *a: (calls.append(1), orig(*a))[1]
        orig_b = lib.xsb_bn_bwd_pool_dxsum
        lib.xsb_bn_bwd_pool_dxsum = lambda *a: (bwd_calls.append(fuse), orig_b(*a))[1]
        try:
            tx = xs.tensor(x, requires_grad=True)
            z = relu(bn(tx))
            y = pool(z)
            yv = y.eval
            am = np.asarray(y.cache["pool_argmax"])
            if g is None:
                g = rs.randn(*yv.shape).astype(np.float32)
            y.backward(gradients=g)
            res.append((yv, am, tx.grad.eval, bn.parameters()[0].grad.eval, len(calls)))
        finally:
            lib.xsb_bn_relu_maxpool_fwd = orig
            lib.xsb_bn_bwd_pool_dxsum = orig_b
            xs.runtime.fuse_bn_relu_pool, xs.runtime.fuse_pool_bn_bwd = keep
    (y1, a1, dx1, dg1, c1), (y0, a0, dx0, dg0, c0) = res
    assert c1 == 1 and c0 == 0
    assert bwd_calls == ([True] if (k, s, p) == (3, 2, 1) else [])
    assert np.array_equal(y1, y0) and np.array_equal(a1, a0)
    # (the BatchNorm backward sums use atomics: summation order, hence the last bits, vary from run to run)
    assert X.rel_err(dx1, dx0) <= 1e-6 and X.rel_err(dg1, dg0) <= 1e-6
    # and against the oracle (fp32 path, 1e-5; arg-max exact)
    xh = x.astype(np.float64)
    zr, xhat, sv, _, _ = X.batch_norm_fwd(xh, np.ones(c), np.zeros(c), np.zeros(c), np.ones(c), 1, 1e-6, 0.99)
    pr, amr = X.max_pool2d_fwd(X.relu_fwd(zr), k, (s, s), p)
    assert X.rel_err(y1, pr) <= 1e-5


@pytest.mark.parametrize("n,c,h,w,oc,k,s,p", [
    (3, 64, 12, 20, 64, 3, 1, 1), (2, 64, 32, 24, 64, 3, 1, 1), (2, 64, 19, 28, 128, 3, 2, 1), (2, 3, 30, 44, 64, 7, 2, 3),
    (2, 128, 9, 14, 256, 3, 1, 1), (1, 64, 16, 40, 64, 3, 1, 1), (2, 64, 21, 13, 128, 1, 2, 0), (2, 4, 17, 33, 64, 5, 2, 2)])
def test_conv_non_square_images_tf32(n, c, h, w, oc, k, s, p):
    """H != W (every kernel takes the two extents separately: im2col, halo fwd / dgrad / wgrad, parity-class dgrad,
    stem): fwd, dX, dW and the fused BN statistics against the float64 oracle."""
    import xshinnosuke_b200 as xs
    from oracle import xs_numpy as X
    from xshinnosuke_b200 import nn
    from xshinnosuke_b200.nn import functional as F
    xs.set_math_mode("tf32")
    try:
        rs = np.random.RandomState(n + c + h + w)
        x = rs.randn(n, c, h, w).astype(np.float32)
        wt = (rs.randn(oc, c, k, k) * 0.1).astype(np.float32)
        b = rs.randn(1, oc).astype(np.float32)
        tx, tw, tb = xs.tensor(x, requires_grad=c >= 32), nn.Parameter(wt), nn.Parameter(b)
        mm, mv = xs.tensor(np.zeros(oc, dtype=np.float32)), xs.tensor(np.ones(oc, dtype=np.float32))
        yc = F.conv2d(tx, tw, tb, (s, s), p)
        yb = F.batch_norm(yc, nn.Parameter(np.ones(oc, dtype=np.float32)), nn.Parameter(np.zeros(oc, dtype=np.float32)), mm, mv,
                          1, True, 1e-6, 0.99)
        conv_out = yc.eval.astype(np.float64)
        yr, col = X.conv2d_fwd(x.astype(np.float64), wt.astype(np.float64), b.astype(np.float64), (s, s), p)
        assert conv_out.shape == yr.shape and X.rel_err(conv_out, yr) <= TOL_TC
        assert X.rel_err(mm.eval, 0.01 * conv_out.mean(axis=(0, 2, 3))) <= 1e-5
        assert X.rel_err(mv.eval, 0.99 + 0.01 * conv_out.var(axis=(0, 2, 3))) <= 1e-5
        g = rs.randn(*yr.shape).astype(np.float32)
        y2 = F.conv2d(tx, tw, tb, (s, s), p)
        y2.backward(gradients=g)
        dx, dw, db = X.conv2d_bwd(g.astype(np.float64), x.shape, wt.astype(np.float64), col, (s, s), p, need_dx=c >= 32)
        assert X.rel_err(tw.grad.eval, dw) <= TOL_TC
        if c >= 32:
            assert X.rel_err(tx.grad.eval, dx) <= TOL_TC
    finally:
        xs.set_math_mode("fp32")


def test_resnet18_tf32_mode(golden):
    """Whole ResNet-18 step in tf32 mode: logits of the first forward within 2e-3 of the reference."""
    import xshinnosuke_b200 as xs
    import parity_suite as S
    xs.set_math_mode("tf32")
    try:
        c = golden.case("resnet18/n4_56")
        net, losses, logits0 = S.resnet_steps(4, 56)
        from oracle import xs_numpy as X
        assert X.rel_err(logits0, c["logits0"]) <= 5e-3     # 20 tf32 convs + BN deep: per-op bound is 2e-3
        assert abs(losses[0] - c["losses"][0]) / c["losses"][0] <= 5e-3
        assert losses[2] < losses[0]
    finally:
        xs.set_math_mode("fp32")


@pytest.mark.parametrize("optimizer", ["sgd", "adam"])
def test_cuda_graph_step_matches_eager(optimizer):
    """a step replayed from a CUDA graph gives the same parameters and losses as the eager loop."""
    import xshinnosuke_b200 as xs
    from oracle import xs_numpy as X
    from xshinnosuke_b200 import nn
    from xshinnosuke_b200.layers import (AvgPooling2D, BatchNormalization, Conv2D, Dense, Flatten, MaxPooling2D, ReLU)

    rs = np.random.RandomState(0)
    xb = [rs.rand(8, 3, 20, 20).astype(np.float32) for _ in range(2)]
    yb = [rs.randint(0, 10, (8,)) for _ in range(2)]

    def build():
        np.random.seed(3)
        net = nn.Sequential(Conv2D(8, 7, stride=2, padding=3), BatchNormalization(), ReLU(True), MaxPooling2D(3, 2, 1),
                            Conv2D(64, 3, padding=1), BatchNormalization(), ReLU(True), AvgPooling2D(2), Flatten(), Dense(10))
        opt = xs.optim.SGD(net.parameters(), lr=0.05) if optimizer == "sgd" else xs.optim.Adam(net.parameters(), lr=0.01)
        return net, opt

    crit = nn.CrossEntropyLoss()
    # eager: 2 warm-up steps on batch 0, then batches 0,1,0
    net_e, opt_e = build()
    losses_e = []
    for i in (0, 0, 0, 1, 0):
        opt_e.zero_grad()
        loss = crit(net_e(xs.tensor(xb[i])), xs.tensor(yb[i]))
        losses_e.append(loss.item())
        loss.backward()
        opt_e.step()
    # graphed: same schedule, the last three steps are replays fed through the static input tensors
    net_g, opt_g = build()
    x_in, y_in = xs.tensor(xb[0]), xs.tensor(yb[0])

    def step():
        opt_g.zero_grad()
        loss = crit(net_g(x_in), y_in)
        loss.backward()
        opt_g.step()
        return loss

    g = xs.GraphedStep(step, warmup=2)
    losses_g = []
    for i in (0, 1, 0):
        x_in.eval = xb[i]
        y_in.eval = yb[i]
        losses_g.append(g().item())
    assert g.launches_per_replay > 20
    np.testing.assert_allclose(losses_g, losses_e[2:], rtol=1e-5)
    for pe, pg in zip(net_e.parameters(), net_g.parameters()):
        if optimizer == "adam" and pe.eval.ndim == 2 and pe.eval.shape[0] == 1 and pe.eval.shape[1] in (8, 64):
            # bias of a conv that feeds a BatchNorm: its gradient is analytically ZERO (BN removes the mean), what is
            # computed is fp32 summation noise (~1e-8), and Adam's g/sqrt(g^2) turns noise into O(lr) steps -- in the
            # reference too, only there the noise is ~1e-17 << eps. The parameter has no effect on any output.
            continue
        assert float(np.max(np.abs(pe.eval - pg.eval))) <= 1e-5 * max(float(np.max(np.abs(pe.eval))), 1e-2)


def test_graphed_step_feeds_prefetched_host_batches():
    """SURVEY 8f-2 on top of 8f-1: g(x_host, y_host) uploads through the double-buffered staging slots and
    g.prefetch(next) overlaps the next upload with the replay; losses equal the eager loop on the same batch order."""
    import torch
    import xshinnosuke_b200 as xs
    from xshinnosuke_b200 import nn
    from xshinnosuke_b200.layers import (AvgPooling2D, BatchNormalization, Conv2D, Dense, Flatten, MaxPooling2D, ReLU)

    rs = np.random.RandomState(5)
    xb = [rs.rand(8, 3, 20, 20).astype(np.float32) for _ in range(3)]
    yb = [rs.randint(0, 10, (8,)).astype(np.int64) for _ in range(3)]
    order = (0, 1, 2, 1, 0, 2)

    def build():
        np.random.seed(11)
        net = nn.Sequential(Conv2D(8, 7, stride=2, padding=3), BatchNormalization(), ReLU(True), MaxPooling2D(3, 2, 1),
                            Conv2D(64, 3, padding=1), BatchNormalization(), ReLU(True), AvgPooling2D(2), Flatten(), Dense(10))
        return net, xs.optim.SGD(net.parameters(), lr=0.05)

    crit = nn.CrossEntropyLoss()
    net_e, opt_e = build()
    losses_e = []
    for i in (0, 0) + order:
        opt_e.zero_grad()
        loss = crit(net_e(xs.tensor(xb[i])), xs.tensor(yb[i]))
        losses_e.append(loss.item())
        loss.backward()
        opt_e.step()
    net_g, opt_g = build()
    x_in, y_in = xs.tensor(xb[0]), xs.tensor(yb[0])

    def step():
        opt_g.zero_grad()
        loss = crit(net_g(x_in), y_in)
        loss.backward()
        opt_g.step()
        return loss

    g = xs.GraphedStep(step, warmup=2, inputs=(x_in, y_in))
    # pinned copies upload asynchronously; plain arrays work too (synchronous path): mix both
    pinned = [(torch.from_numpy(x).pin_memory().numpy(), torch.from_numpy(y).pin_memory().numpy()) for x, y in zip(xb, yb)]
    losses_g = []
    g.prefetch(*pinned[order[0]])
    for k, i in enumerate(order):
        loss = g(*pinned[i]) if k % 2 == 0 else g(xb[i], yb[i])
        if k + 1 < len(order) and (k + 1) % 2 == 0:
            g.prefetch(*pinned[order[k + 1]])
        losses_g.append(loss.item())
    np.testing.assert_allclose(losses_g, losses_e[2:], rtol=1e-5)


def test_static_fit_replayed_as_cuda_graph(golden):
    """Model.fit(..., cuda_graph=True) captures the static graph's train step after two eager batches
    and replays it; the loss trajectory equals the reference's (golden `fit/*`, same tolerance as the eager path),
    including the MLP recipe whose last batch of every epoch is partial (runs eagerly between replays)."""
    import parity_suite as S
    np.random.seed(0)
    xd = np.random.rand(96, 1, 28, 28)
    yd = np.random.randint(0, 10, (96,))
    np.random.seed(1)
    model = S.readme_cnn(lr=0.01)
    prof = model.fit(xd, yd, batch_size=32, epochs=2, verbose=0, shuffle=False, cuda_graph=True)
    np.testing.assert_allclose(prof.data["train_loss"], golden.case("fit/cnn")["losses"], rtol=2e-4)
    np.random.seed(0)
    xd = np.random.rand(80, 784)
    yd = np.random.randint(0, 10, (80,))
    np.random.seed(1)
    mlp = S.readme_mlp(lr=0.05)
    prof = mlp.fit(xd, yd, batch_size=32, epochs=2, verbose=0, shuffle=False, cuda_graph=True)
    np.testing.assert_allclose(prof.data["train_loss"], golden.case("fit/mlp")["losses"], rtol=2e-4)


def test_dataloader_prefetch_matches_plain():
    """the pinned / side-stream prefetching loader yields exactly the reference's batches
    (same order incl. the per-epoch reseeding), and rank shards are contiguous slices."""
    from xshinnosuke_b200.utils.data import DataLoader, DataSet
    rs = np.random.RandomState(0)
    X_, Y_ = rs.rand(70, 3, 9, 9).astype(np.float32), rs.randint(0, 10, (70,))
    for shuffle in (False, True):
        a = DataLoader(DataSet(X_, Y_), 16, shuffle, prefetch=True)
        b = DataLoader(DataSet(X_, Y_), 16, shuffle, prefetch=False)
        for epoch in range(2):
            got = [(x.eval, y.eval) for x, y in a]
            want = [(x.eval, y.eval) for x, y in b]
            assert len(got) == len(want) == 5
            for (gx, gy), (wx, wy) in zip(got, want):
                assert np.array_equal(gx, wx) and np.array_equal(gy, wy)
    full = DataLoader(DataSet(X_, Y_), 16, False, prefetch=False)
    r1 = DataLoader(DataSet(X_, Y_), 16, False, rank=1, world=2)
    for (fx, fy), (sx, sy) in zip(full, r1):
        per = fx.shape[0] // 2
        assert np.array_equal(sx.eval, fx.eval[per:2 * per]) and np.array_equal(sy.eval, fy.eval[per:2 * per])


# ------------------------------------------------------------------ fp32x3: fp32-accurate products on the tensor cores
@pytest.mark.parametrize("n,c,h,oc,k,s,p,n_tc", [
    (4, 64, 14, 64, 3, 1, 1, 9), (8, 64, 28, 128, 3, 2, 1, 9), (8, 64, 28, 128, 1, 2, 0, 9), (4, 128, 14, 256, 3, 1, 1, 9),
    (2, 256, 7, 512, 3, 1, 1, 9), (16, 64, 56, 64, 3, 1, 1, 9), (4, 3, 56, 64, 7, 2, 3, 6), (3, 32, 9, 64, 3, 1, 1, 6),
    (2, 3, 20, 8, 3, 1, 1, 0), (100, 128, 14, 256, 3, 1, 1, 9)])
def test_conv_fp32x3_tensor_cores(n, c, h, oc, k, s, p, n_tc):
    """fp32x3 math mode (nn/x3.py): fwd / dgrad / wgrad as three accumulating tcgen05 kind::tf32 launches over the hi/lo
    operand split, held to the FP32 bar (1e-5 vs the float64 oracle, north_star) -- `n_tc` tensor-core entry calls must
    have served the three ops (3 each inside the envelope; the stem's dgrad and off-grid channel counts run one FFMA call).
    Reductions longer than 2304 products (ResNet-18 layer4: K = 4608) are held to 2e-5: the operands are exact, what is
    left is tcgen05's own fp32 accumulation of one long chain in TMEM (measured 1.1e-5 at K = 4608)."""
    import xshinnosuke_b200 as xs
    from oracle import xs_numpy as X
    from xshinnosuke_b200 import nn
    from xshinnosuke_b200.nn import functional as F
    xs.set_math_mode("fp32x3")
    try:
        rs = np.random.RandomState(n * 1000 + c + h)
        x = rs.randn(n, c, h, h).astype(np.float32)
        w = (rs.randn(oc, c, k, k) * 0.1).astype(np.float32)
        b = rs.randn(1, oc).astype(np.float32)
        s0 = _tc_stats()
        tx, tw, tb = xs.tensor(x, requires_grad=True), nn.Parameter(w), nn.Parameter(b)
        y = F.conv2d(tx, tw, tb, (s, s), p)
        yr, col = X.conv2d_fwd(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), (s, s), p)
        tol = TOL if max(c, oc) * k * k <= 2304 else 2e-5
        assert X.rel_err(y.eval, yr) <= tol
        g = rs.randn(*yr.shape).astype(np.float32)
        y.backward(gradients=g)
        dx, dw, db = X.conv2d_bwd(g.astype(np.float64), x.shape, w.astype(np.float64), col, (s, s), p)
        assert X.rel_err(tx.grad.eval, dx) <= tol
        assert X.rel_err(tw.grad.eval, dw) <= tol
        s1 = _tc_stats()
        assert s1[0] - s0[0] == n_tc and s1[1] == s0[1]       # no silent FFMA fallback inside a split product
    finally:
        xs.set_math_mode("fp32")


def test_dense_and_resnet_fp32x3():
    """Dense products and a whole ResNet-18 step in fp32x3 mode: logits / loss at the FP32 bar of the fp32 (FFMA) mode."""
    import xshinnosuke_b200 as xs
    from oracle import xs_numpy as X
    from xshinnosuke_b200 import nn
    from xshinnosuke_b200.nn import functional as F
    import parity_suite as S
    xs.set_math_mode("fp32x3")
    try:
        rs = np.random.RandomState(3)
        x, w, b = rs.randn(128, 784).astype(np.float32), (rs.randn(784, 500) * 0.1).astype(np.float32), rs.randn(1, 500).astype(np.float32)
        tx, tw, tb = xs.tensor(x, requires_grad=True), nn.Parameter(w), nn.Parameter(b)
        s0 = _tc_stats()
        y = F.addmm(tb, tx, tw)
        assert X.rel_err(y.eval, x.astype(np.float64) @ w.astype(np.float64) + b) <= TOL
        g = rs.randn(128, 500).astype(np.float32)
        y.backward(gradients=g)
        assert X.rel_err(tx.grad.eval, g.astype(np.float64) @ w.astype(np.float64).T) <= TOL
        assert X.rel_err(tw.grad.eval, x.astype(np.float64).T @ g.astype(np.float64)) <= TOL
        assert _tc_stats()[0] - s0[0] == 9
        net, losses, logits0 = S.resnet_steps(4, 56)
        from conftest import Golden
        c = Golden().case("resnet18/n4_56")
        assert X.rel_err(logits0, c["logits0"]) <= 5e-5
        assert abs(losses[0] - c["losses"][0]) / c["losses"][0] <= 2e-5
    finally:
        xs.set_math_mode("fp32")


# ------------------------------------------------------------------ parity AT the benchmarked shapes (BASELINE configs[4])
@pytest.mark.parametrize("name,c,h,oc,k,s,p", [
    ("stem", 3, 224, 64, 7, 2, 3), ("layer1", 64, 56, 64, 3, 1, 1), ("layer2.0", 64, 56, 128, 3, 2, 1),
    ("layer2", 128, 28, 128, 3, 1, 1), ("layer2.sc", 64, 56, 128, 1, 2, 0), ("layer3.0", 128, 28, 256, 3, 2, 1),
    ("layer3", 256, 14, 256, 3, 1, 1), ("layer4.0", 256, 14, 512, 3, 2, 1), ("layer4", 512, 7, 512, 3, 1, 1)])
def test_conv_tf32_at_benchmark_shapes(name, c, h, oc, k, s, p):
    """Every ResNet-18 convolution shape of the benchmarked configuration (N = 128 per GPU, 3 x 224 x 224) in tf32 mode:
    forward and dgrad are batch-separable, so images 0, 77 and 127 of the full-batch GPU result are held to 2e-3 of the
    float64 oracle run on those three images; the weight gradient sums over the batch -- it is held to 2e-3 of the fp32
    (FFMA) mode of the same library on the full batch, which is itself pinned to the oracle at 1e-5 on smaller batches."""
    import xshinnosuke_b200 as xs
    from oracle import xs_numpy as X
    from xshinnosuke_b200 import nn
    from xshinnosuke_b200.nn import functional as F
    n, pick = 128, [0, 77, 127]
    rs = np.random.RandomState(c + h + oc)
    x = rs.randn(n, c, h, h).astype(np.float32)
    w = (rs.randn(oc, c, k, k) * 0.1).astype(np.float32)
    b = rs.randn(1, oc).astype(np.float32)
    oh = (h - k + 2 * p) // s + 1
    g = rs.randn(n, oc, oh, oh).astype(np.float32)

    def run(mode):
        xs.set_math_mode(mode)
        tx, tw, tb = xs.tensor(x, requires_grad=c >= 32), nn.Parameter(w), nn.Parameter(b)
        y = F.conv2d(tx, tw, tb, (s, s), p)
        yv = y.eval[pick]
        y.backward(gradients=g)
        return yv, (tx.grad.eval[pick] if c >= 32 else None), tw.grad.eval
    try:
        s0 = _tc_stats()
        y_tc, dx_tc, dw_tc = run("tf32")
        s1 = _tc_stats()
        assert s1[0] - s0[0] == (3 if c >= 32 else 2)
        yr, col = X.conv2d_fwd(x[pick].astype(np.float64), w.astype(np.float64), b.astype(np.float64), (s, s), p)
        assert X.rel_err(y_tc, yr) <= TOL_TC
        if c >= 32:
            dxr, _, _ = X.conv2d_bwd(g[pick].astype(np.float64), x[pick].shape, w.astype(np.float64), col, (s, s), p)
            assert X.rel_err(dx_tc, dxr) <= TOL_TC
        del col
        _, _, dw_fp32 = run("fp32")
        assert X.rel_err(dw_tc, dw_fp32) <= TOL_TC
    finally:
        xs.set_math_mode("fp32")


@pytest.mark.parametrize("n,c,h,two", [(4, 64, 14, False), (4, 64, 14, True), (3, 8, 9, False), (2, 512, 7, True),
                                       (128, 128, 28, False), (128, 256, 14, True)])
def test_residual_tail_fused_equals_unfused(n, c, h, two):
    """BatchNormalization -> add -> ReLU(inplace) (the tail of a residual block; `two`: the shortcut ends in a BatchNorm as
    well): xsb_bn_add_relu against xsb_bn_apply (x2) + xsb_add_relu -- forward output bit-identical (same arithmetic,
    rounded at the same points), gradients to 1e-6 (atomics), incl. the last two shapes of the benchmarked configuration."""
    import xshinnosuke_b200 as xs
    from oracle import xs_numpy as X
    from xshinnosuke_b200._lib import lib
    from xshinnosuke_b200.layers import BatchNormalization, ReLU
    rs = np.random.RandomState(n + c + h)
    x1 = rs.randn(n, c, h, h).astype(np.float32)
    x2 = rs.randn(n, c, h, h).astype(np.float32)
    g = rs.randn(n, c, h, h).astype(np.float32)
    keep = xs.runtime.fuse_bn_add_relu
    res, calls = [], []
    orig = lib.xsb_bn_add_relu
    lib.xsb_bn_add_relu = lambda *a: (calls.append(1), orig(*a))[1]
    try:
        for fuse in (True, False):
            xs.runtime.fuse_bn_add_relu = fuse
            np.random.seed(0)
            bn_a, bn_b, relu = BatchNormalization(), BatchNormalization(), ReLU(True)
            ta, tb = xs.tensor(x1, requires_grad=True), xs.tensor(x2, requires_grad=True)
            a = bn_a(ta)
            b = bn_b(tb) if two else tb
            z = relu(xs.nn.functional.add(a, b))
            zv = z.eval
            z.backward(gradients=g)
            res.append((zv, ta.grad.eval, tb.grad.eval, bn_a.parameters()[0].grad.eval))
    finally:
        lib.xsb_bn_add_relu = orig
        xs.runtime.fuse_bn_add_relu = keep
    assert len(calls) == 1
    (z1, da1, db1, dg1), (z0, da0, db0, dg0) = res
    assert np.array_equal(z1, z0)
    assert X.rel_err(da1, da0) <= 1e-6 and X.rel_err(db1, db0) <= 1e-6 and X.rel_err(dg1, dg0) <= 1e-5


@pytest.mark.parametrize("n,c,h,mode", [(4, 64, 14, "fp32"), (3, 8, 9, "fp32"), (128, 64, 56, "tf32"), (128, 256, 14, "tf32")])
def test_identity_shortcut_side_output_equals_separate_kernel(n, c, h, mode):
    """Residual block with an identity shortcut, backward: the masked dZ the shortcut needs is written by the BatchNorm
    backward of the other branch (xsb_bn_bwd_dxsum_side; cooperative kernel for the small tensors, two-launch persistent
    kernels for the layer1-sized one) instead of by xsb_relu_bwd_mask -- the same values written first, the conv dgrad
    added second: every gradient equal up to the atomics of the BatchNorm sums."""
    import xshinnosuke_b200 as xs
    from oracle import xs_numpy as X
    from xshinnosuke_b200._lib import lib
    from xshinnosuke_b200.layers import BatchNormalization, Conv2D, ReLU
    rs = np.random.RandomState(n + c + h)
    x = rs.randn(n, c, h, h).astype(np.float32)
    g = rs.randn(n, c, h, h).astype(np.float32)
    keep = xs.runtime.fuse_add_bwd_side
    res, calls = [], []
    orig = lib.xsb_bn_bwd_dxsum_side
    lib.xsb_bn_bwd_dxsum_side = lambda *a: (calls.append(1), orig(*a))[1]
    xs.set_math_mode(mode)
    try:
        for fuse in (True, False):
            xs.runtime.fuse_add_bwd_side = fuse
            np.random.seed(0)
            pre, conv, bn, relu = ReLU(True), Conv2D(c, 3, padding=1), BatchNormalization(), ReLU(True)
            tx = xs.tensor(x, requires_grad=True)
            z0 = pre(BatchNormalization()(tx))
            z = relu(xs.nn.functional.add(bn(conv(z0)), z0))
            zv = z.eval
            z.backward(gradients=g)
            res.append((zv, tx.grad.eval, conv.parameters()[0].grad.eval, bn.parameters()[0].grad.eval))
            del tx, z0, z
    finally:
        lib.xsb_bn_bwd_dxsum_side = orig
        xs.runtime.fuse_add_bwd_side = keep
        xs.set_math_mode("fp32")
    assert len(calls) == 1
    (z1, dx1, dw1, dg1), (z0_, dx0, dw0, dg0) = res
    assert np.array_equal(z1, z0_)
    # (the 26 M-element tensors: two runs of the SAME path differ by 2.5e-5 in dX -- atomics in the BatchNorm sums, measured)
    tol = 1e-6 if mode == "fp32" else 1e-4
    assert X.rel_err(dx1, dx0) <= tol and X.rel_err(dw1, dw0) <= tol and X.rel_err(dg1, dg0) <= 10 * tol


def test_stem_fusions_at_benchmark_shape():
    """BatchNormalization -> ReLU(inplace) -> MaxPooling2D(3, 2, 1) on the stem tensor of the benchmarked configuration
    (128 x 64 x 112 x 112): the fused forward kernel and the pool-gathering BatchNorm backward against the separate
    kernels (which the small-shape tests pin to the oracle) -- pooled values and arg-max bit-identical, dX / dgamma /
    dbeta to 1e-6 (atomics: summation order)."""
    import xshinnosuke_b200 as xs
    from oracle import xs_numpy as X
    from xshinnosuke_b200._lib import lib
    from xshinnosuke_b200.layers import BatchNormalization, MaxPooling2D, ReLU
    rs = np.random.RandomState(11)
    x = rs.randn(128, 64, 112, 112).astype(np.float32)
    g = rs.randn(128, 64, 56, 56).astype(np.float32)
    keep = (xs.runtime.fuse_bn_relu_pool, xs.runtime.fuse_pool_bn_bwd)
    res, calls = [], []
    orig_f, orig_b = lib.xsb_bn_relu_maxpool_fwd, lib.xsb_bn_bwd_pool_dxsum
    lib.xsb_bn_relu_maxpool_fwd = lambda *a: (calls.append("fwd"), orig_f(*a))[1]
    lib.xsb_bn_bwd_pool_dxsum = lambda *a: (calls.append("bwd"), orig_b(*a))[1]
    try:
        for fuse in (True, False):
            xs.runtime.fuse_bn_relu_pool = xs.runtime.fuse_pool_bn_bwd = fuse
            np.random.seed(0)
            bn, relu, pool = BatchNormalization(), ReLU(True), MaxPooling2D(3, 2, 1)
            tx = xs.tensor(x, requires_grad=True)
            y = pool(relu(bn(tx)))
            yv, am = y.eval, np.asarray(y.cache["pool_argmax"])
            y.backward(gradients=g)
            res.append((yv, am, tx.grad.eval, bn.parameters()[0].grad.eval, bn.parameters()[1].grad.eval))
            del tx, y, bn, relu, pool
    finally:
        lib.xsb_bn_relu_maxpool_fwd, lib.xsb_bn_bwd_pool_dxsum = orig_f, orig_b
        xs.runtime.fuse_bn_relu_pool, xs.runtime.fuse_pool_bn_bwd = keep
    assert calls == ["fwd", "bwd"]
    (y1, a1, dx1, dg1, db1), (y0, a0, dx0, dg0, db0) = res
    assert np.array_equal(y1, y0) and np.array_equal(a1, a0)
    assert X.rel_err(dx1, dx0) <= 1e-6 and X.rel_err(dg1, dg0) <= 1e-5 and X.rel_err(db1, db0) <= 1e-5


# ------------------------------------------------------------------ data parallel over NCCL (needs >= 2 GPUs)
@pytest.mark.parametrize("optimizer,graph", [("sgd", False), ("adam", False), ("sgd", True), ("adam", True)])
def test_data_parallel_nccl_matches_per_shard_oracle(optimizer, graph):
    """tests/dp_worker.py under torchrun on 2 GPUs: NCCL bucket all-reduces overlapped with backward (eager and captured
    into the step's CUDA graph), parameters after 4 steps within 2e-5 of the per-shard oracle average."""
    import os
    import socket
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(here, "dp_worker.py"), optimizer] + (["graph"] if graph else [])
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:]
    assert res.stdout.count("DP_RESULT") == 2


def test_model_evaluate_predict_save_load(tmp_path):
    import parity_suite as S
    S.check_model_api(TOL, tmp_path)


@pytest.mark.parametrize("mode,tol", [("fp32", TOL), ("tf32", TOL_TC)])
def test_fused_bias_relu_epilogue(mode, tol):
    """north_star "fused bias+activation epilogue": Dense / Conv2D with activation='relu' = one fused C-ABI call."""
    import contextlib
    import parity_suite as S
    import xshinnosuke_b200 as xs
    from xshinnosuke_b200._lib import lib

    @contextlib.contextmanager
    def spy():
        names = ("xsb_gemm_bias_relu", "xsb_conv2d_fwd_relu", "xsb_relu_fwd", "xsb_gemm", "xsb_conv2d_fwd")
        seen, originals = [], {n: getattr(lib, n) for n in names}
        for n, fn in originals.items():
            setattr(lib, n, (lambda *a, _fn=fn, _n=n: (seen.append(_n), _fn(*a))[1]))
        try:
            yield seen
        finally:
            for n, fn in originals.items():
                setattr(lib, n, fn)
    xs.set_math_mode(mode)
    try:
        S.check_fused_bias_relu(tol, spy)
    finally:
        xs.set_math_mode("fp32")


def test_cross_entropy_label_range():
    import parity_suite as S
    S.check_label_range()


def test_extra_ops(golden):
    """Sigmoid / Tanh / Dropout / ZeroPadding2D / broadcasting add / LayerNorm / GroupNorm."""
    import parity_suite as S
    S.check_extra_ops(golden, TOL)
