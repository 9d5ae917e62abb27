"""CPU: the HOST logic of xshinnosuke_b200 (graph wiring, autograd walk, lazy-zero gradients, flat
optimizer buffers, in-place ReLU protocol, static-graph fit) driven through the public API against
the reference's golden vectors, with the C ABI served by tests/fake_backend.py (NumPy) instead of
the CUDA library. The same checks run on the real kernels in tests/test_parity_gpu.py."""
import pytest

import fake_backend
import parity_suite as S

TOL = 2e-6   # fake kernels compute in float64 and store fp32


@pytest.fixture(autouse=True)
def _fake(monkeypatch):
    import xshinnosuke_b200.core.state as G
    G.INPUTS = G.OUTPUTS = G.GRAPH = None
    yield fake_backend.install(monkeypatch)
    G.INPUTS = G.OUTPUTS = G.GRAPH = None


def test_conv(golden):
    S.check_conv(golden, TOL, [n for n in golden.names("conv") if n != "k3s1p0_f64"] + ["k3s1p0_f64"])


def test_pools(golden):
    S.check_maxpool(golden, TOL)
    S.check_avgpool(golden, TOL)


def test_relu_add(golden):
    S.check_relu_add(golden, TOL)
    S.check_inplace_relu_layer(TOL)


def test_postponed_shortcut_dgrad():
    S.check_postponed_shortcut_dgrad(TOL)


def test_postponed_add_is_dropped_by_zero_grad_and_flushed_by_clone():
    """DevArray.defer_add bookkeeping: a parked kernel belongs to the gradient it was parked on -- zero_grad discards it,
    clone / to_numpy run it first, and a second parked kernel is refused (it runs immediately as the "+=" writer)."""
    import numpy as np
    from xshinnosuke_b200.core.device import DevArray
    ran = []
    a = DevArray.empty((2, 3), pending_zero=True)
    a.cl = False
    assert a.defer_add(lambda acc: ran.append(("first", acc)))
    assert not a.defer_add(lambda acc: ran.append(("second", acc)))      # one parked kernel at most
    assert a.take_acc() == 0 and not ran                                  # a writer in between: "=", nothing flushed
    a.buf.zero_()
    b = a.clone()                                                         # a reader: the parked kernel runs as "+="
    assert ran == [("first", 1)] and not b.pending_zero
    c = DevArray.empty((4,), pending_zero=True)
    c.cl = False
    assert c.defer_add(lambda acc: ran.append(("third", acc)))
    assert np.array_equal(c.to_numpy(), np.zeros(4, dtype=np.float32)) or ran[-1] == ("third", 0)
    assert ran[-1] == ("third", 0)                                        # alone: it is the first writer after all
    import xshinnosuke_b200 as xs
    t = xs.tensor(np.ones((2, 2), dtype=np.float32), requires_grad=True)
    t.zero_grad()
    assert t.grad.arr.defer_add(lambda acc: ran.append(("stale", acc)))
    t.zero_grad()                                                         # discards the parked kernel
    assert np.array_equal(t.grad.eval, np.zeros((2, 2), dtype=np.float32)) and ran[-1] != ("stale", 0)


def test_bn(golden):
    S.check_bn(golden, 1e-5)
    S.check_bn_eval(1e-6)


def test_dense_losses(golden):
    S.check_dense_losses(golden, TOL)


def test_optimizers(golden):
    S.check_optimizers(golden, TOL)
    S.check_weight_decay(golden, TOL)


def test_gradient_check_scripts(golden):
    S.check_gradient_check_scripts(golden, 1e-5)


def test_trajectories(golden):
    S.check_test_fc(golden, 1e-6)
    S.check_test_conv(golden, 1e-6, steps=6)


def test_resnet18_small(golden):
    S.check_resnet18(golden, "n2_33", 2, 33, 1e-4, 1e-5)


def test_fit_static_graph(golden):
    S.check_fit_configs(golden, 1e-6)


def test_fp32x3_mode_host_logic(golden):
    """fp32x3 math mode (nn/x3.py): every GEMM-shaped op inside the tensor-core envelope is issued as the three accumulating
    products of the hi/lo operand split (split, lo.hi, hi.lo, hi.hi+bias), the others as one fp32 call; results stay the
    reference's (the fake kernels are exact, so the split must reassemble x.w to fp32 rounding)."""
    import numpy as np
    import xshinnosuke_b200 as xs
    from oracle import xs_numpy as X
    from xshinnosuke_b200 import nn
    from xshinnosuke_b200.nn import functional as F
    xs.set_math_mode("fp32x3")
    try:
        rs = np.random.RandomState(5)
        x = rs.randn(2, 64, 6, 6).astype(np.float32)
        w = (rs.randn(64, 64, 3, 3) * 0.1).astype(np.float32)
        b = rs.randn(1, 64).astype(np.float32)
        tx, tw, tb = xs.tensor(x, requires_grad=True), nn.Parameter(w), nn.Parameter(b)
        fake = fake_backend.install()
        del fake.calls[:]
        y = F.conv2d(tx, tw, tb, (1, 1), 1)
        yr, col = X.conv2d_fwd(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), (1, 1), 1)
        assert X.rel_err(y.eval, yr) <= 1e-6
        assert fake.calls.count("xsb_conv2d_fwd_acc") == 3 and fake.calls.count("xsb_split_tf32") == 2
        g = rs.randn(*yr.shape).astype(np.float32)
        y.backward(gradients=g)
        dx, dw, db = X.conv2d_bwd(g.astype(np.float64), x.shape, w.astype(np.float64), col, (1, 1), 1)
        assert X.rel_err(tx.grad.eval, dx) <= 1e-6 and X.rel_err(tw.grad.eval, dw) <= 1e-6
        assert fake.calls.count("xsb_conv2d_wgrad") == 3 and fake.calls.count("xsb_conv2d_dgrad") == 3
        # outside the envelope (20 channels): ONE fp32 call, no split
        del fake.calls[:]
        x2 = rs.randn(2, 20, 5, 5).astype(np.float32)
        w2 = (rs.randn(12, 20, 3, 3) * 0.1).astype(np.float32)
        y2 = F.conv2d(xs.tensor(x2), nn.Parameter(w2), None, (1, 1), 1)
        y2.eval
        assert fake.calls.count("xsb_conv2d_fwd") == 1 and "xsb_split_tf32" not in fake.calls
    finally:
        xs.set_math_mode("fp32")


def test_model_evaluate_predict_save_load(tmp_path):
    S.check_model_api(1e-6, tmp_path)


def test_fused_bias_relu_epilogue():
    import contextlib
    fake = fake_backend.install()

    @contextlib.contextmanager
    def spy():
        start = len(fake.calls)
        view = []
        yield view
        view.extend(fake.calls[start:])
    S.check_fused_bias_relu(TOL, spy)


def test_cross_entropy_label_range():
    S.check_label_range()


def test_direct_writes_and_replaced_gradients():
    """a direct write (`t.eval = a`, zero_, fill_) voids whatever was parked on the array (deferred producer,
    postponed add); a gradient tensor the user swapped in after the optimizer flattened its buffers is still applied."""
    import numpy as np
    import xshinnosuke_b200 as xs
    from xshinnosuke_b200 import nn
    from xshinnosuke_b200.core.device import DevArray
    from xshinnosuke_b200.nn import functional as F
    ran = []
    a = DevArray.empty((2, 3), pending_zero=True)
    a.cl = False
    assert a.defer_add(lambda acc: ran.append(acc))
    a.pending_op = lambda m=None: ran.append("producer")
    a.assign_numpy(np.full((2, 3), 7.0, dtype=np.float32))
    assert np.array_equal(a.to_numpy(), np.full((2, 3), 7.0, dtype=np.float32)) and not ran
    t = xs.tensor(np.ones((2, 2), dtype=np.float32), requires_grad=True)
    t.zero_grad()
    assert t.grad.arr.defer_add(lambda acc: ran.append(acc))
    t.grad.zero_()
    assert np.array_equal(t.grad.eval, np.zeros((2, 2), dtype=np.float32)) and not ran
    # deferred conv output overwritten by hand: the producer must not run afterwards
    x = xs.tensor(np.random.RandomState(0).randn(1, 2, 4, 4).astype(np.float32))
    w = nn.Parameter(np.random.RandomState(1).randn(3, 2, 3, 3).astype(np.float32))
    y = F.conv2d(x, w, None, (1, 1), 1)
    y.eval = np.zeros((1, 3, 4, 4), dtype=np.float32)
    assert np.array_equal(y.eval, np.zeros((1, 3, 4, 4), dtype=np.float32))
    # replaced gradient
    p = nn.Parameter(np.ones((4,), dtype=np.float32))
    opt = xs.optim.SGD([p], lr=0.5)
    opt.zero_grad()
    p.grad = xs.tensor(np.array([1.0, 2.0, 3.0, 4.0], dtype=np.float32))
    opt.step()
    assert np.allclose(p.eval, 1.0 - 0.5 * np.array([1.0, 2.0, 3.0, 4.0]))


def test_extra_ops(golden):
    S.check_extra_ops(golden, TOL)


def test_stem_pool_fusions_host_logic():
    """BatchNormalization -> ReLU(inplace) -> MaxPooling2D(3, 2, 1): forward through xsb_bn_relu_maxpool_fwd (the rectified
    tensor stays un-materialised), backward with the pool's gradient PARKED on the ReLU tensor's gradient and consumed by
    xsb_bn_bwd_pool_dxsum; a reader that gets to the parked gradient first receives the scattered copy (plain kernels)."""
    import numpy as np
    import xshinnosuke_b200 as xs
    from xshinnosuke_b200._lib import lib
    from xshinnosuke_b200.layers import BatchNormalization, MaxPooling2D, ReLU
    rs = np.random.RandomState(5)
    x = rs.randn(3, 8, 11, 11).astype(np.float32)
    g = None
    out = {}
    keep = (xs.runtime.fuse_bn_relu_pool, xs.runtime.fuse_pool_bn_bwd)
    try:
        for mode in ("fused", "peek", "plain"):
            xs.runtime.fuse_bn_relu_pool = mode != "plain"
            xs.runtime.fuse_pool_bn_bwd = mode != "plain"
            np.random.seed(0)
            bn, relu, pool = BatchNormalization(), ReLU(True), MaxPooling2D(3, 2, 1)
            tx = xs.tensor(x, requires_grad=True)
            n0 = len(lib.calls)
            z = relu(bn(tx))
            y = pool(z)
            yv = y.eval                                   # (backward releases the graph's intermediate tensors)
            if g is None:
                g = rs.randn(*yv.shape).astype(np.float32)
            if mode == "peek":
                # run the pool's closure by hand, then READ the gradient it parked before the BatchNorm closure runs
                y.grad = xs.tensor(g)
                y.grad_fn(y)
                parked = z.grad.arr._pz[2]
                assert parked is not None and hasattr(parked, "pool")
                dz = z.grad.eval                          # a reader: the scattered copy is made now
                assert z.grad.arr._pz[2] is None and np.abs(dz).sum() > 0
                z.grad_fn(z)
            else:
                y.backward(gradients=g)
            calls = lib.calls[n0:]
            out[mode] = (yv, tx.grad.eval, bn.parameters()[0].grad.eval, bn.parameters()[1].grad.eval, calls)
    finally:
        xs.runtime.fuse_bn_relu_pool, xs.runtime.fuse_pool_bn_bwd = keep
    f, p, q = out["fused"], out["peek"], out["plain"]
    assert "xsb_bn_relu_maxpool_fwd" in f[4] and "xsb_bn_bwd_pool_dxsum" in f[4] and "xsb_maxpool2d_bwd" not in f[4]
    assert "xsb_bn_bwd_pool_dxsum" not in p[4] and "xsb_maxpool2d_bwd" in p[4]
    assert "xsb_bn_relu_maxpool_fwd" not in q[4] and "xsb_bn_bwd_pool_dxsum" not in q[4]
    for a in (f, p):
        for u, v in zip(a[:4], q[:4]):
            assert np.abs(u - v).max() <= TOL * max(1.0, np.abs(v).max())


def test_residual_tail_fusion_host_logic():
    """BatchNormalization -> add -> ReLU(inplace): one xsb_bn_add_relu call reads the BatchNorm INPUT(s) (identity shortcut:
    one parked apply pass; down-sampling shortcut: two); the BatchNorm outputs stay parked and are still produced for a
    later reader; results equal the separate kernels."""
    import numpy as np
    import xshinnosuke_b200 as xs
    from xshinnosuke_b200._lib import lib
    from xshinnosuke_b200.layers import BatchNormalization, ReLU
    rs = np.random.RandomState(3)
    x1 = rs.randn(4, 8, 6, 6).astype(np.float32)
    x2 = rs.randn(4, 8, 6, 6).astype(np.float32)
    g = rs.randn(4, 8, 6, 6).astype(np.float32)
    keep = xs.runtime.fuse_bn_add_relu
    out = {}
    try:
        for two in (False, True):
            for fuse in (True, False):
                xs.runtime.fuse_bn_add_relu = fuse
                np.random.seed(0)
                bn_a, bn_b, relu = BatchNormalization(), BatchNormalization(), ReLU(True)
                ta, tb = xs.tensor(x1, requires_grad=True), xs.tensor(x2, requires_grad=True)
                n0 = len(lib.calls)
                a = bn_a(ta)
                b = bn_b(tb) if two else tb
                z = relu(xs.nn.functional.add(a, b))
                zv = z.eval
                av = a.eval                                  # a later reader of the BatchNorm output: plain apply kernel
                z.backward(gradients=g)
                calls = lib.calls[n0:]
                out[(two, fuse)] = (zv, av, ta.grad.eval, tb.grad.eval, calls)
    finally:
        xs.runtime.fuse_bn_add_relu = keep
    for two in (False, True):
        f, p = out[(two, True)], out[(two, False)]
        assert f[4].count("xsb_bn_add_relu") == 1 and "xsb_add_relu" not in f[4]
        assert f[4].count("xsb_bn_apply") == 1              # only the explicit read of `a` materialised a BatchNorm output
        assert "xsb_bn_add_relu" not in p[4] and p[4].count("xsb_add_relu") == 1
        for u, v in zip(f[:4], p[:4]):
            assert np.abs(u - v).max() <= TOL * max(1.0, np.abs(v).max())


def test_identity_shortcut_gradient_written_by_the_batchnorm_backward():
    """Residual block with an identity shortcut, backward: the masked dZ the shortcut needs is written by the BatchNorm
    backward of the other branch (xsb_bn_bwd_dxsum_side) -- no relu_bwd_mask pass; a reader that gets to the shortcut's
    gradient first still receives it from the plain kernel; results equal the separate kernels."""
    import numpy as np
    import xshinnosuke_b200 as xs
    from xshinnosuke_b200._lib import lib
    from xshinnosuke_b200.layers import BatchNormalization, Conv2D, ReLU
    rs = np.random.RandomState(9)
    x = rs.randn(3, 8, 6, 6).astype(np.float32)
    g = rs.randn(3, 8, 6, 6).astype(np.float32)
    keep = xs.runtime.fuse_add_bwd_side
    out = {}
    try:
        for fuse in (True, False):
            xs.runtime.fuse_add_bwd_side = fuse
            np.random.seed(0)
            pre, conv, bn, relu = ReLU(True), Conv2D(8, 3, padding=1), BatchNormalization(), ReLU(True)
            tx = xs.tensor(x, requires_grad=True)
            n0 = len(lib.calls)
            z0 = pre(BatchNormalization()(tx))               # block input: itself an in-place rectified tensor
            z = relu(xs.nn.functional.add(bn(conv(z0)), z0))
            zv = z.eval
            z.backward(gradients=g)
            calls = lib.calls[n0:]
            out[fuse] = (zv, tx.grad.eval, conv.parameters()[0].grad.eval, calls)
    finally:
        xs.runtime.fuse_add_bwd_side = keep
    f, p = out[True], out[False]
    assert f[3].count("xsb_bn_bwd_dxsum_side") == 1 and p[3].count("xsb_bn_bwd_dxsum_side") == 0
    assert p[3].count("xsb_relu_bwd_mask") == f[3].count("xsb_relu_bwd_mask") + 1
    for u, v in zip(f[:3], p[:3]):
        assert np.abs(u - v).max() <= TOL * max(1.0, np.abs(v).max())
