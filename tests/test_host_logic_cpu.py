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
