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


def test_bn(golden):
    S.check_bn(golden, 1e-5)
    S.check_bn_eval(1e-6)


def test_dense_losses(golden):
    S.check_dense_losses(golden, TOL)


def test_optimizers(golden):
    S.check_optimizers(golden, TOL)


def test_gradient_check_scripts(golden):
    S.check_gradient_check_scripts(golden, 1e-5)


def test_trajectories(golden):
    S.check_test_fc(golden, 1e-6)
    S.check_test_conv(golden, 1e-6, steps=6)


def test_resnet18_small(golden):
    S.check_resnet18(golden, "n2_33", 2, 33, 1e-4, 1e-5)


def test_fit_static_graph(golden):
    S.check_fit_configs(golden, 1e-6)
