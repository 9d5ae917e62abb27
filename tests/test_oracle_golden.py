"""CPU: pin the oracle (oracle/xs_numpy.py, oracle/resnet18.py) against vectors produced by the
unmodified reference (tests/golden/make_golden.py). Tolerance is a few float64 ulps: the oracle
restates the same algorithm in the same precision."""
import numpy as np
import pytest

from oracle import xs_numpy as X
from oracle.resnet18 import ResNet18, synthetic_batch

TIGHT = 1e-12


def close(a, b, tol=TIGHT):
    assert np.asarray(a).shape == np.asarray(b).shape, (np.asarray(a).shape, np.asarray(b).shape)
    assert X.rel_err(a, b) <= tol, X.rel_err(a, b)


def test_conv_cases(golden):
    names = golden.names("conv")
    assert len(names) >= 7
    for name in names:
        c = golden.case(f"conv/{name}")
        s, p = int(c["cfg"][0]), int(c["cfg"][1])
        y, col = X.conv2d_fwd(c["x"], c["w"], c["b"], (s, s), p)
        assert y.dtype == c["y"].dtype
        # fp32 outputs are a cast of the float64 product: allow 1 fp32 ulp
        tol = 2e-7 if y.dtype == np.float32 else TIGHT
        close(y, c["y"], tol)
        dx, dw, db = X.conv2d_bwd(c["g"], c["x"].shape, c["w"], col, (s, s), p)
        close(dx, c["dx"], tol)
        close(dw, c["dw"], tol)
        close(db.reshape(c["db"].shape), c["db"], tol)


def test_maxpool_cases(golden):
    for name in golden.names("maxpool"):
        c = golden.case(f"maxpool/{name}")
        k, s, p = (int(v) for v in c["cfg"])
        y, am = X.max_pool2d_fwd(c["x"], k, s, p)
        assert np.array_equal(y, c["y"]), name
        assert am.dtype == np.int64 and np.array_equal(am, c["argmax"]), name
        dx = X.max_pool2d_bwd(c["g"], am, c["x"].shape, k, s, p)
        close(dx, c["dx"], 2e-7)


def test_maxpool_zero_padding_quirk():
    # all-negative input with padding -> border windows see the 0 pad
    y, am = X.max_pool2d_fwd(-np.ones((1, 1, 4, 4), dtype=np.float32), 3, 2, 1)
    assert y.ravel().tolist() == [0.0, 0.0, 0.0, -1.0]


def test_avgpool_cases(golden):
    for name in golden.names("avgpool"):
        c = golden.case(f"avgpool/{name}")
        k, s, _ = (int(v) for v in c["cfg"])
        close(X.avg_pool2d_fwd(c["x"], k, s), c["y"], 2e-7)
        close(X.avg_pool2d_bwd(c["g"], c["x"].shape, k, s), c["dx"], 2e-7)


def test_relu_add(golden):
    c = golden.case("relu/basic")
    assert np.array_equal(X.relu_fwd(c["x"]), c["y"])
    assert np.array_equal(X.relu_bwd(c["g"], c["x"]), c["dx"])
    assert golden.case("relu/at_zero")["dx"].tolist() == [0.0, 1.0, 1.0]
    assert X.relu_bwd(np.ones(3, np.float32), np.array([-1.0, 0.0, 2.0], np.float32)).tolist() == [0.0, 1.0, 1.0]
    c = golden.case("add/basic")
    assert np.array_equal(c["a"] + c["b"], c["y"])
    assert np.array_equal(c["da"], c["g"]) and np.array_equal(c["db"], c["g"])


def test_batch_norm_cases(golden):
    for name in golden.names("bn"):
        c = golden.case(f"bn/{name}")
        eps, mom = float(c["cfg"][0]), float(c["cfg"][1])
        y, xhat, sv, mm1, mv1 = X.batch_norm_fwd(c["x"], c["gamma"], c["beta"], c["mm0"], c["mv0"], 1, eps, mom)
        close(y, c["y"], 1e-6)
        close(mm1, c["mm1"], 1e-6)
        close(mv1, c["mv1"], 1e-6)
        dx, dg, db = X.batch_norm_bwd(c["g"], xhat, sv, c["gamma"], 1)
        close(dx, c["dx"], 2e-5 if name == "4d_offset" else 2e-6)
        close(dg, c["dgamma"], 2e-6)
        close(db, c["dbeta"], 2e-6)


def test_dense_and_losses(golden):
    for name in golden.names("addmm"):
        c = golden.case(f"addmm/{name}")
        close(X.addmm_fwd(c["b"], c["x"], c["w"]), c["y"], 1e-6)
        dx, dw, db = X.addmm_bwd(c["g"], c["x"], c["w"], c["b"].shape)
        close(dx, c["dx"], 1e-6)
        close(dw, c["dw"], 1e-6)
        close(db.reshape(c["db"].shape), c["db"], 1e-6)
    for name in golden.names("xent"):
        c = golden.case(f"xent/{name}")
        loss, probs = X.cross_entropy_fwd(c["z"], c["labels"])
        close(loss, c["loss"], 1e-6)
        close(X.cross_entropy_bwd(probs, c["labels"]), c["dz"], 1e-6)
    c = golden.case("mse/basic")
    close(X.mse_loss_fwd(c["a"], c["b"]), c["loss"], 1e-6)


def test_optimizers(golden):
    init = golden.case("optim/init")
    n = len(init)
    for opt in ("sgd", "adam", "momentum", "rmsprop", "adagrad", "adadelta"):
        ps = [init[f"p{i}"].copy() for i in range(n)]
        state = {}
        for step in range(4):
            g = golden.case(f"optim/grads{step}")
            gs = [g[f"g{i}"] for i in range(n)]
            if opt == "sgd":
                X.sgd_step(ps, gs, 0.1)
            elif opt == "adam":
                X.adam_step(ps, gs, state, lr=0.01)
            elif opt == "momentum":
                X.momentum_step(ps, gs, state, lr=0.05)
            elif opt == "rmsprop":
                X.rmsprop_step(ps, gs, state, lr=0.01)
            elif opt == "adagrad":
                X.adagrad_step(ps, gs, state, lr=0.05)
            else:
                X.adadelta_step(ps, gs, state)
            want = golden.case(f"optim/{opt}/step{step}")
            for i in range(n):
                close(ps[i], want[f"p{i}"])


def test_check_scripts_known_answers(golden):
    # BASELINE.md section 4 values, reproduced by running the reference here
    c = golden.case("check_cnn")
    np.testing.assert_allclose(c["db"], [[0.55916639]], rtol=1e-7)
    y, col = X.conv2d_fwd(c["x"], c["w"], c["b"])
    out = y.reshape(1, -1)
    close(X.mse_loss_fwd(out, c["y"]), c["loss"])
    g = X.mse_loss_bwd(out, c["y"]).reshape(y.shape)
    dx, dw, db = X.conv2d_bwd(g, c["x"].shape, c["w"], col)
    close(dw, c["dw"])
    close(db.reshape(1, 1), c["db"])
    close(dx, c["dx"])
    c = golden.case("check_dnn")
    np.testing.assert_allclose(c["db"], [[-0.01016655, 0.10660269]], rtol=1e-6)
    c = golden.case("check_bn")
    np.testing.assert_allclose(c["dgamma"], [0.31449195, 0.41223045, 0.33131135], rtol=1e-7)
    np.testing.assert_allclose(c["dbeta"], [-0.08045787, -0.13519585, -0.03582737], rtol=1e-6)


def test_trajectory_known_answers(golden):
    np.testing.assert_allclose(golden.case("test_fc")["losses"], [0.7644428, 4.02763008, 8.16782015], rtol=1e-8)
    l = golden.case("test_conv")["losses"]
    np.testing.assert_allclose(l[:4], [2.19681621, 0.86421591, 0.29898247, 0.1250557], rtol=1e-6)
    np.testing.assert_allclose(l[49], 1.6033873e-05, rtol=1e-5)


@pytest.mark.parametrize("tag,n,hw", [("n4_56", 4, 56), ("n2_33", 2, 33)])
def test_resnet18_trajectory_small(golden, tag, n, hw):
    c = golden.case(f"resnet18/{tag}")
    x, y = synthetic_batch(n, hw)
    net = ResNet18()
    logits = net.forward(x)
    close(logits, c["logits0"], 1e-10)
    losses = [net.train_step(x, y, 0.1) for _ in range(3)]
    np.testing.assert_allclose(losses, c["losses"], rtol=1e-9)
    assert [sum(p.size for p in net.params), len(net.params)] == c["nparams"].tolist()


def test_resnet18_n32_recorded(golden):
    c = golden.case("resnet18/n32_56")
    # the recipe of BASELINE.md section 4, as the reference actually evaluates it in this container
    np.testing.assert_allclose(c["losses"], [6.642040132035239, 0.836408809753542, 0.020209142234972645], rtol=1e-12)
    assert c["nparams"].tolist() == [11232612, 82]


def test_small_net_oracles_reproduce_the_reference_fit_trajectories(golden):
    """oracle/small_nets.py (BASELINE configs[0], [1]; the CPU arm of bench.py --workload cnn|mlp) against the losses the
    unmodified reference's Model.fit / Sequential.fit produced (tests/golden/make_golden.py gen_fit_configs)."""
    from oracle.small_nets import ReadmeCNN, ReadmeMLP
    np.random.seed(0)
    xd, yd = np.random.rand(96, 1, 28, 28), np.random.randint(0, 10, (96,))
    np.random.seed(1)
    net = ReadmeCNN()
    losses = [net.train_step(xd[i:i + 32], yd[i:i + 32], 0.01) for _ in range(2) for i in range(0, 96, 32)]
    np.testing.assert_allclose(losses, golden.case("fit/cnn")["losses"], rtol=1e-12)
    np.random.seed(0)
    xd, yd = np.random.rand(80, 784), np.random.randint(0, 10, (80,))
    np.random.seed(1)
    net = ReadmeMLP()
    losses = [net.train_step(xd[i:i + 32], yd[i:i + 32], 0.05) for _ in range(2) for i in range(0, 80, 32)]
    np.testing.assert_allclose(losses, golden.case("fit/mlp")["losses"], rtol=1e-12)


def test_extra_op_oracles_against_the_reference(golden):
    """oracle restatements of the SURVEY 8f-4 ops against what the unmodified reference produced (sigmoid, tanh, pad2d,
    dropout forward given the reference's mask, broadcasting add). layer_norm / group_norm are unpinned: the reference's
    own implementations raise as shipped (asserted by tests/golden/make_golden.py)."""
    from oracle import xs_numpy as X
    for name, f, b in (("sigmoid", X.sigmoid_fwd, X.sigmoid_bwd), ("tanh", X.tanh_fwd, X.tanh_bwd)):
        c = golden.case(f"extra/{name}")
        y = f(c["x"].astype(np.float64))
        assert X.rel_err(y, c["y"]) <= 1e-6
        assert X.rel_err(b(c["g"].astype(np.float64), y), c["dx"]) <= 1e-6
    c = golden.case("extra/pad2d")
    assert np.array_equal(X.pad2d_fwd(c["x"], (2, 1)), c["y"]) and np.array_equal(X.pad2d_bwd(c["g"], (2, 1)), c["dx"])
    c = golden.case("extra/dropout")
    assert X.rel_err(X.dropout_fwd(c["x"].astype(np.float64), c["mask"], float(c["keep_prob"])), c["y"]) <= 1e-7
    c = golden.case("extra/add_broadcast")
    assert X.rel_err(c["a"] + c["b"], c["y"]) <= 1e-7
    assert X.rel_err(X.add_broadcast_bwd(c["g"].astype(np.float64), c["b"].shape).reshape(-1), c["db"].reshape(-1)) <= 1e-6
