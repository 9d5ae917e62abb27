"""Generate tests/golden/*.npz by running the UNMODIFIED reference (NumPy path).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference is imported through the 3-line shim of SURVEY.md 8c (NumPy aliases restored,
both /root/reference and /root/reference/xs on sys.path, ``np`` injected into xs.optim.adam).
Nothing here is imported by the product; the .npz files are what travels to the GPU box.
"""
import os
import sys

import numpy

numpy.int = int
numpy.float = float
numpy.bool = bool
REF = os.environ.get("XS_REFERENCE", "/root/reference")
sys.path[:0] = [REF, os.path.join(REF, "xs")]

import numpy as np  # noqa: E402
import xs  # noqa: E402
import xs.nn as nn  # noqa: E402
import xs.nn.functional as F  # noqa: E402
import xs.optim  # noqa: E402
from xs.layers import (Activation, AvgPooling2D, BatchNormalization, Conv2D, Dense, Flatten,  # noqa: E402
                       Input, MaxPooling2D, ReLU)

sys.modules["xs.optim.adam"].np = np
for _m in list(sys.modules):
    if _m.endswith("optim.adam"):
        sys.modules[_m].np = np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = {}


def put(case, **arrays):
    for k, v in arrays.items():
        OUT[f"{case}/{k}"] = np.array(v)  # copy: .eval aliases live buffers


def T(a, rg=False):
    return xs.tensor(np.array(a), requires_grad=rg)


# ------------------------------------------------------------------ conv2d fwd + bwd
CONV_CASES = [
    # name, N, C, H, W, OC, k, s, p, dtype
    ("k3s1p1", 2, 4, 9, 9, 8, 3, 1, 1, np.float32),
    ("k3s2p1", 3, 8, 11, 11, 16, 3, 2, 1, np.float32),
    ("k1s2p0", 2, 16, 8, 8, 32, 1, 2, 0, np.float32),
    ("k7s2p3_c3", 2, 3, 21, 21, 8, 7, 2, 3, np.float32),
    ("k2s1p0_c1", 4, 1, 28, 28, 8, 2, 1, 0, np.float32),
    ("k3s1p0_f64", 1, 3, 5, 5, 1, 3, 1, 0, np.float64),
    ("k3s1p1_wide", 2, 64, 6, 6, 64, 3, 1, 1, np.float32),
]


def gen_conv():
    rs = np.random.RandomState(1)
    for name, n, c, h, w, oc, k, s, p, dt in CONV_CASES:
        x = rs.randn(n, c, h, w).astype(dt)
        wt = (rs.randn(oc, c, k, k) * 0.1).astype(dt)
        b = rs.randn(1, oc).astype(dt)
        tx, tw, tb = T(x, True), nn.Parameter(wt), nn.Parameter(b)
        y = F.conv2d(tx, tw, tb, (s, s), p)
        yv = y.eval.copy()
        g = rs.randn(*yv.shape).astype(dt)
        y.backward(gradients=g)
        put(f"conv/{name}", x=x, w=wt, b=b, g=g, y=yv, dx=tx.grad.eval, dw=tw.grad.eval, db=tb.grad.eval,
            cfg=np.array([s, p]))


# ------------------------------------------------------------------ max / avg pool
POOL_CASES = [
    # name, N, C, H, W, k, s, p, kind
    ("k3s2p1_relu", 2, 8, 12, 12, 3, 2, 1, "relu"),
    ("k3s2p1_neg", 1, 4, 7, 7, 3, 2, 1, "neg"),
    ("k2s2p0_odd", 3, 8, 27, 27, 2, 2, 0, "randn"),
    ("k3s2p0", 1, 3, 5, 5, 3, 2, 0, "randn"),
    ("k2s1p0", 2, 5, 6, 6, 2, 1, 0, "relu"),
    ("k3s1p1_ties", 2, 4, 6, 6, 3, 1, 1, "ints"),
]


def gen_pool():
    rs = np.random.RandomState(2)
    for name, n, c, h, w, k, s, p, kind in POOL_CASES:
        x = rs.randn(n, c, h, w).astype(np.float32)
        if kind == "relu":
            x = np.maximum(x, 0)
        elif kind == "neg":
            x = -np.abs(x) - 0.5
        elif kind == "ints":
            x = rs.randint(-1, 2, size=x.shape).astype(np.float32)
        tx = T(x, True)
        y = F.max_pool2d(tx, k, s, p)
        yv = y.eval.copy()
        am = y.cache["pool_argmax"].copy()
        g = rs.randn(*yv.shape).astype(np.float32)
        y.backward(gradients=g)
        put(f"maxpool/{name}", x=x, g=g, y=yv, argmax=am, dx=tx.grad.eval, cfg=np.array([k, s, p]))
    for name, n, c, h, w, k, s in [("k2", 2, 8, 4, 4, 2, 2), ("k3s2", 1, 3, 5, 5, 3, 2), ("k2_7x7", 2, 16, 7, 7, 2, 2)]:
        x = rs.randn(n, c, h, w).astype(np.float32)
        tx = T(x, True)
        y = F.avg_pool2d(tx, k, s)
        yv = y.eval.copy()
        g = rs.randn(*yv.shape).astype(np.float32)
        y.backward(gradients=g)
        put(f"avgpool/{name}", x=x, g=g, y=yv, dx=tx.grad.eval, cfg=np.array([k, s, 0]))


# ------------------------------------------------------------------ relu, add
def gen_relu_add():
    rs = np.random.RandomState(3)
    x = rs.randn(3, 5, 4, 4).astype(np.float32)
    x.ravel()[::7] = 0.0  # exact zeros: gradient passes (Q3)
    tx = T(x, True)
    y = F.relu(tx)
    yv = y.eval.copy()
    g = rs.randn(*x.shape).astype(np.float32)
    y.backward(gradients=g)
    put("relu/basic", x=x, g=g, y=yv, dx=tx.grad.eval)
    tx = T(np.array([-1.0, 0.0, 2.0], dtype=np.float32), True)
    y = F.relu(tx)
    y.backward(gradients=np.ones(3, dtype=np.float32))
    put("relu/at_zero", dx=tx.grad.eval)
    a = rs.randn(2, 3, 4, 4).astype(np.float32)
    b = rs.randn(2, 3, 4, 4).astype(np.float32)
    ta, tb = T(a, True), T(b, True)
    y = F.add(ta, tb)
    yv = y.eval.copy()
    g = rs.randn(*a.shape).astype(np.float32)
    y.backward(gradients=g)
    put("add/basic", a=a, b=b, g=g, y=yv, da=ta.grad.eval, db=tb.grad.eval)


# ------------------------------------------------------------------ batch norm
def gen_bn():
    rs = np.random.RandomState(4)
    for name, shape, eps, mom in [("4d", (4, 6, 5, 5), 1e-5, 0.9), ("4d_layer_defaults", (8, 16, 3, 3), 1e-6, 0.99),
                                  ("2d", (16, 10), 1e-5, 0.9), ("4d_offset", (2, 3, 7, 7), 1e-5, 0.9)]:
        c = shape[1]
        x = rs.randn(*shape).astype(np.float32)
        if name == "4d_offset":
            x = x * 0.01 + 100.0  # exposes E[x^2]-E[x]^2 cancellation
        gm = (rs.rand(c) + 0.5).astype(np.float32)
        bt = rs.randn(c).astype(np.float32)
        mm0 = rs.randn(c).astype(np.float32)
        mv0 = (rs.rand(c) + 0.5).astype(np.float32)
        tx, tg, tb = T(x, True), nn.Parameter(gm), nn.Parameter(bt)
        tmm, tmv = T(mm0.copy()), T(mv0.copy())
        y = F.batch_norm(tx, tg, tb, tmm, tmv, 1, True, eps, mom)
        yv = y.eval.copy()
        g = rs.randn(*shape).astype(np.float32)
        y.backward(gradients=g)
        put(f"bn/{name}", x=x, gamma=gm, beta=bt, mm0=mm0, mv0=mv0, g=g, y=yv, mm1=tmm.eval, mv1=tmv.eval,
            dx=tx.grad.eval, dgamma=tg.grad.eval, dbeta=tb.grad.eval, cfg=np.array([eps, mom]))


# ------------------------------------------------------------------ dense, losses
def gen_dense_loss():
    rs = np.random.RandomState(5)
    for name, n, i, o in [("n8", 8, 20, 7), ("n1", 1, 6, 4), ("mlp", 16, 784, 500)]:
        x = rs.randn(n, i).astype(np.float32)
        w = (rs.randn(i, o) * 0.1).astype(np.float32)
        b = rs.randn(1, o).astype(np.float32)
        tx, tw, tb = T(x, True), nn.Parameter(w), nn.Parameter(b)
        y = F.addmm(tb, tx, tw)
        yv = y.eval.copy()
        g = rs.randn(n, o).astype(np.float32)
        y.backward(gradients=g)
        put(f"addmm/{name}", x=x, w=w, b=b, g=g, y=yv, dx=tx.grad.eval, dw=tw.grad.eval, db=tb.grad.eval)
    for name, n, c in [("n8c10", 8, 10), ("n32c100", 32, 100), ("n1c3", 1, 3)]:
        z = (rs.randn(n, c) * 3).astype(np.float32)
        lab = rs.randint(0, c, (n,))
        tz, tl = T(z, True), T(lab)
        loss = F.cross_entropy(tz, tl)
        lv = loss.eval.copy()
        loss.backward()
        put(f"xent/{name}", z=z, labels=lab, loss=lv, dz=tz.grad.eval)
    a = rs.randn(4, 9).astype(np.float32)
    b = rs.randn(4, 9).astype(np.float32)
    ta, tb = T(a, True), T(b)
    loss = F.mse_loss(ta, tb)
    lv = loss.eval.copy()
    loss.grad_fn = F.MSELossBackward if loss.grad_fn is None else loss.grad_fn
    put("mse/basic", a=a, b=b, loss=lv)


# ------------------------------------------------------------------ optimizers
def gen_optim():
    rs = np.random.RandomState(6)
    shapes = [(5, 3, 3, 3), (1, 5), (7,), (20, 4)]
    p0 = [rs.randn(*s) for s in shapes]
    gs = [[rs.randn(*s) for s in shapes] for _ in range(4)]
    import xs.optim.sgd, xs.optim.rmsprop, xs.optim.adagrad, xs.optim.adadelta  # noqa: E401
    for m in ("sgd", "rmsprop", "adagrad", "adadelta"):
        sys.modules["xs.optim." + m].np = np      # same missing-import defect as adam.py
    for oname, ctor in [("sgd", lambda ps: xs.optim.SGD(ps, lr=0.1)), ("adam", lambda ps: xs.optim.Adam(ps, lr=0.01)),
                        ("momentum", lambda ps: xs.optim.Momentum(ps, lr=0.05)),
                        ("rmsprop", lambda ps: xs.optim.RMSprop(ps, lr=0.01)),
                        ("adagrad", lambda ps: xs.optim.AdaGrad(ps, lr=0.05)),
                        ("adadelta", lambda ps: xs.optim.AdaDelta(ps))]:
        ps = [nn.Parameter(p.copy()) for p in p0]
        opt = ctor(ps)
        for step in range(4):
            opt.zero_grad()
            for p, g in zip(ps, gs[step]):
                p.grad.eval[...] = g
            opt.step()
            put(f"optim/{oname}/step{step}", **{f"p{i}": p.eval for i, p in enumerate(ps)})
    put("optim/init", **{f"p{i}": p for i, p in enumerate(p0)})
    for step in range(4):
        put(f"optim/grads{step}", **{f"g{i}": g for i, g in enumerate(gs[step])})


# ------------------------------------------------------------------ reference test scripts as recipes
def gen_check_scripts():
    # tests/gradient_check/check_cnn.py:7-28
    xs.manual_seed_all(0)
    x = xs.randn(1, 3, 5, 5, requires_grad=True)
    y = xs.randn(1, 9)
    cnn = nn.Sequential(Conv2D(out_channels=1, kernel_size=3, use_bias=True), Flatten())
    out = cnn(x)
    loss = nn.MSELoss()(out, y)
    lv = loss.eval.copy()
    loss.backward()
    ps = sorted(cnn.parameters(), key=lambda p: -p.eval.size)
    put("check_cnn", x=x.eval, y=y.eval, w=ps[0].eval, b=ps[1].eval, loss=lv, dw=ps[0].grad.eval, db=ps[1].grad.eval,
        dx=x.grad.eval)
    # tests/gradient_check/check_dnn.py:7-23
    xs.manual_seed_all(0)
    x = xs.randn(3, 10, requires_grad=True)
    y = xs.randn(3, 2)
    dnn = Dense(out_features=2)
    out = dnn(x)
    loss = nn.MSELoss()(out, y)
    lv = loss.eval.copy()
    loss.backward()
    w_, b_ = dnn.parameters()
    put("check_dnn", x=x.eval, y=y.eval, w=w_.eval, b=b_.eval, loss=lv, dw=w_.grad.eval, db=b_.grad.eval,
        dx=x.grad.eval)
    # tests/gradient_check/check_bn.py:7-28
    xs.manual_seed_all(0)
    x = xs.randn(1, 3, 5, 5, requires_grad=True)
    y = xs.randn(1, 75)
    bn_layer = BatchNormalization()
    bn = nn.Sequential(bn_layer, Flatten())
    out = bn(x)
    loss = nn.MSELoss()(out, y)
    lv = loss.eval.copy()
    loss.backward()
    gam, bet = bn_layer.parameters()
    put("check_bn", x=x.eval, y=y.eval, loss=lv, dgamma=gam.grad.eval, dbeta=bet.grad.eval, dx=x.grad.eval)


def gen_test_conv():
    """tests/test_conv.py:58-99 (XS half): Conv-BN-ReLU-Conv, view/mean, cross_entropy, Adam lr 0.1."""
    np.random.seed(0)
    xd = np.random.rand(2, 1, 6, 6).astype(np.float32)
    yd = np.random.randint(0, 2, (2,), dtype=np.int64)
    w1 = np.random.rand(3, 1, 3, 3).astype(np.float32)
    b1 = np.random.rand(3).astype(np.float32)
    g1 = np.random.rand(3).astype(np.float32)
    be1 = np.random.rand(3).astype(np.float32)
    w2 = np.random.rand(5, 3, 3, 3).astype(np.float32)
    b2 = np.random.rand(5).astype(np.float32)
    l1 = Conv2D(out_channels=3, kernel_size=3)
    bn1 = BatchNormalization(epsilon=1e-5, momentum=0.9)
    l2 = Conv2D(out_channels=5, kernel_size=3)
    # Parameter() aliases the ndarray and training updates it in place: hand over copies
    l1.parameters([nn.Parameter(w1.copy()), nn.Parameter(b1.copy())])
    l2.parameters([nn.Parameter(w2.copy()), nn.Parameter(b2.copy())])
    bn1.parameters([nn.Parameter(g1.copy()), nn.Parameter(be1.copy())])
    bn1.moving_mean = xs.zeros(3)
    bn1.moving_variance = xs.ones(3)
    net = nn.Sequential(l1, bn1, ReLU(), l2)
    opt = xs.optim.Adam(net.parameters(), lr=0.1)
    losses = []
    for _ in range(50):
        x, y = xs.tensor(xd), xs.tensor(yd)
        opt.zero_grad()
        pred = net(x).view(x.size(0), -1, 2).mean(axis=1)
        loss = F.cross_entropy(pred, y)
        losses.append(loss.item())
        loss.backward()
        opt.step()
    put("test_conv", x=xd, y=yd, w1=w1, b1=b1, g1=g1, be1=be1, w2=w2, b2=b2, losses=np.array(losses))


def gen_test_fc_run():
    """tests/test_fc.py:44-72 (XS half): Dense(5)-ReLU-Dense(2), cross_entropy, SGD lr 0.5, f64, 3 steps."""
    np.random.seed(0)
    xd = np.random.rand(2, 10)
    yd = np.random.randint(0, 2, (2,), dtype=np.int64)
    w1 = np.random.rand(10, 5)
    b1 = np.random.rand(1, 5)
    w2 = np.random.rand(5, 2)
    b2 = np.random.rand(1, 2)
    d1, d2 = Dense(5, input_shape=(10,)), Dense(2)
    d1.parameters([nn.Parameter(w1.copy()), nn.Parameter(b1.copy())])
    d2.parameters([nn.Parameter(w2.copy()), nn.Parameter(b2.copy())])
    net = nn.Sequential(d1, ReLU(), d2)
    opt = xs.optim.SGD(net.parameters(), lr=0.5)
    losses = []
    for _ in range(3):
        x, y = xs.tensor(xd), xs.tensor(yd)
        opt.zero_grad()
        loss = F.cross_entropy(net(x), y)
        losses.append(loss.item())
        loss.backward()
        opt.step()
    put("test_fc", x=xd, y=yd, w1=w1, b1=b1, w2=w2, b2=b2, losses=np.array(losses),
        w1_end=d1.parameters()[0].eval, w2_end=d2.parameters()[0].eval)


# ------------------------------------------------------------------ ResNet-18 (tests/resnet/resnet18_dg.py)
def gen_resnet():
    src = open(os.path.join(REF, "tests/resnet/resnet18_dg.py")).read()
    ns = {"nn": nn, "Conv2D": Conv2D, "BatchNormalization": BatchNormalization, "ReLU": ReLU,
          "MaxPooling2D": MaxPooling2D, "AvgPooling2D": AvgPooling2D, "Flatten": Flatten, "Dense": Dense}
    exec(src[src.index("class BasicBlock"):src.index("# random generate data")], ns)
    for tag, n, hw in [("n32_56", 32, 56), ("n4_56", 4, 56), ("n2_33", 2, 33)]:
        np.random.seed(0)
        xd = np.random.rand(n, 3, hw, hw)
        yd = np.random.randint(0, 100, (n,))
        net = ns["ResNet18"]()
        opt = xs.optim.SGD(net.parameters(), lr=0.1)
        crit = nn.CrossEntropyLoss()
        losses, first_logits = [], None
        for step in range(3):
            x, y = xs.tensor(xd), xs.tensor(yd)
            opt.zero_grad()
            pred = net(x)
            if step == 0:
                first_logits = pred.eval.copy()
            loss = crit(pred, y)
            losses.append(loss.item())
            loss.backward()
            opt.step()
        put(f"resnet18/{tag}", losses=np.array(losses), logits0=first_logits,
            nparams=np.array([sum(p.eval.size for p in net.parameters()), len(net.parameters())]))


# ------------------------------------------------------------------ BASELINE.json configs[0], configs[1] via Model/Sequential.fit
def gen_fit_configs():
    # configs[0] in its working form: Input-Conv2D(8,2x2)-Activation(relu)-MaxPooling2D(2)-Flatten-Dense(10)
    np.random.seed(0)
    xd = np.random.rand(96, 1, 28, 28)
    yd = np.random.randint(0, 10, (96,))
    np.random.seed(1)
    inp = Input((1, 28, 28))
    h = Conv2D(8, (2, 2))(inp)
    h = Activation("relu")(h)
    h = MaxPooling2D(2)(h)
    h = Flatten()(h)
    out = Dense(10)(h)
    model = nn.Model(inputs=inp, outputs=out)
    model.compile(optimizer="sgd", loss="cross_entropy", lr=0.01)
    prof = model.fit(xd, yd, batch_size=32, epochs=2, verbose=0, shuffle=False)
    put("fit/cnn", losses=np.array(_profile_losses(prof)))
    # configs[1]: Sequential MLP 784-500(relu)-10
    np.random.seed(0)
    xd = np.random.rand(80, 784)
    yd = np.random.randint(0, 10, (80,))
    np.random.seed(1)
    mlp = nn.Sequential(Dense(500, activation="relu", input_shape=(784,)), Dense(10))
    mlp.compile(optimizer="sgd", loss="cross_entropy", lr=0.05)
    prof = mlp.fit(xd, yd, batch_size=32, epochs=2, verbose=0, shuffle=False)
    put("fit/mlp", losses=np.array(_profile_losses(prof)))


# ------------------------------------------------------------------ SURVEY 8f-4 ops: sigmoid / tanh / pad / dropout / LN / GN / broadcast add
def gen_extra_ops():
    rs = np.random.RandomState(21)
    for name, fn in (("sigmoid", F.sigmoid), ("tanh", F.tanh)):
        x = rs.randn(3, 5, 4, 6).astype(np.float32)
        tx = T(x, True)
        y = fn(tx)
        yv = y.eval.copy()
        g = rs.randn(*yv.shape).astype(np.float32)
        y.backward(gradients=g)
        put(f"extra/{name}", x=x, g=g, y=yv, dx=tx.grad.eval)
    # zero padding
    x = rs.randn(2, 3, 5, 4).astype(np.float32)
    tx = T(x, True)
    y = F.pad2d(tx, (2, 1))
    yv = y.eval.copy()
    g = rs.randn(*yv.shape).astype(np.float32)
    y.backward(gradients=g)
    put("extra/pad2d", x=x, g=g, y=yv, dx=tx.grad.eval)
    # dropout: the mask the reference drew is stored; parity is defined given the mask
    x = rs.randn(4, 6, 3, 3).astype(np.float32)
    tx = T(x, True)
    np.random.seed(5)
    y = F.dropout2d(tx, 0.7)
    mask = np.array(y.cache["mask"])
    yv = y.eval.copy()
    # (the shipped Dropout2DBackward multiplies the int64 mask in place by keep_prob and raises a casting error,
    # grad_fn.py:323: only the forward can be pinned; the backward formula is its own: dX += mask * keep_prob * dY)
    put("extra/dropout", x=x, y=yv, mask=mask, keep_prob=np.array(0.7))
    # layer_norm / group_norm cannot be pinned: as shipped both raise in the FORWARD (np.subtract(x, mean, out=mean) with a
    # keepdims mean, functional.py:754 / :789) -- parity for them is against oracle/xs_numpy.py's restatement of the formulas
    for fn, args in ((F.layer_norm, (T(rs.randn(2, 3, 2, 2)), T(np.ones((3, 2, 2))), T(np.zeros((3, 2, 2))))),
                     (F.group_norm, (T(rs.randn(2, 4, 2, 2)), T(np.ones(4)), T(np.zeros(4)), 1e-5, 2))):
        try:
            fn(*args)
            raise SystemExit(f"{fn.__name__} runs on this reference: pin it")
        except ValueError:
            pass
    # broadcasting add: [M, N] + [1, N]
    a, b = rs.randn(6, 5).astype(np.float32), rs.randn(1, 5).astype(np.float32)
    ta, tb = T(a, True), T(b, True)
    y = F.add(ta, tb)
    yv = y.eval.copy()
    g = rs.randn(6, 5).astype(np.float32)
    y.backward(gradients=g)
    put("extra/add_broadcast", a=a, b=b, g=g, y=yv, da=ta.grad.eval, db=tb.grad.eval)


def _profile_losses(prof):
    return list(prof.data["train_loss"])  # xs/utils/toolkit.py:54-56


if __name__ == "__main__":
    if "--only-extra" in sys.argv:      # add the 8f-4 arrays to the existing file without re-running the long cases
        path = os.path.join(HERE, "reference_vectors.npz")
        old = np.load(path)
        OUT.update({k: old[k] for k in old.files if not k.startswith("extra/")})
        gen_extra_ops()
        np.savez_compressed(path, **OUT)
        print("wrote", path, len(OUT), "arrays", os.path.getsize(path) // 1024, "KiB")
        sys.exit(0)
    gen_conv()
    gen_pool()
    gen_relu_add()
    gen_bn()
    gen_dense_loss()
    gen_optim()
    gen_check_scripts()
    gen_test_conv()
    gen_test_fc_run()
    gen_resnet()
    gen_fit_configs()
    gen_extra_ops()
    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **OUT)
    print("wrote", path, len(OUT), "arrays", os.path.getsize(path) // 1024, "KiB")
