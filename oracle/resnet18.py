"""Oracle ResNet-18 training step on NumPy (TEST INFRASTRUCTURE ONLY).

Restates, on top of ``oracle.xs_numpy``, the network of the reference's
``tests/resnet/resnet18_dg.py:12-75`` (BasicBlock 12-34, ResNet18 37-75) and the
training step of ``resnet18_dg.py:92-100`` (zero_grad, forward, cross-entropy, backward,
SGD/Adam step). Parameter creation order equals the reference's lazy first-forward order
so that ``np.random.seed(s)`` yields the same weights: conv kernels/Dense kernel are
``normal(0, 0.1)`` (xs/nn/initializers.py:60-67), biases zeros, BN gamma ones / beta zeros
/ moving stats zeros / ones (xs/layers/normalize.py:21-33), BN eps 1e-6, momentum 0.99.
"""
import numpy as np

from . import xs_numpy as X

BN_EPS = 1e-6
BN_MOM = 0.99


class _Conv:
    def __init__(self, cin, cout, k, stride, pad, params):
        self.w = np.random.normal(loc=0.0, scale=0.1, size=(cout, cin, k, k))
        self.b = np.zeros((1, cout))
        self.stride, self.pad = (stride, stride), pad
        params += [self.w, self.b]

    def fwd(self, x):
        self.x_shape = x.shape
        y, self.col = X.conv2d_fwd(x, self.w, self.b, self.stride, self.pad)
        return y

    def bwd(self, dy, grads, need_dx=True):
        dx, dw, db = X.conv2d_bwd(dy, self.x_shape, self.w, self.col, self.stride, self.pad, need_dx)
        self.col = None
        grads[id(self.w)] = dw
        grads[id(self.b)] = db.reshape(1, -1)
        return dx


class _BN:
    def __init__(self, c, params):
        self.g = np.ones(c)
        self.b = np.zeros(c)
        self.mm = np.zeros(c)
        self.mv = np.ones(c)
        params += [self.g, self.b]

    def fwd(self, x):
        y, self.xhat, self.sv, self.mm, self.mv = X.batch_norm_fwd(
            x, self.g, self.b, self.mm, self.mv, 1, BN_EPS, BN_MOM)
        return y

    def bwd(self, dy, grads):
        dx, dg, db = X.batch_norm_bwd(dy, self.xhat, self.sv, self.g, 1)
        self.xhat = self.sv = None
        grads[id(self.g)] = dg
        grads[id(self.b)] = db
        return dx


class _Block:
    """resnet18_dg.py:12-34: conv3x3(s)-BN-ReLU-conv3x3-BN (+ 1x1(s)-BN shortcut) - add - ReLU."""

    def __init__(self, cin, cout, stride, params):
        self.c1 = _Conv(cin, cout, 3, stride, 1, params)
        self.b1 = _BN(cout, params)
        self.c2 = _Conv(cout, cout, 3, 1, 1, params)
        self.b2 = _BN(cout, params)
        self.proj = stride != 1 or cin != cout
        if self.proj:
            self.cs = _Conv(cin, cout, 1, stride, 0, params)
            self.bs = _BN(cout, params)

    def fwd(self, x):
        self.pre1 = self.b1.fwd(self.c1.fwd(x))
        h = self.b2.fwd(self.c2.fwd(X.relu_fwd(self.pre1)))
        s = self.bs.fwd(self.cs.fwd(x)) if self.proj else x
        self.pre2 = h + s
        return X.relu_fwd(self.pre2)

    def bwd(self, dy, grads):
        d = X.relu_bwd(dy, self.pre2)
        dh = self.c2.bwd(self.b2.bwd(d, grads), grads)
        dx = self.c1.bwd(self.b1.bwd(X.relu_bwd(dh, self.pre1), grads), grads)
        dx = dx + (self.cs.bwd(self.bs.bwd(d, grads), grads) if self.proj else d)
        self.pre1 = self.pre2 = None
        return dx


class ResNet18:
    """Build AFTER seeding np.random (weights are drawn at construction, in forward order)."""

    def __init__(self, in_channels=3, num_classes=100, feat=None):
        self.params = []
        p = self.params
        self.stem = _Conv(in_channels, 64, 7, 2, 3, p)
        self.stem_bn = _BN(64, p)
        self.blocks = []
        cin = 64
        for cout, stride in ((64, 1), (128, 2), (256, 2), (512, 2)):
            for s in (stride, 1):
                self.blocks.append(_Block(cin, cout, s, p))
                cin = cout
        self.num_classes = num_classes
        self.fc_w = None  # drawn lazily: in-features depend on the input size
        self.fc_b = None
        if feat is not None:
            self._init_fc(feat)

    def _init_fc(self, feat):
        self.fc_w = np.random.normal(loc=0.0, scale=0.1, size=(feat, self.num_classes))
        self.fc_b = np.zeros((1, self.num_classes))
        self.params += [self.fc_w, self.fc_b]

    def forward(self, x):
        self.pre0 = self.stem_bn.fwd(self.stem.fwd(x))
        a = X.relu_fwd(self.pre0)
        self.pool_in_shape = a.shape
        a, self.argmax = X.max_pool2d_fwd(a, 3, (2, 2), 1)
        for blk in self.blocks:
            a = blk.fwd(a)
        self.avg_in_shape = a.shape
        a = X.avg_pool2d_fwd(a, 2)
        self.flat_shape = a.shape
        self.feat = a.reshape(a.shape[0], -1)
        if self.fc_w is None:
            self._init_fc(self.feat.shape[1])
        return X.addmm_fwd(self.fc_b, self.feat, self.fc_w)

    def backward(self, dlogits):
        grads = {}
        dfeat, dw, db = X.addmm_bwd(dlogits, self.feat, self.fc_w, self.fc_b.shape)
        grads[id(self.fc_w)] = dw
        grads[id(self.fc_b)] = db.reshape(self.fc_b.shape)
        d = X.avg_pool2d_bwd(dfeat.reshape(self.flat_shape), self.avg_in_shape, 2)
        for blk in reversed(self.blocks):
            d = blk.bwd(d, grads)
        d = X.max_pool2d_bwd(d, self.argmax, self.pool_in_shape, 3, (2, 2), 1)
        d = X.relu_bwd(d, self.pre0)
        self.stem.bwd(self.stem_bn.bwd(d, grads), grads, need_dx=False)
        return [grads[id(p)] for p in self.params]

    def train_step(self, x, y, lr=0.1, optimizer="sgd", adam_state=None):
        """One step of resnet18_dg.py:92-100; returns the float64 loss before the update."""
        logits = self.forward(x)
        loss, probs = X.cross_entropy_fwd(logits, y)
        grads = self.backward(X.cross_entropy_bwd(probs, y))
        if optimizer == "sgd":
            X.sgd_step(self.params, grads, lr)
        else:
            X.adam_step(self.params, grads, adam_state, lr=lr)
        return float(loss[0])


def synthetic_batch(n, hw=56, classes=100, seed=0):
    """resnet18_dg.py:79-82 recipe: seed, X = rand(n,3,hw,hw), Y = randint(0,classes,(n,))."""
    np.random.seed(seed)
    x = np.random.rand(n, 3, hw, hw)
    y = np.random.randint(0, classes, (n,))
    return x, y
