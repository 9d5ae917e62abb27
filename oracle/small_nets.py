"""Oracle restatement of BASELINE.json configs[0] / configs[1] on NumPy (TEST INFRASTRUCTURE ONLY).

configs[0], in the form that runs on the reference (README.md:53-68 recipe):
    Input(1,28,28) -> Conv2D(8, 2x2) -> Activation('relu') -> MaxPooling2D(2) -> Flatten -> Dense(10), SGD, cross entropy
configs[1] (README.md:94):  Sequential Dense(784 -> 500, relu) -> Dense(10), SGD, cross entropy.
One train step = the body of xs/nn/models.py:100-106 (zero_grad, forward, loss, backward, optimizer step) written with
``oracle.xs_numpy``. Parameters are drawn at construction in the reference's creation order -- kernels normal(0, 0.1)
(xs/nn/initializers.py:60-67), biases zeros -- so ``np.random.seed(s)`` reproduces the reference's weights; pinned to the
reference's own fit() trajectories (tests/golden: fit/cnn, fit/mlp) by tests/test_oracle_golden.py.
"""
import numpy as np

from . import xs_numpy as X


class ReadmeCNN:
    def __init__(self):
        self.w = np.random.normal(loc=0.0, scale=0.1, size=(8, 1, 2, 2))
        self.b = np.zeros((1, 8))
        self.fw = np.random.normal(loc=0.0, scale=0.1, size=(8 * 13 * 13, 10))
        self.fb = np.zeros((1, 10))
        self.params = [self.w, self.b, self.fw, self.fb]

    def train_step(self, x, y, lr):
        c, col = X.conv2d_fwd(x, self.w, self.b, (1, 1), 0)
        r = X.relu_fwd(c)
        p, am = X.max_pool2d_fwd(r, 2, (2, 2), 0)
        feat = p.reshape(p.shape[0], -1)
        z = X.addmm_fwd(self.fb, feat, self.fw)
        loss, probs = X.cross_entropy_fwd(z, y)
        dfeat, dfw, dfb = X.addmm_bwd(X.cross_entropy_bwd(probs, y), feat, self.fw, self.fb.shape)
        dr = X.max_pool2d_bwd(dfeat.reshape(p.shape), am, r.shape, 2, (2, 2), 0)
        _, dw, db = X.conv2d_bwd(X.relu_bwd(dr, c), x.shape, self.w, col, (1, 1), 0, need_dx=False)
        X.sgd_step(self.params, [dw, db.reshape(1, -1), dfw, dfb.reshape(1, -1)], lr)
        return float(loss[0])


class ReadmeMLP:
    def __init__(self, n_in=784, hidden=500, classes=10):
        self.w1 = np.random.normal(loc=0.0, scale=0.1, size=(n_in, hidden))
        self.b1 = np.zeros((1, hidden))
        self.w2 = np.random.normal(loc=0.0, scale=0.1, size=(hidden, classes))
        self.b2 = np.zeros((1, classes))
        self.params = [self.w1, self.b1, self.w2, self.b2]

    def train_step(self, x, y, lr):
        h = X.addmm_fwd(self.b1, x, self.w1)
        a = X.relu_fwd(h)
        z = X.addmm_fwd(self.b2, a, self.w2)
        loss, probs = X.cross_entropy_fwd(z, y)
        da, dw2, db2 = X.addmm_bwd(X.cross_entropy_bwd(probs, y), a, self.w2, self.b2.shape)
        _, dw1, db1 = X.addmm_bwd(X.relu_bwd(da, h), x, self.w1, self.b1.shape)
        X.sgd_step(self.params, [dw1, db1.reshape(1, -1), dw2, db2.reshape(1, -1)], lr)
        return float(loss[0])
