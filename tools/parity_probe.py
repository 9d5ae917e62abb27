"""Whole-network parity probe: ResNet-18 step 0, GPU (given math mode) vs the float64 oracle, error of the logits, the loss
and EVERY parameter gradient in creation order -- shows how the per-op error of a math mode grows through the backward pass."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xshinnosuke_b200 as xs  # noqa: E402
from oracle import xs_numpy as X  # noqa: E402
from oracle.resnet18 import ResNet18 as OracleNet  # noqa: E402
from xshinnosuke_b200 import nn  # noqa: E402
from xshinnosuke_b200.recipes import ResNet18  # noqa: E402

n, hw = int(sys.argv[1]), int(sys.argv[2])
modes = sys.argv[3].split(",")
rs = np.random.RandomState(0)
x, y = rs.rand(n, 3, hw, hw).astype(np.float32), rs.randint(0, 100, (n,))


def oracle(tf32):
    X.TF32_OPERANDS = tf32
    try:
        np.random.seed(0)
        onet = OracleNet(num_classes=100)
        lg = onet.forward(x.astype(np.float64))
        ls, probs = X.cross_entropy_fwd(lg, y)
        return lg, ls, onet.backward(X.cross_entropy_bwd(probs, y))
    finally:
        X.TF32_OPERANDS = False


logits, loss, grads = oracle(False)
lt, losst, gt = oracle(True)
ge = [float(X.rel_err(a, b)) for a, b in zip(gt, grads) if a.ndim == 4]
print(f"tf32-operand oracle vs exact oracle: logits {X.rel_err(lt, logits):.2e} loss {abs(losst[0] - loss[0]) / loss[0]:.2e} "
      "conv dW errs first->last: " + " ".join(f"{e:.1e}" for e in ge), flush=True)
for mode in modes:
    xs.set_math_mode(mode)
    np.random.seed(0)
    net = ResNet18(num_classes=100)
    opt = xs.optim.SGD(net.parameters(), lr=0.1)
    opt.zero_grad()
    out = net(xs.tensor(x))
    gl = out.eval
    l = nn.CrossEntropyLoss()(out, xs.tensor(y))
    lv = l.item()
    l.backward()
    ps = sorted(net.parameters(), key=lambda p: p.ticket)
    errs = [float(X.rel_err(p.grad.eval.reshape(g.shape), g)) for p, g in zip(ps, grads)]
    w_errs = [e for p, e in zip(ps, errs) if len(p.shape) == 4]
    print(f"{mode} n={n} hw={hw}: logits {X.rel_err(gl, logits):.2e} loss {abs(lv - loss[0]) / loss[0]:.2e} "
          f"conv dW errs first->last: " + " ".join(f"{e:.1e}" for e in w_errs) + f" | fc dW {errs[-2]:.1e}", flush=True)
    if mode == "tf32":
        e2 = [float(X.rel_err(p.grad.eval.reshape(g.shape), g)) for p, g in zip(ps, gt)]
        print(f"   vs the tf32-operand oracle: logits {X.rel_err(gl, lt):.2e} loss {abs(lv - losst[0]) / losst[0]:.2e} conv dW errs: "
              + " ".join(f"{e:.1e}" for p, e in zip(ps, e2) if len(p.shape) == 4) + f" | fc dW {e2[-2]:.1e}", flush=True)
