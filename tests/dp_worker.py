"""Worker for the data-parallel tests: runs on every rank (gloo + fake backend on CPU, or NCCL on GPUs).

Net (bias-free convs): Conv(8,3x3,p1)-BN-ReLU(in place)-MaxPool(2)-Conv(16,3x3,p1)-BN-ReLU-AvgPool(2)-Flatten-Dense(10).
DP parity oracle: run the NumPy oracle separately on each rank's shard from identical
weights, average the parameter gradients, apply one update; every rank must hold those parameters."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def oracle_shard_grads(x, y, P):
    from oracle import xs_numpy as X
    w1, g1, be1, w2, g2, be2, wf, bf = P
    b1 = b2 = None   # bias-free convs: a bias in front of a BN has zero gradient, which Adam turns into noise
    c1, col1 = X.conv2d_fwd(x, w1, b1, (1, 1), 1)
    n1, xh1, sv1, _, _ = X.batch_norm_fwd(c1, g1, be1, np.zeros(8), np.ones(8), 1, 1e-6, 0.99)
    r1 = X.relu_fwd(n1)
    p1, am = X.max_pool2d_fwd(r1, 2, None, 0)
    c2, col2 = X.conv2d_fwd(p1, w2, b2, (1, 1), 1)
    n2, xh2, sv2, _, _ = X.batch_norm_fwd(c2, g2, be2, np.zeros(16), np.ones(16), 1, 1e-6, 0.99)
    r2 = X.relu_fwd(n2)
    a2 = X.avg_pool2d_fwd(r2, 2)
    f = a2.reshape(len(x), -1)
    z = X.addmm_fwd(bf, f, wf)
    loss, pr = X.cross_entropy_fwd(z, y)
    dz = X.cross_entropy_bwd(pr, y)
    df, dwf, dbf = X.addmm_bwd(dz, f, wf, bf.shape)
    dr2 = X.avg_pool2d_bwd(df.reshape(a2.shape), r2.shape, 2)
    dc2, dg2, dbe2 = X.batch_norm_bwd(X.relu_bwd(dr2, n2), xh2, sv2, g2, 1)
    dp1, dw2, db2 = X.conv2d_bwd(dc2, p1.shape, w2, col2, (1, 1), 1)
    dr1 = X.max_pool2d_bwd(dp1, am, r1.shape, 2, None, 0)
    dc1, dg1, dbe1 = X.batch_norm_bwd(X.relu_bwd(dr1, n1), xh1, sv1, g1, 1)
    _, dw1, db1 = X.conv2d_bwd(dc1, x.shape, w1, col1, (1, 1), 1, need_dx=False)
    return float(loss[0]), [dw1, dg1, dbe1, dw2, dg2, dbe2, dwf, dbf.reshape(1, -1)]


def run(rank, world, port, use_fake, steps, optimizer, out_queue=None, bucket_mb=0.001, use_graph=False):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    if use_fake:
        import fake_backend
        fake_backend.install()
    import xshinnosuke_b200 as xs
    from oracle import xs_numpy as X
    from xshinnosuke_b200 import nn
    from xshinnosuke_b200.layers import (AvgPooling2D, BatchNormalization, Conv2D, Dense, Flatten, MaxPooling2D, ReLU)
    from xshinnosuke_b200.parallel import DataParallel, shard

    rs = np.random.RandomState(0)
    X_all = rs.rand(4 * world, 3, 8, 8).astype(np.float32)
    Y_all = rs.randint(0, 10, (4 * world,))
    np.random.seed(5)
    net = nn.Sequential(Conv2D(8, 3, padding=1, use_bias=False), BatchNormalization(), ReLU(True), MaxPooling2D(2),
                        Conv2D(16, 3, padding=1, use_bias=False), BatchNormalization(), ReLU(True), AvgPooling2D(2), Flatten(),
                        Dense(10))
    lr = 0.1 if optimizer == "sgd" else 0.01
    opt = xs.optim.SGD(net.parameters(), lr=lr) if optimizer == "sgd" else xs.optim.Adam(net.parameters(), lr=lr)
    dp = DataParallel(net, opt, bucket_mb=bucket_mb)
    xs_x, xs_y = shard(X_all, rank, world), shard(Y_all, rank, world)
    ref_params, adam_state, errs, losses = None, {}, [], []
    tx, ty = xs.tensor(xs_x), xs.tensor(xs_y)
    graphed = None
    for step in range(steps):
        if use_graph and step == 1:
            # steps 1.. are replays of ONE captured step (kernels + bucket all-reduces); GraphedStep's own
            # warm-up step is step 1, the capture itself executes nothing
            def one_step():
                opt.zero_grad()
                l_ = nn.CrossEntropyLoss()(net(tx), ty)
                l_.backward()
                opt.step()
                return l_
            graphed = xs.GraphedStep(one_step, warmup=1)
            loss = graphed.result
            losses.append(float("nan"))
        elif graphed is not None:
            losses.append(graphed().item())
        else:
            opt.zero_grad()
            loss = nn.CrossEntropyLoss()(net(tx), ty)
            if ref_params is None:   # parameters exist after the first forward
                ref_params = [p.eval.astype(np.float64) for p in net.parameters()]
            losses.append(loss.item())
            loss.backward()
            opt.step()
        # oracle: every shard separately, averaged gradients, one update
        grads, ref_losses = None, []
        for r in range(world):
            l, g = oracle_shard_grads(shard(X_all, r, world).astype(np.float64), shard(Y_all, r, world), ref_params)
            ref_losses.append(l)
            grads = g if grads is None else [a + b for a, b in zip(grads, g)]
        grads = [g / world for g in grads]
        if optimizer == "sgd":
            X.sgd_step(ref_params, grads, lr)
        else:
            X.adam_step(ref_params, grads, adam_state, lr=lr)
        # conv biases in front of a BN have a (mathematically) zero gradient and stay ~0: floor the denominator
        errs.append(max(float(np.max(np.abs(p.eval - r)) / max(np.max(np.abs(r)), 1e-3))
                        for p, r in zip(net.parameters(), ref_params)))
        if losses[-1] == losses[-1]:   # (the capture step has no readable loss)
            errs.append(abs(losses[-1] - ref_losses[rank]) / ref_losses[rank])
    result = dict(rank=rank, max_err=max(errs), n_allreduce=dp.n_allreduce, n_buckets=len(dp._buckets),
                  losses=losses)
    dp.dist.barrier()
    if out_queue is not None:
        out_queue.put(result)
    return result


if __name__ == "__main__":   # torchrun entry (GPU): python tests/dp_worker.py [sgd|adam]
    opt_name = sys.argv[1] if len(sys.argv) > 1 else "sgd"
    r = run(int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("MASTER_PORT", 29511)),
            False, 4, opt_name, use_graph="graph" in sys.argv[2:])
    print("DP_RESULT", r, flush=True)
    assert r["max_err"] <= 2e-5, r
