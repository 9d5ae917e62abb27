"""CPU: the data-parallel path at world_size 2 over gloo (host logic + fake backend kernels):
bucket planning, overlapped bucket all-reduce hooks, grad_scale = 1/world in the fused update,
checked against the per-shard oracle average."""
import socket

import pytest
import torch.multiprocessing as mp

import dp_worker
from xshinnosuke_b200.parallel import plan_buckets


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_plan_buckets_walks_backwards_and_covers_everything():
    offsets, sizes = [0, 64, 128, 1152, 1216], [10, 64, 1000, 64, 30]
    b = plan_buckets(offsets, sizes, 1280, 512)
    assert [x[:2] for x in b] == [(128, 1280), (0, 128)]          # launch order: tail of the buffer first
    assert [x[2] for x in b] == [[4, 3, 2], [1, 0]]
    assert sorted(i for x in b for i in x[2]) == [0, 1, 2, 3, 4]
    assert plan_buckets([0], [5], 64, 10 ** 9) == [(0, 64, [0])]
    assert plan_buckets([], [], 0, 10) == []
    # small tail bucket: the first-created parameters (their gradients come last) are reduced on their own
    t = plan_buckets(offsets, sizes, 1280, 512, tail_floats=128)
    assert [x[:2] for x in t] == [(128, 1280), (0, 128)] and t[-1][2] == [1, 0]
    t = plan_buckets(offsets, sizes, 1280, 10 ** 9, tail_floats=100)
    assert [x[:2] for x in t] == [(64, 1280), (0, 64)] and t[-1][2] == [0]
    assert sorted(i for x in t for i in x[2]) == [0, 1, 2, 3, 4]


@pytest.mark.parametrize("optimizer", ["sgd", "adam"])
def test_two_ranks_match_per_shard_oracle(optimizer):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=dp_worker.run, args=(r, 2, port, True, 3, optimizer, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in results:
        assert r["max_err"] <= 5e-6, r
        assert r["n_buckets"] >= 3                       # tiny bucket size -> several buckets
        assert r["n_allreduce"] == 3 * r["n_buckets"]    # every bucket reduced exactly once per step


def _accum_worker(rank, world, port, q):
    """Two backward passes per step: refused without no_sync(), correct with it (gradients of both micro-batches summed
    locally, reduced once)."""
    import os
    import numpy as np
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import fake_backend
    fake_backend.install()
    import xshinnosuke_b200 as xs
    from xshinnosuke_b200 import nn
    from xshinnosuke_b200.layers import Dense
    from xshinnosuke_b200.parallel import DataParallel
    rs = np.random.RandomState(rank)
    xa, xb = rs.rand(4, 6).astype(np.float32), rs.rand(4, 6).astype(np.float32)
    ya, yb = rs.randint(0, 3, (4,)), rs.randint(0, 3, (4,))
    np.random.seed(0)
    net = nn.Sequential(Dense(5), Dense(3))
    opt = xs.optim.SGD(net.parameters(), lr=0.1)
    dp = DataParallel(net, opt, bucket_mb=1e-5)
    crit = nn.CrossEntropyLoss()
    opt.zero_grad()
    crit(net(xs.tensor(xa)), xs.tensor(ya)).backward()          # warm-up step: builds the flat buffers / bucket plan
    opt.step()
    w0 = [p.eval.astype(np.float64) for p in net.parameters()]
    # (1) plain second backward in the same step: refused
    opt.zero_grad()
    crit(net(xs.tensor(xa)), xs.tensor(ya)).backward()
    refused = False
    try:
        crit(net(xs.tensor(xb)), xs.tensor(yb)).backward()
    except RuntimeError as e:
        refused = "no_sync" in str(e)
    opt.step()                                                  # finishes the step so that the ranks stay in lockstep
    for p, v in zip(net.parameters(), w0):
        p.eval = v
    # (2) with no_sync(): local sum of both micro-batches, one reduction
    opt.zero_grad()
    with dp.no_sync():
        crit(net(xs.tensor(xa)), xs.tensor(ya)).backward()
    crit(net(xs.tensor(xb)), xs.tensor(yb)).backward()
    g_local = [p.grad.eval.astype(np.float64) for p in net.parameters()]   # (already reduced: sums over ranks)
    opt.step()
    w1 = [p.eval.astype(np.float64) for p in net.parameters()]
    q.put(dict(rank=rank, refused=refused, w0=w0, w1=w1, g=g_local))


def test_gradient_accumulation_needs_no_sync():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_accum_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda r: r["rank"])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import numpy as np
    assert all(r["refused"] for r in res)
    for a, b in zip(res[0]["w1"], res[1]["w1"]):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-7)          # replicas stay identical
    for w0, w1, g in zip(res[0]["w0"], res[0]["w1"], res[0]["g"]):
        np.testing.assert_allclose(w1, w0 - 0.1 * g / 2, rtol=0, atol=2e-6)   # update = lr * (sum over ranks) / world
