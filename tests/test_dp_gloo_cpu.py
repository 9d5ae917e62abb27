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
