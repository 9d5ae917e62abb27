"""Batch-sharded data parallelism for the xs training loop: one process per GPU, NCCL over NVLink.

The reference has no distributed code at all; this is the B200-side addition the
north_star asks for. Semantics: every rank is "the reference run on its own shard" (per-replica
BatchNorm statistics, loss = mean over the local batch); parameter gradients are SUMMED across ranks
and the fused optimizer kernel applies grad_scale = 1/world (so the update uses the average).

Gradient accumulation (several backward() calls before one step()): a bucket is all-reduced IN PLACE once per step, so a
backward that starts after buckets of this step were already reduced would add local gradients on top of a cross-rank
sum. That is refused with an error; wrap every backward but the last in `with dp.no_sync():` -- inside it nothing is
reduced, and the last backward (or step()) reduces the accumulated local gradients once.

Mechanics: parameter gradients already live in one flat fp32 buffer laid out in parameter-creation
order (optim/optimizer.py). It is cut into buckets from the END (fc -> layer4 -> ... -> conv1: the order
in which backward produces them). `core.autograd.backward` calls `after_node` after every closure;
as soon as every parameter of the next bucket has received its gradient the bucket's all-reduce is
issued on a dedicated communication stream, overlapping the remaining backward kernels. `step()`
waits for the communication stream and runs the single fused update kernel.
"""
import os

from .core import state as GLOBAL
from .core.device import rt, torch


def init_process_group(backend=None):
    """Idempotent torch.distributed init from RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT."""
    import torch.distributed as dist
    if not dist.is_available():
        raise RuntimeError("torch.distributed is not available")
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch().cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        kw = {}
        if backend == "nccl":
            rt.ensure()
            kw["device_id"] = rt.device
        dist.init_process_group(backend=backend, **kw)
    return dist


def plan_buckets(offsets, sizes, total, bucket_floats, tail_floats=0):
    """Cut [0, total) into contiguous ranges, walking parameters from the last one backwards.
    Returns a list of (start, stop, [param indices]) in launch order (last parameters first).

    tail_floats > 0: the bucket launched LAST (the first-created parameters: the stem, whose gradients appear at the very
    end of backward) is kept that small -- its all-reduce has nothing left to overlap with, so it should be latency-
    rather than bandwidth-sized (measured on 2 x B200: 43 us exposed for an 8 MB tail bucket)."""
    head = 0            # parameters [0, head) form the small tail bucket
    tail_floats = min(tail_floats, total // 8)      # (a tiny model is not mostly "tail")
    if tail_floats > 0 and len(offsets) > 1:
        while head + 1 < len(offsets) and offsets[head + 1] <= tail_floats:
            head += 1
        head = max(head, 1)
    buckets, cur, stop = [], [], total
    for i in range(len(offsets) - 1, head - 1, -1):
        cur.append(i)
        start = offsets[i]
        if stop - start >= bucket_floats or i == head:
            buckets.append((start, stop, cur))
            cur, stop = [], start
    if head > 0:
        buckets.append((0, offsets[head], list(range(head - 1, -1, -1))))
    return buckets


class DataParallel:
    """Wrap a (model, optimizer) pair:

        opt = xs.optim.SGD(net.parameters(), lr=0.1)
        dp = DataParallel(net, opt)           # after torchrun / init_process_group
        ... usual loop: opt.zero_grad(); loss = crit(net(x_shard), y_shard); loss.backward(); opt.step()
    """

    def __init__(self, model, optimizer, bucket_mb=8.0, overlap=True, broadcast_params=True, tail_kb=64.0):
        self.dist = init_process_group()
        self.world = self.dist.get_world_size()
        self.rank = self.dist.get_rank()
        self.model, self.opt = model, optimizer
        self.bucket_floats = int(bucket_mb * (1 << 20) / 4)
        self.tail_floats = int(tail_kb * 1024 / 4)
        self.overlap = overlap
        self.broadcast_params = broadcast_params
        self._flat = None
        self._buckets = []
        self._next = 0
        self._comm = None
        self._sync = True          # False inside no_sync(): backward passes only accumulate locally
        self.n_allreduce = 0
        optimizer.grad_scale = 1.0 / self.world
        optimizer.before_step = self._before_step
        GLOBAL.AFTER_NODE_BACKWARD = self._after_node
        GLOBAL.BEFORE_BACKWARD = self._before_backward

    # ---- plumbing
    def _cuda(self):
        return rt.device is not None and rt.device.type == "cuda"

    def _comm_stream(self):
        if self._comm is None and self._cuda():
            self._comm = torch().cuda.Stream(device=rt.device)
        return self._comm

    def _sync_plan(self):
        flat = self.opt.flat
        if flat is self._flat:
            return
        first = self._flat is None
        self._flat = flat
        sizes = [p.numel() for p in flat.params]
        self._buckets = plan_buckets(flat.offsets, sizes, flat.total, self.bucket_floats, self.tail_floats)
        self._next = 0
        if self.broadcast_params and flat.total:
            # replicas start from rank 0's weights (identical seeds make this a no-op, but do not rely on it)
            self.dist.broadcast(flat.flat_p, src=0)
        del first

    def _launch(self, b):
        start, stop, _ = self._buckets[b]
        if stop <= start:
            return
        view = self._flat.flat_g[start:stop]
        if self._cuda():
            comm = self._comm_stream()
            comm.wait_stream(torch().cuda.current_stream())
            with torch().cuda.stream(comm):
                self.dist.all_reduce(view, op=self.dist.ReduceOp.SUM)
        else:
            self.dist.all_reduce(view, op=self.dist.ReduceOp.SUM)
        self.n_allreduce += 1

    def _bucket_ready(self, b):
        params = self._flat.params
        for i in self._buckets[b][2]:
            p = params[i]
            if p.pending_uses > 0 or p._grad._arr.pending_zero:
                return False
        return True

    # ---- gradient accumulation
    def no_sync(self):
        """Context manager: backward passes inside it do not all-reduce (use for all but the last micro-batch)."""
        dp = self

        class _NoSync:
            def __enter__(self):
                self.prev, dp._sync = dp._sync, False

            def __exit__(self, *exc):
                dp._sync = self.prev
                return False
        return _NoSync()

    # ---- hooks
    def _before_backward(self):
        if self.world > 1 and self._next > 0:
            raise RuntimeError(
                "DataParallel: backward() started after gradient buckets of this step were already all-reduced "
                f"({self._next} of {len(self._buckets)}); with several backward passes per step() wrap all but the last in "
                "`with dp.no_sync():`")

    def _after_node(self, node):
        if not self.overlap or not self._sync or self.world == 1 or self._flat is None or self.opt.flat is not self._flat:
            return
        while self._next < len(self._buckets) and self._bucket_ready(self._next):
            self._launch(self._next)
            self._next += 1

    def _before_step(self, opt):
        self._sync_plan()
        if self.world > 1:
            while self._next < len(self._buckets):
                self._launch(self._next)
                self._next += 1
            if self._cuda():
                torch().cuda.current_stream().wait_stream(self._comm_stream())
        self._next = 0
        # a forward that never saw its backward (validation without no_grad) must not hold buckets back for ever
        for p in self._flat.params:
            p.pending_uses = 0

    def close(self):
        if GLOBAL.AFTER_NODE_BACKWARD == self._after_node:
            GLOBAL.AFTER_NODE_BACKWARD = None
        if GLOBAL.BEFORE_BACKWARD == self._before_backward:
            GLOBAL.BEFORE_BACKWARD = None
        self.opt.before_step = None
        self.opt.grad_scale = 1.0


def shard(array, rank, world):
    """Contiguous batch shard `rank` of `world`."""
    n = array.shape[0]
    per = n // world
    return array[rank * per:(rank + 1) * per]
