"""Whole-step CUDA-graph capture.

At small inputs (ResNet-18 at 56x56, batch 32) a training step is ~215 kernel launches of a few microseconds
each and the Python that issues them costs more than the GPU work. `GraphedStep` records one full step --
zero_grad, forward, loss, backward, fused optimizer update, with every launch going through the C ABI on the
capturing stream -- into a CUDA graph and replays it per batch: no tracing compiler, the captured kernels,
memsets and TMA tensor maps are exactly the eager ones.

Feeding batches: pass the captured input tensors as `inputs=(x, y)`; `g(x_host, y_host)` then
uploads the host batch into a device staging slot, lays it out into the captured tensors (NCHW -> channels-last on the
device) and replays; `g.prefetch(x_next, y_next)` starts the upload of the NEXT batch on a side stream so that it
overlaps the running replay (two staging slots, event-ordered; pinned host arrays copy asynchronously).

Contract: the step function must be shape-static and must not synchronise (`.item()`, `.eval`) inside; feed new
batches by assigning into the SAME input tensors (`x.eval = batch`) before calling the graphed step; read the
loss of the returned tensor afterwards. Adam's step counter lives on the device, so bias correction advances
across replays; BatchNorm running statistics and parameters are updated in place by the captured kernels.
"""
import numpy as np

from .._lib import lib
from ..core import state as GLOBAL
from ..core.device import rt, torch


class _Slot:
    """One device staging copy of a host batch: raw (logical NCHW / row-major) buffers, filled on the side stream."""

    def __init__(self, inputs):
        t = torch()
        self.bufs = [t.empty(max(x.arr.size, 1), dtype=t.float32 if x.arr.dtype == "float32" else t.int64, device=rt.device)
                     for x in inputs]
        self.uploaded = t.cuda.Event()     # recorded on the side stream after the H2D copies
        self.consumed = None               # recorded on the main stream after the commit that read this slot
        self.key = None                    # identities of the host arrays it holds
        self.keep = None                   # host tensors kept alive until the copy ran


class GraphedStep:
    def __init__(self, step_fn, warmup=3, inputs=None):
        rt.ensure()
        t = torch()
        self._fn = step_fn
        self._inputs = list(inputs) if inputs is not None else None
        self._slots = None
        self._next_slot = 0
        self._side_in = None
        if warmup > 0:
            side = t.cuda.Stream(device=rt.device)
            side.wait_stream(t.cuda.current_stream())
            with t.cuda.stream(side):      # warm-up off the default stream: builds flat buffers, scratch, attributes
                for _ in range(warmup):
                    step_fn()
            t.cuda.current_stream().wait_stream(side)
        # warmup == 0: the caller has already run this very step eagerly (Model.fit captures after its first batches,
        # every one of which is a real update) -- capture only
        t.cuda.synchronize()
        GLOBAL.INPUTS = GLOBAL.OUTPUTS = GLOBAL.GRAPH = None
        self.graph = t.cuda.CUDAGraph()
        l0 = rt.launches
        with t.cuda.graph(self.graph):
            self.result = step_fn()
        self.launches_per_replay = rt.launches - l0
        rt.launches = l0
        self.replays = 0

    # ---- input feeding
    def _ensure_slots(self):
        if self._inputs is None:
            raise ValueError("GraphedStep(..., inputs=(x, y)) is needed to feed host batches")
        if self._slots is None:
            self._slots = [_Slot(self._inputs), _Slot(self._inputs)]
            self._side_in = torch().cuda.Stream(device=rt.device)

    def _host_tensor(self, a, x):
        t = torch()
        want = np.float32 if x.arr.dtype == "float32" else np.int64
        a = np.ascontiguousarray(np.asarray(a), dtype=want)   # no copy when the caller's array already fits
        if a.size != x.arr.size:
            raise ValueError(f"batch has {a.size} elements, the captured input holds {x.arr.size}")
        if not a.flags.writeable:
            a = a.copy()
        return t.from_numpy(a.reshape(-1))

    def prefetch(self, *host):
        """Start uploading a host batch (async when the arrays are pinned) into the free staging slot."""
        self._ensure_slots()
        t = torch()
        slot = self._slots[self._next_slot]
        self._next_slot ^= 1
        if slot.consumed is not None:
            self._side_in.wait_event(slot.consumed)          # the commit that read this slot has finished
        srcs = [self._host_tensor(a, x) for a, x in zip(host, self._inputs)]
        with t.cuda.stream(self._side_in):
            for dst, src in zip(slot.bufs, srcs):
                dst.copy_(src, non_blocking=True)
            slot.uploaded.record(self._side_in)
        slot.key = tuple(id(a) for a in host)
        slot.keep = (srcs, host)
        return slot

    def _commit(self, slot):
        t = torch()
        main = t.cuda.current_stream()
        main.wait_event(slot.uploaded)
        for x, buf in zip(self._inputs, slot.bufs):
            arr = x.arr
            n = max(arr.size, 1)
            if arr.cl and arr.shape[1] > 1 and arr.shape[2] * arr.shape[3] > 1:
                nn_, c, h, w = arr.shape
                lib.xsb_nchw_to_nhwc(buf.data_ptr(), arr.wptr, nn_, c, h, w, 0, rt.stream())
                rt.launches += 1
            else:
                arr.wptr  # noqa: B018
                arr.buf[:n].copy_(buf[:n])
        if slot.consumed is None:
            slot.consumed = t.cuda.Event()
        slot.consumed.record(main)
        slot.key = None

    def __call__(self, *host):
        if host:
            self._ensure_slots()
            key = tuple(id(a) for a in host)
            slot = next((s for s in self._slots if s.key == key), None)
            if slot is None:
                slot = self.prefetch(*host)
            self._commit(slot)
        self.graph.replay()
        self.replays += 1
        rt.launches += self.launches_per_replay
        return self.result
