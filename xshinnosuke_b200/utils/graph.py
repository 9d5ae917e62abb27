"""Whole-step CUDA-graph capture.

At small inputs (ResNet-18 at 56x56, batch 32) a training step is ~215 kernel launches of a few microseconds
each and the Python that issues them costs more than the GPU work. `GraphedStep` records one full step --
zero_grad, forward, loss, backward, fused optimizer update, with every launch going through the C ABI on the
capturing stream -- into a CUDA graph and replays it per batch: no tracing compiler, the captured kernels,
memsets and TMA tensor maps are exactly the eager ones.

Contract: the step function must be shape-static and must not synchronise (`.item()`, `.eval`) inside; feed new
batches by assigning into the SAME input tensors (`x.eval = batch`) before calling the graphed step; read the
loss of the returned tensor afterwards. Adam's step counter lives on the device, so bias correction advances
across replays; BatchNorm running statistics and parameters are updated in place by the captured kernels.
"""
from ..core import state as GLOBAL
from ..core.device import rt, torch


class GraphedStep:
    def __init__(self, step_fn, warmup=3):
        rt.ensure()
        t = torch()
        self._fn = step_fn
        side = t.cuda.Stream(device=rt.device)
        side.wait_stream(t.cuda.current_stream())
        with t.cuda.stream(side):          # warm-up off the default stream: builds flat buffers, scratch, attributes
            for _ in range(max(warmup, 1)):
                step_fn()
        t.cuda.current_stream().wait_stream(side)
        t.cuda.synchronize()
        GLOBAL.INPUTS = GLOBAL.OUTPUTS = GLOBAL.GRAPH = None
        self.graph = t.cuda.CUDAGraph()
        l0 = rt.launches
        with t.cuda.graph(self.graph):
            self.result = step_fn()
        self.launches_per_replay = rt.launches - l0
        rt.launches = l0
        self.replays = 0

    def __call__(self):
        self.graph.replay()
        self.replays += 1
        rt.launches += self.launches_per_replay
        return self.result
