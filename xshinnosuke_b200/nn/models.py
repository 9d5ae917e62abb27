"""Model containers (mirror of xs/nn/models.py:14-363): Sequential, functional Model, dynamic Module.

`fit` / `evaluate` / `predict` / `save` / `load` keep the reference's behaviour; host batches are
uploaded once per step (the reference's `batch_x.cuda_()` copy, models.py:95-97). Parameters are
kept in a ParameterSet that iterates in creation order (the reference's plain `set` has hash order,
SURVEY.md Q10) so the optimizer's flat buffers are deterministic across ranks.
"""
import pickle
import time

import numpy as np

from ..core import autograd
from ..core import state as GLOBAL
from ..core.tensor import Parameter, Tensor
from ..optim import get_optimizer
from ..utils.data import DataLoader, DataSet
from ..utils.toolkit import ProgressBar, SummaryProfile
from . import functional as F
from . import objectives


class ParameterSet:
    """Set semantics (`add`, `in`, `len`, iteration) with creation-order iteration."""

    def __init__(self, items=()):
        self._d = {}
        for p in items:
            self.add(p)

    def add(self, p):
        self._d[id(p)] = p

    def __contains__(self, p):
        return id(p) in self._d

    def __iter__(self):
        return iter(sorted(self._d.values(), key=lambda p: getattr(p, "ticket", 0)))

    def __len__(self):
        return len(self._d)


class _Base:
    def __init__(self):
        self._parameters = ParameterSet()

    def train(self):
        """The reference writes IS_TRAINING, which no op ever reads (models.py:18-22, SURVEY Q5: its BatchNorm stays on batch
        statistics in `evaluate`). SURVEY 8f-4 "fixing the train()/eval() flag": here both flags move, so eval() puts
        BatchNorm on its running statistics and stops in-place ReLUs from recording masks."""
        GLOBAL.IS_TRAINING = GLOBAL.TRAINING = True

    def eval(self):
        GLOBAL.IS_TRAINING = GLOBAL.TRAINING = False

    def parameters(self, parameters=None):
        if parameters is None:
            return self._parameters
        self._parameters = ParameterSet(parameters)

    def to(self, dst):
        return self   # everything already lives on the device


class _Model(_Base):
    def __init__(self):
        super().__init__()
        self.loss = None
        self.optimizer = None
        self._graph = None

    def compile(self, optimizer, loss):
        pass

    def init_graph_out_tensor(self, x, y):
        raise NotImplementedError

    # ------------------------------------------------------------------ training
    def _train_batch(self, bx, by):
        """One update on the static graph: the body of the reference's batch loop (xs/nn/models.py:100-106)."""
        self.train()
        self.optimizer.zero_grad()
        pred = self.forward(bx)
        loss = self.loss.forward(pred, by)
        self.backward(loss)
        self.optimizer.step()
        return pred, loss

    def fit(self, x, y, batch_size=None, epochs=1, verbose=1, shuffle=True, validation_data=None,
            validation_split=0.0, initial_epoch=0, cuda_graph=False):
        """xs/nn/models.py:61-129. `cuda_graph=True` (no reference counterpart): the static graph's
        train step -- zero_grad, forward into the preallocated `out=` buffers, loss, backward, optimizer step -- is
        captured into ONE CUDA graph after the first two full batches (which run eagerly: every batch is exactly one
        update either way) and replayed for every further full-size batch; a partial last batch runs eagerly."""
        from ..core.device import rt
        x, y = np.asarray(x), np.asarray(y)
        columns = ["train_loss", "train_acc"]
        if validation_data is None and 0.0 < validation_split < 1.0:
            n_val = int(x.shape[0] * validation_split)
            validation_data, x, y = (x[-n_val:], y[-n_val:]), x[:-n_val], y[:-n_val]
            columns += ["validation_loss", "validation_acc"]
        loader = DataLoader(DataSet(x, y), batch_size, shuffle)
        replay = _StaticStepReplay(self) if (cuda_graph and validation_data is None and batch_size is not None and
                                             rt.device is not None and rt.device.type == "cuda") else None
        val = (0.0, 0.0)
        with SummaryProfile(*columns, verbose=verbose) as profile:
            for epoch in range(initial_epoch, epochs):
                started = time.time()
                if verbose != 0:
                    print("\033[1;31m Epoch[%d/%d]\033[0m" % (epoch + 1, epochs))
                bar = ProgressBar(max_iter=len(x), verbose=verbose)
                last = len(loader) - 1
                for idx, (bx, by) in enumerate(loader):
                    if replay is not None and bx.shape[0] == batch_size:
                        pred, loss, by = replay.step(bx, by, first_of_epoch=idx == 0)
                    else:
                        # output buffers are (re)sized on the first batch and shrunk to a prefix view on a partial last one
                        if idx in (0, last) or replay is not None:
                            self.init_graph_out_tensor(bx, by)
                        pred, loss = self._train_batch(bx, by)
                    acc, loss_value = self.loss.calc_acc(pred, by), loss.item()
                    profile.step("train_acc", acc)
                    profile.step("train_loss", loss_value)
                    if validation_data is not None:
                        val = self.evaluate(validation_data[0], validation_data[1], batch_size=batch_size)
                        profile.step("validation_loss", val[1])
                        profile.step("validation_acc", val[0])
                    bar.update(bx.shape[0])
                    bar.console(verbose, time.time() - started, loss_value, acc, val[1], val[0])
                if verbose != 0:
                    print()
        return profile

    def __call__(self, x, *args, **kwargs):
        return self.forward(x)

    def forward(self, x, *args, **kwargs):
        raise NotImplementedError

    def backward(self, loss):
        """Static-graph backward (models.py:138-144): the loss closure, then every layer in reverse."""
        if loss.grad is None or loss.grad.arr.pending_zero:
            loss.zero_grad()
            loss.grad.fill_(1.0)
        if GLOBAL.BEFORE_BACKWARD is not None:
            GLOBAL.BEFORE_BACKWARD()
        loss.grad_fn(loss)
        hook = GLOBAL.AFTER_NODE_BACKWARD
        for layer in reversed(self._graph):
            layer.backward()
            if hook is not None:
                hook(layer)
        GLOBAL.INPUTS = None

    # ------------------------------------------------------------------ inference (xs/nn/models.py:146-186)
    def evaluate(self, x, y, batch_size=None):
        """-> (accuracy, loss) with BatchNorm on its running statistics; batched: the mean over the batches."""
        self.eval()
        x, y = np.asarray(x), np.asarray(y)
        if batch_size is None:
            tx, ty = Tensor(x), Tensor(y)
            self.init_graph_out_tensor(tx, ty)
            return self.loss.metric(self.forward(tx), ty)
        assert type(batch_size) is int
        loader = DataLoader(DataSet(x, y), batch_size)
        last, scores = len(loader) - 1, []
        for idx, (bx, by) in enumerate(loader):
            if idx in (0, last):
                self.init_graph_out_tensor(bx, by)
            scores.append(self.loss.metric(self.forward(bx), by))
        return float(np.mean([a for a, _ in scores])), float(np.mean([l for _, l in scores]))

    def predict(self, x, batch_size=None):
        self.eval()
        x = np.asarray(x)
        step = batch_size if batch_size is not None else max(len(x), 1)
        chunks = []
        for lo in range(0, len(x), step):
            tx = Tensor(x[lo:lo + step])
            self.init_graph_out_tensor(tx, None)
            chunks.append(self.forward(tx).eval)
        return Tensor(np.concatenate(chunks, axis=0))

    # ------------------------------------------------------------------ persistence
    CHECKPOINT_FORMAT = "xshinnosuke_b200.checkpoint"
    CHECKPOINT_VERSION = 2

    def save(self, save_path):
        """Pickle of a versioned dict of HOST arrays: parameters (creation order), BatchNorm running statistics and the
        optimizer's state (per-parameter moment buffers out of the flat device buffers, step counters, hyper-parameters).
        Intentionally NOT the reference's format -- that pickles the live `[graph, optimizer, loss]` objects
        (xs/nn/models.py:188-192), which here hold device buffers; `load` says so when handed one."""
        if not save_path.endswith(".pkl"):
            save_path += ".pkl"
        params = list(self._parameters)
        state = {"format": self.CHECKPOINT_FORMAT, "version": self.CHECKPOINT_VERSION,
                 "params": [p.eval for p in params],
                 "running": [(l.moving_mean.eval, l.moving_variance.eval) for l in self._graph
                             if getattr(l, "moving_mean", None) is not None],
                 "optimizer": self.optimizer.state_dict() if self.optimizer is not None else None}
        with open(save_path, "wb") as f:
            pickle.dump(state, f)

    def load(self, model_path):
        """Restore what `save` wrote into THIS (already compiled) model; the architecture must match."""
        if not model_path.endswith(".pkl"):
            model_path += ".pkl"
        with open(model_path, "rb") as f:
            state = pickle.load(f)
        if isinstance(state, (list, tuple)):
            raise ValueError(f"{model_path} looks like a reference-format checkpoint (a pickled [graph, optimizer, loss] list, "
                             "xs/nn/models.py:188-192); this backend stores a versioned dict of host arrays -- re-save it with "
                             "Model.save of xshinnosuke_b200")
        if not isinstance(state, dict) or state.get("format") != self.CHECKPOINT_FORMAT:
            if isinstance(state, dict) and "params" in state and "format" not in state:
                state = dict(state, format=self.CHECKPOINT_FORMAT, version=1, optimizer=None)    # round-1 files
            else:
                raise ValueError(f"{model_path} is not a {self.CHECKPOINT_FORMAT} file")
        if state["version"] > self.CHECKPOINT_VERSION:
            raise ValueError(f"checkpoint version {state['version']} is newer than this library ({self.CHECKPOINT_VERSION})")
        params = list(self._parameters)
        if len(params) != len(state["params"]):
            raise ValueError(f"checkpoint holds {len(state['params'])} parameter tensors, the model has {len(params)}")
        for i, (p, v) in enumerate(zip(params, state["params"])):
            if tuple(p.shape) != tuple(np.shape(v)):
                raise ValueError(f"parameter {i}: checkpoint shape {tuple(np.shape(v))} != model shape {tuple(p.shape)}")
        bns = [l for l in self._graph if getattr(l, "moving_mean", None) is not None]
        if len(bns) != len(state["running"]):
            raise ValueError(f"checkpoint holds {len(state['running'])} BatchNorm statistics, the model has {len(bns)}")
        for p, v in zip(params, state["params"]):
            p.eval = v
        for l, (mm, mv) in zip(bns, state["running"]):
            l.moving_mean.eval = mm
            l.moving_variance.eval = mv
        if state.get("optimizer") is not None and self.optimizer is not None:
            self.optimizer.load_state_dict(state["optimizer"])

    def __str__(self):
        if self._graph is None:
            raise ValueError("Please compile Model!")
        lines = ["*" * 75, "%-25s %-20s %-10s %-15s" % ("Layer(type)", "Output Shape", "Param", "Connected to"),
                 "#" * 75]
        counts = {}
        for layer in self._graph:
            if layer.name in (None, "None"):
                base = layer.__class__.__name__.lower()
                layer.name = base + str(counts.get(base, 0))
                counts[base] = counts.get(base, 0) + 1
        for layer in self._graph:
            conn = ", ".join(str(p.name) for p in layer.in_bounds) if layer.in_bounds else ""
            shape = (None,) + tuple(layer.shape) if layer.shape is not None else None
            lines.append("%-25s %-20s %-10s %-15s" % ("%s (%s)" % (layer.name, layer.__class__.__name__), shape,
                                                      layer.params_count(), conn))
            lines.append("-" * 75)
        total = sum(p.numel() for p in self._parameters)
        trainable = sum(p.numel() for p in self._parameters if p.requires_grad)
        lines += ["*" * 75, "Total params: %d" % total, "Trainable params: %d" % trainable,
                  "Non-trainable params: %d" % (total - trainable)]
        return "\n".join(lines)


class _StaticStepReplay:
    """Model.fit(cuda_graph=True): full-size batches are copied into two fixed input tensors on the device; the first
    two run eagerly (they build the flat optimizer buffers and size every `out=` buffer), then the step is captured once
    and replayed."""

    def __init__(self, model):
        self.model, self.x, self.y, self.graphed, self.done = model, None, None, None, 0

    def _feed(self, bx, by):
        if self.x is None:
            self.x, self.y = Tensor(bx.arr.clone()), Tensor(by.arr.clone())
            return
        for dst, src in ((self.x, bx), (self.y, by)):
            src.arr.ptr  # noqa: B018 -- materialise
            dst.arr.wptr  # noqa: B018
            dst.arr.buf[: max(dst.arr.size, 1)].copy_(src.arr.buf[: max(src.arr.size, 1)])

    def step(self, bx, by, first_of_epoch):
        self._feed(bx, by)
        if self.graphed is None or first_of_epoch:
            self.model.init_graph_out_tensor(self.x, self.y)
        if self.graphed is None and self.done >= 2:
            from ..utils.graph import GraphedStep
            self.graphed = GraphedStep(lambda: self.model._train_batch(self.x, self.y), warmup=0)   # capture only
        pred, loss = self.graphed() if self.graphed is not None else self.model._train_batch(self.x, self.y)
        self.done += 1
        return pred, loss, self.y


class Sequential(_Model):
    def __init__(self, *layers):
        super().__init__()
        self._graph = list(layers)
        self._initialized = False

    def add(self, layer):
        self._graph.append(layer)

    def init_graph_out_tensor(self, x, y):
        for layer in self._graph:
            layer.init_layer_out_tensor(x)
            x = layer.data
        if y is not None:
            self.loss.init_layer_out_tensor(x, y)

    def init_params(self, input_shape=None):
        from ..layers.base import Layer
        for layer in self._graph:
            if isinstance(layer, Layer):
                if len(layer.parameters()) == 0:
                    layer.init_params(input_shape)
                for v in layer.parameters():
                    if v is not None:
                        self._parameters.add(v)
                input_shape = layer.compute_output_shape(input_shape)
            self._initialized = True

    def compile(self, optimizer, loss, **kwargs):
        assert self._graph
        prev = None
        for layer in self._graph:
            layer.connect(prev)
            prev = layer
        self.init_params(self._graph[0]._input_shape)
        self.loss = objectives.get_objective(loss)
        self.optimizer = get_optimizer(optimizer, parameters=self._parameters, **kwargs)

    def forward(self, x, *args, **kwargs):
        if not self._initialized:
            self.init_params(x.shape[1:])
        for layer in self._graph:
            x = layer.forward(x)
        return x

    def pop(self, index):
        self._graph.pop(index)


class Model(_Model):
    """Functional API: Model(inputs=Input(...), outputs=last_layer)."""

    def __init__(self, inputs=None, outputs=None):
        super().__init__()
        self.inputs = inputs
        self.outputs = outputs

    def init_graph_out_tensor(self, x, y):
        self.inputs._data = x
        for layer in self._graph[1:]:
            layer.init_layer_out_tensor()
        if y is not None:
            self.loss.init_layer_out_tensor(self._graph[-1].data, y)

    def compile(self, optimizer, loss, **kwargs):
        assert self.inputs is not None and self.outputs is not None
        self._graph = _layer_topological_sort(self.inputs, self.outputs)
        for layer in self._graph:
            if layer is not self.inputs and len(layer.parameters()) == 0:
                layer.init_params(layer._input_shape)
            for v in layer.parameters():
                if v is not None:
                    self._parameters.add(v)
        self.loss = objectives.get_objective(loss)
        self.optimizer = get_optimizer(optimizer, parameters=self._parameters, **kwargs)

    def forward(self, x, *args, **kwargs):
        self.inputs._in_bounds = [x]
        out = None
        for layer in self._graph:
            out = layer.forward()
        return out


def _layer_topological_sort(ins, outs):
    topo, seen = [], set()

    def visit(v):
        if v is ins:
            if id(v) not in seen:
                seen.add(id(v))
                topo.append(v)
            return
        if id(v) in seen:
            return
        seen.add(id(v))
        for child in v.in_bounds:
            visit(child)
        topo.append(v)

    visit(outs)
    return topo


class Module(_Base):
    """PyTorch-style dynamic module (models.py:341-363): subclass, define forward(x)."""

    def __call__(self, x, *args, **kwargs):
        out = self.forward(x)
        self._collect_variables(x)
        out.is_leaf = True
        return out

    def _collect_variables(self, x):
        queue, seen = [x], {id(x)}
        while queue:
            vertex = queue.pop(0)
            for n in vertex.out_bounds:
                if id(n) not in seen:
                    for v in n.in_bounds:
                        if isinstance(v, Parameter):
                            self._parameters.add(v)
                    queue.append(n)
                    seen.add(id(n))

    def forward(self, x):
        raise NotImplementedError
