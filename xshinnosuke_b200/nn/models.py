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
        GLOBAL.IS_TRAINING = True

    def eval(self):
        GLOBAL.IS_TRAINING = False

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

    def fit(self, x, y, batch_size=None, epochs=1, verbose=1, shuffle=True, validation_data=None,
            validation_split=0.0, initial_epoch=0, cuda_graph=False):
        """xs/nn/models.py:61-129. `cuda_graph=True` (no reference counterpart): the static graph's
        train step -- zero_grad, forward into the preallocated `out=` buffers, loss, backward, optimizer step -- is
        captured into ONE CUDA graph after the first two full batches (which run eagerly: every batch is exactly one
        update either way) and replayed for every further full-size batch; a partial last batch runs eagerly."""
        x, y = np.asarray(x), np.asarray(y)
        names = ["train_loss", "train_acc"]
        if validation_data is None and 0.0 < validation_split < 1.0:
            split = int(x.shape[0] * validation_split)
            validation_data = (x[-split:], y[-split:])
            train_x, train_y = x[:-split], y[:-split]
            names += ["validation_loss", "validation_acc"]
        else:
            train_x, train_y = x, y
        loader = DataLoader(DataSet(train_x, train_y), batch_size, shuffle)
        v_acc = v_loss = 0.0
        from ..core.device import rt
        use_graph = bool(cuda_graph) and validation_data is None and batch_size is not None and \
            rt.device is not None and rt.device.type == "cuda"
        gx = gy = graphed = None
        n_full = 0

        def train_step_static():
            self.train()
            self.optimizer.zero_grad()
            pred_ = self.forward(gx)
            loss_ = self.loss.forward(pred_, gy)
            self.backward(loss_)
            self.optimizer.step()
            return pred_, loss_
        with SummaryProfile(*names, verbose=verbose) as profile:
            for epoch in range(initial_epoch, epochs):
                t0 = time.time()
                if verbose != 0:
                    print("\033[1;31m Epoch[%d/%d]\033[0m" % (epoch + 1, epochs))
                bar = ProgressBar(max_iter=len(train_x), verbose=verbose)
                for idx, (batch_x, batch_y) in enumerate(loader):
                    full = use_graph and batch_x.shape[0] == batch_size
                    if full:
                        # fixed input tensors: every full batch is copied into them on the device
                        if gx is None:
                            gx, gy = Tensor(batch_x.arr.clone()), Tensor(batch_y.arr.clone())
                        else:
                            for dst, src in ((gx, batch_x), (gy, batch_y)):
                                src.arr.ptr  # noqa: B018 -- materialise
                                dst.arr.wptr  # noqa: B018
                                dst.arr.buf[: max(dst.arr.size, 1)].copy_(src.arr.buf[: max(src.arr.size, 1)])
                        if graphed is None or idx == 0:
                            self.init_graph_out_tensor(gx, gy)
                        if graphed is None and n_full >= 2:
                            from ..utils.graph import GraphedStep
                            graphed = GraphedStep(train_step_static, warmup=0)   # capture only, executes nothing
                        pred, loss = graphed() if graphed is not None else train_step_static()
                        n_full += 1
                        batch_y = gy
                    else:
                        if idx == 0 or idx == len(loader) - 1 or use_graph:
                            self.init_graph_out_tensor(batch_x, batch_y)
                        self.train()
                        self.optimizer.zero_grad()
                        pred = self.forward(batch_x)
                        loss = self.loss.forward(pred, batch_y)
                        self.backward(loss)
                        self.optimizer.step()
                    train_acc = self.loss.calc_acc(pred, batch_y)
                    loss_value = loss.item()
                    profile.step("train_acc", train_acc)
                    profile.step("train_loss", loss_value)
                    if validation_data is not None:
                        v_acc, v_loss = self.evaluate(validation_data[0], validation_data[1], batch_size=batch_size)
                        profile.step("validation_loss", v_loss)
                        profile.step("validation_acc", v_acc)
                    bar.update(batch_x.shape[0])
                    bar.console(verbose, time.time() - t0, loss_value, train_acc, v_loss, v_acc)
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

    def evaluate(self, x, y, batch_size=None):
        self.eval()
        x, y = np.asarray(x), np.asarray(y)
        if batch_size is not None:
            assert type(batch_size) is int
            loader = DataLoader(DataSet(x, y), batch_size)
            accs, losses = [], []
            for idx, (bx, by) in enumerate(loader):
                if idx == 0 or idx == len(loader) - 1:
                    self.init_graph_out_tensor(bx, by)
                pred = self.forward(bx)
                acc, loss = self.loss.metric(pred, by)
                accs.append(acc)
                losses.append(loss)
            return float(np.mean(accs)), float(np.mean(losses))
        tx, ty = Tensor(x), Tensor(y)
        self.init_graph_out_tensor(tx, ty)
        pred = self.forward(tx)
        return self.loss.metric(pred, ty)

    def predict(self, x, batch_size=None):
        self.eval()
        x = np.asarray(x)
        outs = []
        bs = batch_size if batch_size is not None else len(x)
        for i in range(0, len(x), bs):
            tx = Tensor(x[i:i + bs])
            self.init_graph_out_tensor(tx, None)
            outs.append(self.forward(tx).eval)
        return Tensor(np.concatenate(outs, axis=0))

    def save(self, save_path):
        """Pickle of host copies of every parameter (device buffers do not pickle)."""
        if not save_path.endswith(".pkl"):
            save_path += ".pkl"
        state = {"params": [p.eval for p in self._parameters],
                 "running": [(l.moving_mean.eval, l.moving_variance.eval) for l in self._graph
                             if getattr(l, "moving_mean", None) is not None]}
        with open(save_path, "wb") as f:
            pickle.dump(state, f)

    def load(self, model_path):
        if not model_path.endswith(".pkl"):
            model_path += ".pkl"
        with open(model_path, "rb") as f:
            state = pickle.load(f)
        for p, v in zip(self._parameters, state["params"]):
            p.eval = v
        bns = [l for l in self._graph if getattr(l, "moving_mean", None) is not None]
        for l, (mm, mv) in zip(bns, state["running"]):
            l.moving_mean.eval = mm
            l.moving_variance.eval = mv

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
