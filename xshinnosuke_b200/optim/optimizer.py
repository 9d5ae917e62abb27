"""SGD / Adam as ONE fused kernel launch over a flat parameter buffer (mirror of xs/optim/optimizer.py:5-20,
sgd.py:4-9, adam.py:4-30).

On the first `zero_grad()` / `step()` the optimizer flattens its parameters (creation order, each
tensor padded to a 64-float boundary) into one device buffer, rebinds every Parameter's storage to a
view of it, and lays the gradients out in a parallel flat buffer. `zero_grad()` then costs nothing
(gradients are flagged logically zero, see DevArray.pending_zero) and `step()` is a single
128-bit-vectorised kernel (xsb_sgd_step / xsb_adam_step). The parameter collection may be a live set
that is still empty at construction (tests/resnet/resnet18_dg.py:88 builds the optimizer before the
first forward); it is re-flattened whenever it grows. `weight_decay`: the reference accepts it and
never applies it (Q8); here a non-zero value adds the L2 term wd*p to the gradient inside the fused kernel
(g' = grad_scale*g + wd*p), the default 0 is the reference's arithmetic. The data-parallel engine (parallel.py) all-reduces buckets of the flat gradient
buffer and passes grad_scale = 1/world.
"""
import numpy as np

from ..core.device import DevArray, fill_floats, rt, torch
from ..core.tensor import Tensor
from .._lib import lib

ALIGN = 64  # floats; keeps every tensor 256-byte aligned inside the flat buffers (TMA-friendly)


class FlatState:
    """Flat parameter / gradient (/ moment) buffers and the per-tensor views into them."""

    def __init__(self, params, n_extra=0):
        self.params = list(params)
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.total = off
        t = torch()
        self.flat_p = t.zeros(max(off, 1), dtype=t.float32, device=rt.device)
        self.flat_g = t.zeros(max(off, 1), dtype=t.float32, device=rt.device)
        self.extra = [t.zeros(max(off, 1), dtype=t.float32, device=rt.device) for _ in range(n_extra)]
        self.gviews = []
        for p, o in zip(self.params, self.offsets):
            n = p.numel()
            old = p._arr
            view = self.flat_p[o:o + n]
            view.copy_(old.buf[:n]) if not old.pending_zero else view.zero_()
            old.buf = view
            old.pending_zero = False
            gview = self.flat_g[o:o + n]
            self.gviews.append(gview)
            pending = True
            if p._grad is not None and p._grad._arr is not None and not p._grad._arr.pending_zero:
                gview.copy_(p._grad._arr.buf[:n])
                pending = False
            p._grad = Tensor(DevArray(gview, old.shape, "float32", cl=old.cl, pending_zero=pending))

    def owns(self, params):
        return len(params) == len(self.params) and all(a is b for a, b in zip(params, self.params))


class Optimizer:
    n_state = 0

    def __init__(self, parameters=None, lr=0.01, weight_decay=0.0):
        self._parameters = parameters if parameters is not None else []
        self.lr = lr
        self.weight_decay = weight_decay
        self.iterations = 0
        self.flat = None
        self.grad_scale = 1.0        # set to 1/world by the data-parallel engine
        self.before_step = None      # hook: wait for gradient all-reduces

    def _ensure_flat(self):
        rt.ensure()
        params = [p for p in self._parameters if p is not None]
        if self.flat is None or not self.flat.owns(params):
            old = self.flat
            self.flat = FlatState(params, self.n_state)
            if old is not None:
                self._migrate_state(old, self.flat)
        return self.flat

    def _migrate_state(self, old, new):
        for k in range(self.n_state):
            for p, o in zip(old.params, old.offsets):
                if p in new.params:
                    no = new.offsets[new.params.index(p)]
                    new.extra[k][no:no + p.numel()].copy_(old.extra[k][o:o + p.numel()])

    # ---- persistence (Model.save / Model.load)
    _hyper = ("lr", "weight_decay")

    def state_dict(self):
        """Host copy of everything a resumed run needs: hyper-parameters, step counters and the per-parameter slices of
        the flat state buffers (Adam's v / m, Momentum's v, ...), in parameter order."""
        rt.ensure()
        flat = self._ensure_flat()
        slots = [[buf[o:o + p.numel()].cpu().numpy().reshape(p.shape) for p, o in zip(flat.params, flat.offsets)]
                 for buf in flat.extra]
        return {"class": type(self).__name__, "iterations": int(self.iterations), "slots": slots,
                "hyper": {k: getattr(self, k) for k in self._hyper if hasattr(self, k)},
                "step": int(self._step_dev.item()) if getattr(self, "_step_dev", None) is not None else None}

    def load_state_dict(self, state):
        if state["class"] != type(self).__name__:
            raise ValueError(f"checkpoint optimizer is {state['class']}, the model was compiled with {type(self).__name__}")
        flat = self._ensure_flat()
        if len(state["slots"]) != len(flat.extra) or any(len(s) != len(flat.params) for s in state["slots"]):
            raise ValueError("optimizer state does not match the model's parameters")
        t = torch()
        for buf, slot in zip(flat.extra, state["slots"]):
            for p, o, v in zip(flat.params, flat.offsets, slot):
                if tuple(v.shape) != tuple(p.shape):
                    raise ValueError(f"optimizer state shape {tuple(v.shape)} != parameter shape {tuple(p.shape)}")
                buf[o:o + p.numel()].copy_(t.from_numpy(np.ascontiguousarray(v, dtype=np.float32).reshape(-1)))
        for k, v in state["hyper"].items():
            setattr(self, k, v)
        self.iterations = int(state["iterations"])
        if state.get("step") is not None and hasattr(self, "_step_dev"):
            if self._step_dev is None:
                self._step_dev = t.zeros(1, dtype=t.int32, device=rt.device)
            self._step_dev.fill_(int(state["step"]))

    def zero_grad(self):
        """xs/optim/optimizer.py:12-17 -- logical: each gradient view is flagged zero, no kernel runs."""
        flat = self._ensure_flat()
        for p in flat.params:
            p._grad._arr.pending_zero = True
            p._grad._arr._pz[2] = None      # a postponed add belongs to the gradient that is being discarded
            p.pending_uses = 0              # forwards whose backward never ran (e.g. validation) are forgotten here

    def _materialize_grads(self):
        for p, gview in zip(self.flat.params, self.flat.gviews):
            g = p._grad._arr if p._grad is not None else None
            if g is None or g.buf.data_ptr() != gview.data_ptr():
                # the user replaced the gradient tensor (p.grad = Tensor(...)): bring it back into the flat buffer the
                # fused kernel reads, and rebind the view
                if g is not None and not g.pending_zero:
                    src = g if g.cl == p._arr.cl else (g.as_cl() if p._arr.cl else g.logical_rowmajor())
                    src.flush()
                    gview.copy_(src.buf[: p.numel()])
                    pending = False
                else:
                    pending = True
                p._grad = Tensor(DevArray(gview, p._arr.shape, "float32", cl=p._arr.cl, pending_zero=pending))
            if p._grad._arr.pending_zero:      # parameter received no gradient this step
                fill_floats(p._grad._arr.buf, 0.0)
                p._grad._arr.pending_zero = False

    def step(self):
        self.iterations += 1


class SGD(Optimizer):
    """p -= lr * g (xs/optim/sgd.py:5-9)."""

    def step(self):
        flat = self._ensure_flat()
        self._materialize_grads()
        if self.before_step is not None:
            self.before_step(self)
        if flat.total:
            lib.xsb_sgd_step(flat.flat_p.data_ptr(), flat.flat_g.data_ptr(), flat.total, float(self.lr),
                             float(self.grad_scale), float(self.weight_decay), rt.stream())
            rt.launches += 1
        super().step()


class Adam(Optimizer):
    """xs/optim/adam.py:4-30 (v = first moment, m = second moment, bias-corrected, eps outside the sqrt).
    The step counter lives on the device so a CUDA-graph replay advances the bias correction."""
    n_state = 2
    _hyper = ("lr", "weight_decay", "beta1", "beta2", "epsilon")

    def __init__(self, parameters=None, lr=0.001, weight_decay=0.0, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon
        self._step_dev = None
        super().__init__(parameters=parameters, lr=lr, weight_decay=weight_decay)

    def step(self):
        flat = self._ensure_flat()
        if self._step_dev is None:
            t = torch()
            self._step_dev = t.zeros(1, dtype=t.int32, device=rt.device)
        self._materialize_grads()
        if self.before_step is not None:
            self.before_step(self)
        self.iterations += 1
        st = rt.stream()
        lib.xsb_counter_inc(self._step_dev.data_ptr(), st)
        if flat.total:
            lib.xsb_adam_step(flat.flat_p.data_ptr(), flat.flat_g.data_ptr(), flat.extra[0].data_ptr(),
                              flat.extra[1].data_ptr(), flat.total, float(self.lr), float(self.beta1),
                              float(self.beta2), float(self.epsilon), float(self.grad_scale), float(self.weight_decay),
                              self._step_dev.data_ptr(), st)
        rt.launches += 2


class _RuleOptimizer(Optimizer):
    """Momentum / RMSprop / AdaGrad / AdaDelta as one fused launch of xsb_rule_step over the flat
    buffers. All four raise NameError in the reference as shipped (missing `np` import, Q7); the formulas are
    theirs (xs/optim/sgd.py:12-29, rmsprop.py:4-21, adagrad.py:4-18, adadelta.py:4-26)."""
    kind = 0
    n_state = 1
    rho = 0.0
    epsilon = 0.0
    _hyper = ("lr", "weight_decay", "rho", "epsilon")

    def step(self):
        flat = self._ensure_flat()
        self._materialize_grads()
        if self.before_step is not None:
            self.before_step(self)
        if flat.total:
            s2 = flat.extra[1].data_ptr() if self.n_state > 1 else None
            lib.xsb_rule_step(self.kind, flat.flat_p.data_ptr(), flat.flat_g.data_ptr(), flat.extra[0].data_ptr(), s2,
                              flat.total, float(self.lr), float(self.rho), float(self.epsilon), float(self.grad_scale),
                              float(self.weight_decay), rt.stream())
            rt.launches += 1
        self.iterations += 1


class Momentum(_RuleOptimizer):
    kind = 0

    def __init__(self, parameters=None, lr=0.01, weight_decay=0.0, rho=0.9):
        self.rho = rho
        super().__init__(parameters=parameters, lr=lr, weight_decay=weight_decay)


class RMSprop(_RuleOptimizer):
    kind = 1

    def __init__(self, parameters=None, lr=0.001, weight_decay=0.0, rho=0.9, epsilon=1e-7):
        self.rho, self.epsilon = rho, epsilon
        super().__init__(parameters=parameters, lr=lr, weight_decay=weight_decay)


class AdaGrad(_RuleOptimizer):
    kind = 2

    def __init__(self, parameters=None, lr=0.01, weight_decay=0.0, epsilon=1e-7):
        self.epsilon = epsilon
        super().__init__(parameters=parameters, lr=lr, weight_decay=weight_decay)


class AdaDelta(_RuleOptimizer):
    kind = 3
    n_state = 2

    def __init__(self, parameters=None, lr=1.0, weight_decay=0.0, rho=0.95, epsilon=1e-7):
        self.rho, self.epsilon = rho, epsilon
        super().__init__(parameters=parameters, lr=lr, weight_decay=weight_decay)
