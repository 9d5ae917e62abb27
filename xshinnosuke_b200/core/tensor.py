"""Tensor / Parameter: the type every op reads and writes (mirror of xs/core/base.py:6-371).

Same surface as the reference (`.eval` NCHW read/write port, `.grad`, `.cache`, in/out bounds,
`grad_fn`, `retain_grad`, static/dynamic flag, prefix `slices`, operator overloads, `backward()`),
but the payload is a DevArray in HBM. `.eval` / `.numpy()` / `.item()` are synchronising D2H reads
that return HOST COPIES in the reference's logical layout (mutating the returned ndarray does not
write through -- assign `t.eval = arr` instead).
"""
import numpy as np

from . import state as GLOBAL
from .device import DevArray, rt


class Tensor(object):
    def __new__(cls, *args, **kwargs):
        if len(args) and isinstance(args[0], Tensor):
            return args[0]
        return super().__new__(cls)

    def __init__(self, data=None, requires_grad=False, dtype=None, name=None, **kwargs):
        if isinstance(data, Tensor):
            return
        self._arr = None
        self._grad = None
        self.reset_(data, requires_grad, name, dtype, **kwargs)

    def reset_(self, data=None, requires_grad=False, name=None, dtype=None, **kwargs):
        if data is not None:
            if isinstance(data, DevArray):
                self._arr = data
            else:
                a = np.asarray(data, dtype=dtype)
                self._arr = DevArray.from_numpy(a)
        g = kwargs.pop("grad", None)
        self._grad = g if (g is None or isinstance(g, Tensor)) else Tensor(g)
        self._in_bounds = []
        self._out_bounds = []
        self.name = name
        self.grad_fn = None
        self.requires_grad = requires_grad
        self._slice = kwargs.pop("slices", slice(None, None, None))
        self._static = kwargs.pop("static_graph_tensor", False)
        self._cache = {}
        self._retain_grad = kwargs.pop("retain_grad", False)
        self._is_leaf = None
        self.next_layers = []

    # ------------------------------------------------------------------ storage
    def free_memory(self):
        self._arr = None
        self._grad = None

    @property
    def arr(self):
        """The active DevArray (prefix view when a static tensor holds a partial batch)."""
        a = self._arr
        if a is None:
            return None
        stop = self._slice.stop
        if stop is None or not a.shape or stop >= a.shape[0]:
            return a
        return a.prefix(stop)

    @property
    def data(self):
        if self._arr is None:
            return None
        t = Tensor(self._arr, slices=self._slice, static_graph_tensor=self._static, grad=self._grad)
        return t

    @data.setter
    def data(self, v):
        if isinstance(v, Tensor):
            self.eval = v.eval
        else:
            raise TypeError("Unknown data type {}".format(type(v)))

    @property
    def eval(self):
        return self.arr.to_numpy()

    @eval.setter
    def eval(self, v):
        self.arr.assign_numpy(np.asarray(v))

    def numpy(self):
        return self.eval

    @property
    def shape_capacity(self):
        return self._arr.shape

    @property
    def shape(self):
        s = self.arr.shape
        return s if len(s) != 0 else (1,)

    @property
    def dtype(self):
        return self._arr.dtype

    @property
    def device(self):
        return "cuda"

    def slices(self, slices=None):
        if slices is None:
            return self._slice
        assert isinstance(slices, slice)
        if slices.start not in (None, 0) or slices.step not in (None, 1):
            raise NotImplementedError("only prefix slices (slice(None, n)) are supported")
        self._slice = slices

    # ------------------------------------------------------------------ graph bookkeeping
    @property
    def is_leaf(self):
        return self._is_leaf if self._is_leaf is not None else len(self._in_bounds) == 0

    @is_leaf.setter
    def is_leaf(self, flag):
        self._is_leaf = flag

    @property
    def grad(self):
        if self._grad is not None:
            self._grad._slice = self._slice
        return self._grad

    @grad.setter
    def grad(self, g):
        self._grad = g if (g is None or isinstance(g, Tensor)) else Tensor(g)

    @property
    def cache(self):
        return self._cache

    @property
    def is_static(self):
        return self._static

    @property
    def is_dynamic(self):
        return not self._static

    def retain_grad(self, retain=None):
        if retain is None:
            return self._retain_grad
        self._retain_grad = retain

    def add_in_bounds(self, *ins):
        for i in ins:
            if isinstance(i, Tensor) or i is None:
                self._in_bounds.append(i)
            else:
                raise TypeError("Not a Tensor")

    def add_out_bounds(self, *outs):
        for o in outs:
            if isinstance(o, Tensor):
                self._out_bounds.append(o)
            else:
                raise TypeError("Not a Tensor")

    def requires_grad_(self, flag):
        self.requires_grad = flag

    @property
    def out_bounds(self):
        return self._out_bounds

    @property
    def in_bounds(self):
        return self._in_bounds

    # ------------------------------------------------------------------ device moves (always resident)
    def to(self, dst="static"):
        if dst == "static":
            self._static = True
            return self
        if dst == "dynamic":
            self._static = False
            return self
        if dst in ("cuda", "cpu"):
            # the reference returns a fresh tensor holding a converted copy (xs/core/base.py:143-159);
            # here the payload already lives in HBM, so the new tensor shares it.
            ret = Tensor(self._arr, requires_grad=self.requires_grad, name=self.name, slices=self._slice)
            if self._grad is not None:
                ret._grad = self._grad
            return ret
        raise ValueError(dst)

    def cuda(self):
        return self.to("cuda")

    def cpu(self):
        return self.to("cpu")

    def cuda_(self):
        pass

    def cpu_(self):
        pass

    # ------------------------------------------------------------------ python protocol
    def __repr__(self):
        rg = f", requires_grad={self.requires_grad}" if self.requires_grad else ""
        gf = f", grad_fn=<{self.grad_fn.__name__}>" if self.grad_fn else ""
        body = self.eval if self._arr is not None else None
        return f"{self.__class__.__name__}({body}{rg}{gf})"

    __str__ = __repr__

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other

    def __add__(self, other):
        from ..nn import functional as F
        return F.add(self, Tensor(other))

    __radd__ = __add__

    def __matmul__(self, other):
        from ..nn import functional as F
        return F.mm(self, Tensor(other))

    def dot(self, other):
        return self.__matmul__(other)

    mm = dot

    def numel(self):
        return self._arr.size

    def item(self):
        assert self._arr.size == 1
        v = self.eval.reshape(-1)[0].item()
        if v != v and self._cache.get("label_check") and rt.take_label_error():
            # the reference's NumPy indexing raises at the call (xs/nn/td_functional.py:119-127); here at the first read
            raise IndexError("cross_entropy: a label is outside [0, number of classes)")
        return v

    def long_(self):
        if self._arr.dtype != "int64":
            self._arr = DevArray.from_numpy(self._arr.to_numpy().astype(np.int64))

    def zero_(self):
        self.arr.fill(0.0)

    def ones_(self):
        self.fill_(1.0)

    def fill_(self, value):
        self.arr.fill(value)

    def zero_grad(self):
        """Logical zeroing: no memset, the first kernel that writes the gradient overwrites it."""
        if self._grad is None or self._grad._arr is None or self._grad._arr.shape != self._arr.shape:
            self._grad = Tensor(DevArray.empty(self._arr.shape, "float32", pending_zero=True))
        else:
            self._grad._arr.pending_zero = True
            self._grad._arr._pz[2] = None      # a postponed add belongs to the gradient that is being discarded

    def sum(self, axis=None, keepdims=False):
        from ..nn import functional as F
        return F.sum(self, axis, keepdims)

    def mean(self, axis=None, keepdims=False):
        from ..nn import functional as F
        return F.mean(self, axis, keepdims)

    def l2norm(self):
        return Tensor(np.asarray(np.sqrt(np.sum(np.square(self.eval.astype(np.float64))))))

    def view(self, *shapes):
        from ..nn import functional as F
        if len(shapes) == 1 and isinstance(shapes[0], (tuple, list)):
            shapes = tuple(shapes[0])
        return F.view(self, shapes)

    def size(self, *dims):
        if len(dims) == 0:
            return self.shape
        if len(dims) == 1:
            return self.shape[dims[0]]
        return (self.shape[d] for d in dims)

    def backward(self, gradients=None, retain_graph=False):
        from . import autograd
        if self.grad_fn is None:
            raise ValueError("can not solve grad on %s" % self)
        if gradients is not None:
            g = gradients.eval if isinstance(gradients, Tensor) else np.asarray(gradients)
            assert g.size == self._arr.size and tuple(g.shape) == tuple(self.shape)
            self._grad = Tensor(np.asarray(g, dtype=np.float32))
        else:
            if self._arr.size == 1:
                if self._grad is None or self._grad._arr is None or self._grad._arr.pending_zero:
                    self._grad = Tensor(DevArray.full(self._arr.shape, 1.0))
            else:
                raise ValueError("grad can be implicitly created only for scalar outputs")
        GLOBAL.OUTPUTS = self
        autograd.backward(self, GLOBAL.INPUTS, retain_graph)


class Parameter(Tensor):
    """Trainable leaf (xs/core/base.py:360-371). Carries a creation ticket: the flat optimizer /
    gradient buffers are laid out in creation order (the reference keeps an unordered set, Q10)."""
    _ticket = 0

    def __new__(cls, *args, **kwargs):
        if len(args) and isinstance(args[0], Parameter):
            return args[0]
        return object.__new__(cls)

    def __init__(self, t=None, **kwargs):
        if isinstance(t, Parameter):
            return
        if isinstance(t, Tensor):
            Tensor.__init__(self, t._arr, t.requires_grad, None, t.name, **kwargs)
        else:
            Tensor.__init__(self, t, requires_grad=True)
        Parameter._ticket += 1
        self.ticket = Parameter._ticket
        self.pending_uses = 0   # forward uses whose backward has not run yet (data-parallel bucket readiness)
