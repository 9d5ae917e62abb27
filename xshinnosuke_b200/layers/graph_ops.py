"""Input / Reshape / Add layers for the functional (static-graph) front end (mirror of xs/layers/ops.py)."""
from ..core import state as GLOBAL
from ..core.device import rt
from ..nn import functional as F
from ..nn.grad_fn import accumulate_into
from ..utils.common import initialize_ops_grad
from .._lib import lib
from .base import Layer, StaticOutput


class Input(Layer):
    def __init__(self, input_shape, **kwargs):
        super().__init__(input_shape=input_shape, **kwargs)
        self._shape = tuple(input_shape)

    def call(self, x, *args, **kwargs):
        self._data = x
        return self._data

    def forward(self, x=None, *args, **kwargs):
        if x is None:
            x = self._in_bounds[-1]
        self._data = x
        return x

    def backward(self, gradients=None):
        pass


class Reshape(Layer):
    def __init__(self, shape, **kwargs):
        super().__init__(shape=tuple(shape), **kwargs)

    def call(self, x, *args, **kwargs):
        out = self._data if (self._data is not None and self._data.is_static) else None
        self._data = F.view(x, (-1,) + tuple(self._shape), out)
        return self._data

    def compute_output_shape(self, input_shape=None):
        return self._shape


class ZeroPadding2D(Layer):
    """xs/layers/ops.py:26-37: zero padding of H and W by padding = (ph, pw) on both sides."""

    def __init__(self, padding, **kwargs):
        self.padding = (padding, padding) if isinstance(padding, int) else tuple(padding)
        super().__init__(**kwargs)

    def call(self, x, *args, **kwargs):
        out = self._data if (self._data is not None and self._data.is_static) else None
        self._data = F.pad2d(x, self.padding, out)
        return self._data

    def compute_output_shape(self, input_shape=None):
        self._shape = (input_shape[0], input_shape[1] + 2 * self.padding[0], input_shape[2] + 2 * self.padding[1])
        return self._shape


class Add(Layer):
    """Residual merge of the static graph (xs/layers/ops.py:40-82): out = sum of inbound layer outputs."""

    def __call__(self, inbounds, *args, **kwargs):
        for inbound in inbounds:
            self._in_bounds.append(inbound)
            inbound.add_out_bounds(self)
            self._shape = inbound.shape
        return self

    def init_layer_out_tensor(self, x=None):
        x = self._in_bounds[0].data if x is None else x
        self._data = StaticOutput.fit(self._data, x.shape[0], self.shape, self.trainable, [b.data for b in self._in_bounds])

    def forward(self, x=None, *args, **kwargs):
        ins = [b.data for b in self._in_bounds]
        y = self._data.arr
        if len(ins) == 1:
            src = ins[0].arr.as_cl()
            src.flush()
            y.buf[: y.size].copy_(src.buf[: y.size])
            y.pending_zero = False
        else:
            a, b = ins[0].arr.as_cl(), ins[1].arr.as_cl()
            lib.xsb_add(a.ptr, b.ptr, y.wptr, y.size, rt.stream())
            rt.launches += 1
            for extra in ins[2:]:
                lib.xsb_axpy(y.ptr, extra.arr.as_cl().ptr, 1.0, y.size, rt.stream())
                rt.launches += 1
        for t in ins:
            if GLOBAL.TRAINING and t.requires_grad:
                initialize_ops_grad(t)
            self._data.requires_grad = self._data.requires_grad or t.requires_grad
        return self._data

    def compute_output_shape(self, input_shape=None):
        return self._shape

    def backward(self, gradients=None):
        for inbound in self._in_bounds:
            t = inbound.data
            if t.requires_grad:
                accumulate_into(t, self._data.grad.arr)
        self._data.zero_grad()
