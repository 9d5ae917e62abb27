"""Layer base class (mirror of xs/layers/base.py:6-134).

Two call protocols, as in the reference: `layer(tensor)` runs the op eagerly (dynamic graph, lazy
parameter creation from the input shape); `layer(prev_layer)` only records topology for the
functional `Model` (static graph), whose `fit` then drives `init_layer_out_tensor` / `forward` /
`backward` per batch with preallocated, reusable output buffers.
"""
from ..core import state as GLOBAL
from ..core.tensor import Tensor
from ..nn.initializers import Zeros


class Layer:
    def __init__(self, trainable=True, name=None, in_bounds=None, input_shape=None, shape=None, parameters=None,
                 **kwargs):
        self.trainable = trainable
        self._in_bounds = in_bounds if in_bounds is not None else []
        self._out_bounds = []
        self._input_shape = input_shape
        self.name = str(name)
        self._shape = shape
        self._parameters = parameters if parameters is not None else []
        self._data = None
        super().__init__(**kwargs)

    def __call__(self, x, *args, **kwargs):
        if isinstance(x, Tensor):
            if len(self._parameters) == 0:
                self.init_params(x.shape[1:])
            x.next_layers.append(self)
            return self.call(x, *args, **kwargs)
        self._shape = self.compute_output_shape(x.shape)
        self._in_bounds.append(x)
        self._input_shape = x.shape
        x.add_out_bounds(self)
        return self

    def connect(self, inbound=None):
        if inbound is None:
            if self._input_shape is None:
                raise ValueError("must specify input_shape")
        else:
            self._input_shape = inbound.shape
        self._shape = self.compute_output_shape(self._input_shape)
        if inbound is not None:
            self._in_bounds.append(inbound)
            inbound.add_out_bounds(self)

    def params_count(self):
        return sum(v.numel() for v in self._parameters if v is not None)

    @property
    def shape(self):
        return self._shape

    @property
    def out_bounds(self):
        return self._out_bounds

    @property
    def in_bounds(self):
        return self._in_bounds

    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, v):
        self._data = v

    def parameters(self, parameters=None):
        if parameters is None:
            return self._parameters
        self._parameters = parameters

    def add_out_bounds(self, *outs):
        for out in outs:
            if isinstance(out, Layer):
                self._out_bounds.append(out)

    def init_params(self, input_shape=None, *args):
        pass

    def _new_static(self, x, shape, in_tensors):
        t = Zeros()((x.shape[0],) + tuple(shape), requires_grad=self.trainable)
        t.to("static")
        t.add_in_bounds(*in_tensors)
        return t

    def init_layer_out_tensor(self, x=None):
        """Static graph: (re)allocate the persistent output buffer, or shrink it to a prefix view
        for a partial last batch (xs/layers/base.py:88-104)."""
        x = self._in_bounds[0].data if x is None else x
        if self._data is None or x.shape[0] > self._data.shape_capacity[0] or \
                (x.shape[0] < self._data.shape_capacity[0] and not GLOBAL.TRAINING):
            self._data = self._new_static(x, self.shape, [x] + list(self.parameters()))
        elif x.shape[0] < self._data.shape_capacity[0]:
            self._data.slices(slice(None, x.shape[0], None))
        else:
            self._data.slices(slice(None, None, None))

    def compute_output_shape(self, input_shape=None):
        self._shape = (input_shape,) if isinstance(input_shape, int) else tuple(input_shape)
        return self._shape

    def call(self, x, *args, **kwargs):
        raise NotImplementedError

    def forward(self, x=None, *args, **kwargs):
        if x is None:
            x = self._in_bounds[-1]
        if len(self._parameters) == 0:
            self.init_params(x.shape[1:])
        if isinstance(x, Tensor):
            if x.is_dynamic:
                x.next_layers.append(self)
            return self.call(x, *args, **kwargs)
        return self.call(x.data, *args, **kwargs)

    def backward(self, gradients=None):
        if gradients is not None:
            assert isinstance(gradients, Tensor) and gradients.shape == self._data.shape
            self._data.grad = gradients
        if self._data.grad_fn is not None:
            self._data.grad_fn(self._data)
            self._data.zero_grad()
