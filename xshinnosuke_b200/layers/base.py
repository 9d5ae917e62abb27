"""Layer base class (the API of xs/layers/base.py:6-134).

A layer is used in two ways, as in the reference:
  * eagerly -- `layer(tensor)`: parameters are created on first use from the input shape and the op runs (dynamic graph);
  * symbolically -- `layer(prev_layer)` / `connect(prev_layer)`: only the topology and the output shape are recorded; the
    functional `Model` / `Sequential` then drive `init_layer_out_tensor` / `forward` / `backward` per batch.

Two pieces the reference spreads over its methods live in one place each here: `_link` (the symbolic edge + shape
inference) and `StaticOutput` (the policy of the preallocated, reusable output buffer of the static graph).
"""
from ..core import state as GLOBAL
from ..core.tensor import Tensor
from ..nn.initializers import Zeros


class StaticOutput:
    """The static graph's per-layer output buffer: allocated once for the largest batch seen, narrowed to a prefix view for
    a smaller (last) batch while training, re-allocated at the exact size when not training (xs/layers/base.py:88-104)."""

    @staticmethod
    def fit(current, batch, sample_shape, requires_grad, operands):
        """-> the buffer to use for a batch of `batch` samples (`current` re-sliced, or a fresh one wired to `operands`)."""
        capacity = current.shape_capacity[0] if current is not None else -1
        if batch > capacity or (batch < capacity and not GLOBAL.TRAINING):
            fresh = Zeros()((batch,) + tuple(sample_shape), requires_grad=requires_grad)
            fresh.to("static")
            fresh.add_in_bounds(*operands)
            return fresh
        current.slices(slice(None, batch if batch < capacity else None, None))
        return current


class Layer:
    def __init__(self, trainable=True, name=None, in_bounds=None, input_shape=None, shape=None, parameters=None,
                 **kwargs):
        self.trainable = trainable
        self.name = str(name)
        self._in_bounds = [] if in_bounds is None else in_bounds
        self._out_bounds = []
        self._input_shape, self._shape = input_shape, shape
        self._parameters = [] if parameters is None else parameters
        self._data = None
        super().__init__(**kwargs)

    # ------------------------------------------------------------------ topology
    def _link(self, upstream):
        """Record the symbolic edge upstream -> self and infer the output shape from upstream's."""
        self._input_shape = upstream.shape
        self._shape = self.compute_output_shape(self._input_shape)
        self._in_bounds.append(upstream)
        upstream.add_out_bounds(self)

    def __call__(self, x, *args, **kwargs):
        if not isinstance(x, Tensor):          # a layer: functional-API wiring, nothing runs
            self._link(x)
            return self
        self._ensure_params(x)
        x.next_layers.append(self)
        return self.call(x, *args, **kwargs)

    def connect(self, inbound=None):
        """Sequential.compile: chain to the previous layer; the first layer needs an explicit input_shape."""
        if inbound is not None:
            self._link(inbound)
        elif self._input_shape is None:
            raise ValueError("must specify input_shape")
        else:
            self._shape = self.compute_output_shape(self._input_shape)

    @property
    def in_bounds(self):
        return self._in_bounds

    @property
    def out_bounds(self):
        return self._out_bounds

    def add_out_bounds(self, *outs):
        self._out_bounds.extend(o for o in outs if isinstance(o, Layer))

    @property
    def shape(self):
        return self._shape

    def compute_output_shape(self, input_shape=None):
        """Default: shape-preserving (activations, normalisation). An int means a flat feature count."""
        self._shape = (input_shape,) if isinstance(input_shape, int) else tuple(input_shape)
        return self._shape

    # ------------------------------------------------------------------ parameters
    def init_params(self, input_shape=None, *args):
        pass

    def _ensure_params(self, x):
        if not self._parameters:
            self.init_params(x.shape[1:])

    def parameters(self, parameters=None):
        if parameters is None:
            return self._parameters
        self._parameters = parameters

    def params_count(self):
        return sum(p.numel() for p in self._parameters if p is not None)

    # ------------------------------------------------------------------ data
    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, v):
        self._data = v

    def init_layer_out_tensor(self, x=None):
        x = self._in_bounds[0].data if x is None else x
        self._data = StaticOutput.fit(self._data, x.shape[0], self.shape, self.trainable, [x] + list(self.parameters()))

    # ------------------------------------------------------------------ execution
    def call(self, x, *args, **kwargs):
        raise NotImplementedError

    def forward(self, x=None, *args, **kwargs):
        """Static-graph step: x defaults to the upstream LAYER, whose current output tensor is consumed."""
        source = self._in_bounds[-1] if x is None else x
        self._ensure_params(source)
        if not isinstance(source, Tensor):
            return self.call(source.data, *args, **kwargs)
        if source.is_dynamic:
            source.next_layers.append(self)
        return self.call(source, *args, **kwargs)

    def backward(self, gradients=None):
        out = self._data
        if gradients is not None:
            assert isinstance(gradients, Tensor) and gradients.shape == out.shape
            out.grad = gradients
        if out.grad_fn is not None:
            out.grad_fn(out)       # accumulates into the operands' gradients ...
            out.zero_grad()        # ... after which this buffer's own gradient is spent
