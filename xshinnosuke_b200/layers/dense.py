"""Dense and Flatten layers (API of xs/layers/linear.py:5-76; ops: nn.functional.addmm / mm / flatten)."""
from functools import reduce
from operator import mul

from ..nn import functional as F
from ._weighted import WeightedLayer
from .base import Layer


class Dense(WeightedLayer):
    def __init__(self, out_features, activation=None, use_bias=True, kernel_initializer="normal",
                 bias_initializer="zeros", **kwargs):
        self.out_features = out_features
        super().__init__(use_bias, activation, kernel_initializer, bias_initializer, **kwargs)

    # the layer's visible output is the activation's when there is one (xs/layers/linear.py:15-17)
    @property
    def data(self):
        return self.activation.data if self.activation is not None else self._data

    @data.setter
    def data(self, v):
        if self.activation is not None:
            self.activation.data = v
        else:
            self._data = v

    def compute_output_shape(self, input_shape=None):
        return (self.out_features,)

    def _activation_input_shape(self):
        return self.out_features

    def init_params(self, input_shape=None, *args):
        if input_shape is not None:
            self._input_shape = input_shape
        # kernel [in, out] -- NOT transposed (xs/layers/linear.py:32), bias [1, out]
        self._make_params(tuple(self._input_shape) + (self.out_features,), (1, self.out_features), keep_bias_slot=False)

    def call(self, x, *args, **kwargs):
        weight, bias = self._kernel_and_bias()
        out = self._static_out()
        self._data = F.mm(x, weight, out) if bias is None else F.addmm(bias, x, weight, out)
        return self._activate()


class Flatten(Layer):
    def __init__(self, start_dim=1, **kwargs):
        if start_dim < 1:
            raise ValueError("start_dim must be > 0")
        self.start_dim = start_dim
        super().__init__(**kwargs)

    def compute_output_shape(self, input_shape=None):
        assert len(input_shape) >= 3
        lead = self.start_dim - 1                        # input_shape excludes the batch axis
        return tuple(input_shape[:lead]) + (reduce(mul, input_shape[lead:]),)

    def call(self, x, *args, **kwargs):
        out = self._data if (self._data is not None and self._data.is_static) else None
        self._data = F.flatten(x, self.start_dim, out)
        return self._data
