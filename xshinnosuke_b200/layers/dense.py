"""Dense and Flatten layers (mirror of xs/layers/linear.py:5-76)."""
from functools import reduce

from ..core.tensor import Parameter
from ..nn import functional as F
from ..nn.initializers import get_initializer
from .act import get_activator
from .base import Layer


class Dense(Layer):
    def __init__(self, out_features, activation=None, use_bias=True, kernel_initializer="normal",
                 bias_initializer="zeros", **kwargs):
        self.out_features = out_features
        self.use_bias = use_bias
        self.kernel_initializer = get_initializer(kernel_initializer)
        self.bias_initializer = get_initializer(bias_initializer)
        self.activation = get_activator(activation) if activation is not None else None
        super().__init__(**kwargs)

    @property
    def data(self):
        return self._data if self.activation is None else self.activation.data

    @data.setter
    def data(self, v):
        if self.activation is None:
            self._data = v
        else:
            self.activation.data = v

    def compute_output_shape(self, input_shape=None):
        return (self.out_features,)

    def init_params(self, input_shape=None, *args):
        if input_shape is not None:
            self._input_shape = input_shape
        # kernel is [in, out], NOT transposed (xs/layers/linear.py:32); bias [1, out]
        self._parameters.append(Parameter(self.kernel_initializer(tuple(self._input_shape) + (self.out_features,),
                                                                   requires_grad=True)))
        if self.use_bias:
            self._parameters.append(Parameter(self.bias_initializer((1, self.out_features), requires_grad=True)))

    def init_layer_out_tensor(self, x=None):
        super().init_layer_out_tensor(x)
        if self.activation is not None:
            self.activation.compute_output_shape(self.out_features)
            self.activation.init_layer_out_tensor(self._data)

    def call(self, x, *args, **kwargs):
        out = self._data if (self._data is not None and self._data.is_static) else None
        if self.use_bias:
            weight, bias = self._parameters
            self._data = F.addmm(bias, x, weight, out)
        else:
            weight, = self._parameters
            self._data = F.mm(x, weight, out)
        if self.activation is not None:
            return self.activation.forward(self._data)
        return self._data

    def backward(self, gradients=None):
        if self.activation is not None:
            self.activation.backward()
        super().backward()


class Flatten(Layer):
    def __init__(self, start_dim=1, **kwargs):
        if start_dim < 1:
            raise ValueError("start_dim must be > 0")
        self.start_dim = start_dim
        super().__init__(**kwargs)

    def compute_output_shape(self, input_shape=None):
        assert len(input_shape) >= 3
        flat = reduce(lambda a, b: a * b, input_shape[self.start_dim - 1:])
        return tuple(input_shape[: self.start_dim - 1]) + (flat,)

    def call(self, x, *args, **kwargs):
        out = self._data if (self._data is not None and self._data.is_static) else None
        self._data = F.flatten(x, self.start_dim, out)
        return self._data
