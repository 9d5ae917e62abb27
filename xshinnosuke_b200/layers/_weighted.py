"""What Conv2D and Dense share: two parameters (kernel, optional bias) created lazily from the input shape, and an
optional activation layer chained behind the op.

Reference behaviour kept (xs/layers/convolutional.py:61-83, xs/layers/linear.py:15-17, 43-60): the activation is a real
layer object whose output buffer is preallocated next to the op's own for the static graph, whose forward runs right
after the op, and whose backward runs right before the op's. (Dense exposes the activation's output as its `data`;
Conv2D does not -- SURVEY quirk: a functional `Model` reads the pre-activation buffer of a Conv2D.)
"""
from ..core.tensor import Parameter
from ..nn.initializers import get_initializer
from .act import get_activator
from .base import Layer


class WeightedLayer(Layer):
    def __init__(self, use_bias, activation, kernel_initializer, bias_initializer, **kwargs):
        self.use_bias = use_bias
        self.activation = None if activation is None else get_activator(activation)
        self.kernel_initializer = get_initializer(kernel_initializer)
        self.bias_initializer = get_initializer(bias_initializer)
        super().__init__(**kwargs)

    # ---- parameters
    def _make_params(self, kernel_shape, bias_shape, keep_bias_slot):
        """kernel, then bias; `keep_bias_slot`: a bias-less layer still has a (None) second entry."""
        self._parameters.append(Parameter(self.kernel_initializer(tuple(kernel_shape), requires_grad=True)))
        if self.use_bias:
            self._parameters.append(Parameter(self.bias_initializer(tuple(bias_shape), requires_grad=True)))
        elif keep_bias_slot:
            self._parameters.append(None)

    def _kernel_and_bias(self):
        bias = self._parameters[1] if (self.use_bias and len(self._parameters) > 1) else None
        return self._parameters[0], bias

    # ---- the chained activation
    def _static_out(self):
        """The preallocated output buffer to write into (static graph), else None."""
        buf = self._data
        return buf if (buf is not None and buf.is_static) else None

    def _activate(self):
        return self._data if self.activation is None else self.activation.forward(self._data)

    def init_layer_out_tensor(self, x=None):
        super().init_layer_out_tensor(x)
        if self.activation is not None:
            self.activation.compute_output_shape(self._activation_input_shape())
            self.activation.init_layer_out_tensor(self._data)

    def _activation_input_shape(self):
        return self._shape

    def backward(self, gradients=None):
        if self.activation is not None:
            self.activation.backward()
        super().backward()
