"""Conv2D layer (mirror of xs/layers/convolutional.py:7-83)."""
from ..core.tensor import Parameter
from ..nn import functional as F
from ..nn.initializers import get_initializer
from ..utils.common import pair
from .act import get_activator
from .base import Layer


class Conv2D(Layer):
    def __init__(self, out_channels, kernel_size, use_bias=True, stride=1, padding=0, activation=None,
                 kernel_initializer="Normal", bias_initializer="zeros", **kwargs):
        self.out_channels = out_channels
        self.kernel_size = pair(kernel_size)
        self.use_bias = use_bias
        self.stride = pair(stride)
        self.padding = padding
        self.activation = get_activator(activation) if activation is not None else None
        self.kernel_initializer = get_initializer(kernel_initializer)
        self.bias_initializer = get_initializer(bias_initializer)
        self.pad_size = padding
        super().__init__(**kwargs)

    def init_params(self, input_shape=None, *args):
        if input_shape is not None:
            self._input_shape = input_shape
        kh, kw = self.kernel_size
        # weight [OC, C, kh, kw] (stored channels-last on the device), bias [1, OC] (Q13)
        weight = Parameter(self.kernel_initializer((self.out_channels, self._input_shape[0], kh, kw), requires_grad=True))
        self._parameters.append(weight)
        if self.use_bias:
            self._parameters.append(Parameter(self.bias_initializer((1, self.out_channels), requires_grad=True)))
        else:
            self._parameters.append(None)
        if isinstance(self.padding, str):
            self.compute_output_shape(self._input_shape)

    def compute_output_shape(self, input_shape=None):
        assert len(input_shape) == 3
        _, h, w = input_shape
        kh, kw = self.kernel_size
        if isinstance(self.padding, str):
            mode = self.padding.upper()
            if mode == "SAME":
                oh, ow = h, w
                self.pad_size = (self.stride[0] * (h - 1) - h + kh) // 2
            elif mode == "VALID":
                oh, ow = (h - kh) // self.stride[0] + 1, (w - kw) // self.stride[1] + 1
                self.pad_size = 0
            else:
                raise TypeError("Unknown padding type!plz inputs SAME or VALID or an integer")
        else:
            assert isinstance(self.padding, int)
            oh = (h - kh + 2 * self.padding) // self.stride[0] + 1
            ow = (w - kw + 2 * self.padding) // self.stride[1] + 1
            self.pad_size = self.padding
        return self.out_channels, oh, ow

    def init_layer_out_tensor(self, x=None):
        super().init_layer_out_tensor(x)
        if self.activation is not None:
            self.activation.compute_output_shape(self._shape)
            self.activation.init_layer_out_tensor(self._data)

    def call(self, x, *args, **kwargs):
        weight, bias = self._parameters[0], (self._parameters[1] if self.use_bias else None)
        out = self._data if (self._data is not None and self._data.is_static) else None
        self._data = F.conv2d(x, weight, bias, self.stride, self.pad_size, out)
        if self.activation is not None:
            return self.activation.forward(self._data)
        return self._data

    def backward(self, gradients=None):
        if self.activation is not None:
            self.activation.backward()
        super().backward()
