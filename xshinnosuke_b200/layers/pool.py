"""MaxPooling2D / AvgPooling2D layers (mirror of xs/layers/pooling.py:5-55)."""
from ..nn import functional as F
from ..utils.common import pair
from .base import Layer


class _Pool2D(Layer):
    def __init__(self, kernel_size=2, stride=None, padding=0, **kwargs):
        if not isinstance(kernel_size, (int, float)):
            # the reference fails the same way (xs/layers/pooling.py:19): kernel_size must be an int
            raise TypeError("kernel_size must be an int")
        self.kernel_size = int(kernel_size)
        self.stride = pair(kernel_size) if stride is None else pair(stride)
        self.padding = padding
        super().__init__(**kwargs)

    def compute_output_shape(self, input_shape=None):
        c, h, w = input_shape
        oh = (h - self.kernel_size + 2 * self.padding) // self.stride[0] + 1
        ow = (w - self.kernel_size + 2 * self.padding) // self.stride[1] + 1
        return c, oh, ow

    def _out(self):
        return self._data if (self._data is not None and self._data.is_static) else None


class MaxPooling2D(_Pool2D):
    def call(self, x, *args, **kwargs):
        self._data = F.max_pool2d(x, self.kernel_size, self.stride, self.padding, self._out())
        return self._data


class AvgPooling2D(_Pool2D):
    def call(self, x, *args, **kwargs):
        self._data = F.avg_pool2d(x, self.kernel_size, self.stride, self.padding, self._out())
        return self._data
