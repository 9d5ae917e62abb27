"""Conv2D layer (API of xs/layers/convolutional.py:7-83; the op is nn.functional.conv2d -> xsb_conv2d_fwd)."""
from ..nn import functional as F
from ..utils.common import pair
from ._weighted import WeightedLayer


def _window_out(size, k, stride, pad):
    return (size - k + 2 * pad) // stride + 1


class Conv2D(WeightedLayer):
    def __init__(self, out_channels, kernel_size, use_bias=True, stride=1, padding=0, activation=None,
                 kernel_initializer="Normal", bias_initializer="zeros", **kwargs):
        self.out_channels = out_channels
        self.kernel_size, self.stride = pair(kernel_size), pair(stride)
        self.padding = self.pad_size = padding          # pad_size: the integer the kernel gets ('SAME'/'VALID' resolved)
        super().__init__(use_bias, activation, kernel_initializer, bias_initializer, **kwargs)

    def init_params(self, input_shape=None, *args):
        if input_shape is not None:
            self._input_shape = input_shape
        in_channels = self._input_shape[0]
        # filter [OC, C, kh, kw] (channels-last on the device), bias [1, OC] (xs/layers/convolutional.py:31-37)
        self._make_params((self.out_channels, in_channels) + tuple(self.kernel_size), (1, self.out_channels),
                          keep_bias_slot=True)
        if isinstance(self.padding, str):                # dynamic graph: nobody else resolves the string
            self.compute_output_shape(self._input_shape)

    def _resolve_padding(self, h, kh):
        """-> integer padding for 'SAME' / 'VALID' / int (xs/layers/convolutional.py:43-59: SAME is derived from the
        height and the vertical stride only)."""
        if isinstance(self.padding, str):
            mode = self.padding.upper()
            if mode == "SAME":
                return (self.stride[0] * (h - 1) - h + kh) // 2
            if mode == "VALID":
                return 0
            raise TypeError("Unknown padding type!plz inputs SAME or VALID or an integer")
        assert isinstance(self.padding, int)
        return self.padding

    def compute_output_shape(self, input_shape=None):
        assert len(input_shape) == 3
        _, h, w = input_shape
        (kh, kw), (sh, sw) = self.kernel_size, self.stride
        self.pad_size = self._resolve_padding(h, kh)
        if isinstance(self.padding, str) and self.padding.upper() == "SAME":
            return self.out_channels, h, w               # the reference reports the input size, whatever the stride
        return self.out_channels, _window_out(h, kh, sh, self.pad_size), _window_out(w, kw, sw, self.pad_size)

    def call(self, x, *args, **kwargs):
        weight, bias = self._kernel_and_bias()
        self._data = F.conv2d(x, weight, bias, self.stride, self.pad_size, self._static_out())
        return self._activate()
