"""Normalisation layers (mirror of xs/layers/normalize.py): BatchNormalization 19-52, Dropout 7-16, LayerNormalization 55-74,
GroupNormalization 77-101."""
from ..core import state as GLOBAL
from ..core.tensor import Parameter
from ..nn import functional as F
from ..nn.initializers import get_initializer
from .base import Layer


class BatchNormalization(Layer):
    """Layer defaults differ from the functional ones: epsilon=1e-6, momentum=0.99 (normalize.py:20)."""

    def __init__(self, epsilon=1e-6, momentum=0.99, axis=1, gamma_initializer="ones", beta_initializer="zeros",
                 moving_mean_initializer="zeros", moving_variance_initializer="ones", **kwargs):
        self.epsilon = epsilon
        self.axis = axis
        self.momentum = momentum
        self.gamma_initializer = get_initializer(gamma_initializer)
        self.beta_initializer = get_initializer(beta_initializer)
        self.moving_mean_initializer = get_initializer(moving_mean_initializer)
        self.moving_variance_initializer = get_initializer(moving_variance_initializer)
        self.moving_mean = None
        self.moving_variance = None
        super().__init__(**kwargs)

    def init_params(self, input_shape=None, *args):
        if input_shape is not None:
            self._input_shape = input_shape
        assert len(self._input_shape) >= 1
        n_in = self._input_shape[self.axis - 1]
        self._parameters.append(Parameter(self.gamma_initializer(n_in, requires_grad=True)))
        self._parameters.append(Parameter(self.beta_initializer(n_in, requires_grad=True)))
        self.moving_mean = self.moving_mean_initializer(n_in)
        self.moving_variance = self.moving_variance_initializer(n_in)

    def call(self, x, *args, **kwargs):
        gamma, beta = self._parameters
        if self.moving_mean is None:   # parameters injected by hand (tests/test_conv.py:26-29)
            n_in = gamma.numel()
            self.moving_mean = self.moving_mean_initializer(n_in)
            self.moving_variance = self.moving_variance_initializer(n_in)
        out = self._data if (self._data is not None and self._data.is_static) else None
        self._data = F.batch_norm(x, gamma, beta, self.moving_mean, self.moving_variance, self.axis, GLOBAL.TRAINING,
                                  self.epsilon, self.momentum, out)
        return self._data


class Dropout(Layer):
    """keep_prob = probability of KEEPING a unit (xs/layers/normalize.py:7-16)."""

    def __init__(self, keep_prob, **kwargs):
        self.keep_prob = keep_prob
        super().__init__(**kwargs)

    def call(self, x, *args, **kwargs):
        out = self._data if (self._data is not None and self._data.is_static) else None
        self._data = F.dropout2d(x, self.keep_prob, out)
        return self._data


class LayerNormalization(Layer):
    """Per-sample normalisation with per-element gamma / beta of the input shape (xs/layers/normalize.py:55-74)."""

    def __init__(self, epsilon=1e-10, gamma_initializer="ones", beta_initializer="zeros", **kwargs):
        self.epsilon = epsilon
        self.gamma_initializer = get_initializer(gamma_initializer)
        self.beta_initializer = get_initializer(beta_initializer)
        super().__init__(**kwargs)

    def init_params(self, input_shape=None, *args):
        if input_shape is not None:
            self._input_shape = input_shape
        self._parameters.append(Parameter(self.gamma_initializer(tuple(self._input_shape), requires_grad=True)))
        self._parameters.append(Parameter(self.beta_initializer(tuple(self._input_shape), requires_grad=True)))

    def call(self, x, *args, **kwargs):
        gamma, beta = self._parameters
        out = self._data if (self._data is not None and self._data.is_static) else None
        self._data = F.layer_norm(x, gamma, beta, self.epsilon, out)
        return self._data


class GroupNormalization(Layer):
    """xs/layers/normalize.py:77-101: per-(sample, group) statistics, per-channel gamma / beta."""

    def __init__(self, epsilon=1e-5, groups=16, gamma_initializer="ones", beta_initializer="zeros", **kwargs):
        self.epsilon = epsilon
        self.G = groups
        self.gamma_initializer = get_initializer(gamma_initializer)
        self.beta_initializer = get_initializer(beta_initializer)
        super().__init__(**kwargs)

    def init_params(self, input_shape=None, *args):
        if input_shape is not None:
            self._input_shape = input_shape
        c = self._input_shape[0]
        assert c % self.G == 0
        self._parameters.append(Parameter(self.gamma_initializer(c, requires_grad=True)))
        self._parameters.append(Parameter(self.beta_initializer(c, requires_grad=True)))

    def call(self, x, *args, **kwargs):
        gamma, beta = self._parameters
        out = self._data if (self._data is not None and self._data.is_static) else None
        self._data = F.group_norm(x, gamma, beta, self.epsilon, self.G, out)
        return self._data
