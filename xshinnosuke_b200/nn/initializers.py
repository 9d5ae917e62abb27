"""Weight initializers and tensor factories (mirror of xs/nn/initializers.py + xs/nn/rng.py).

Host-side and one-off: values are drawn from the GLOBAL NumPy RNG exactly like the reference
(np.random.normal / uniform ..., float64) so a seeded run produces the oracle's weights, then
uploaded (fp32) to the device.
"""
import copy

import numpy as np

from ..core.tensor import Tensor


class Initializer:
    def __init__(self, seed=None):
        if seed is not None:
            np.random.seed(seed)

    @staticmethod
    def decompose_size(shape):
        shape = np.array(shape)
        if shape.size == 2:
            fan_in, fan_out = shape
        elif shape.size in (4, 5):
            field = np.prod(shape[2:])
            fan_in, fan_out = shape[1] * field, shape[0] * field
        else:
            fan_in = fan_out = int(np.sqrt(np.prod(shape)))
        return fan_in, fan_out


def _shape(shape):
    return (shape,) if isinstance(shape, (int, np.integer)) else tuple(shape)


class Zeros(Initializer):
    def __call__(self, shape, **kw):
        return Tensor(np.zeros(_shape(shape)), **kw)


class Ones(Initializer):
    def __call__(self, shape, **kw):
        return Tensor(np.ones(_shape(shape)), **kw)


class Empty(Zeros):
    pass


class Matrix(Initializer):
    def __init__(self, value=0.0, seed=None):
        self.value = value
        super().__init__(seed)

    def __call__(self, shape, **kw):
        return Tensor(np.full(_shape(shape), self.value), **kw)


class Uniform(Initializer):
    def __init__(self, scale=0.05, seed=None):
        self.scale = scale
        super().__init__(seed)

    def __call__(self, shape, **kw):
        return Tensor(np.random.uniform(-self.scale, self.scale, size=shape), **kw)


class Normal(Initializer):
    """Default for Conv2D / Dense kernels: N(0, 0.1) (xs/nn/initializers.py:60-67)."""

    def __init__(self, std=0.1, mean=0.0, seed=None):
        self.std, self.mean = std, mean
        super().__init__(seed)

    def __call__(self, shape, **kw):
        return Tensor(np.random.normal(loc=self.mean, scale=self.std, size=shape), **kw)


class LecunUniform(Initializer):
    def __call__(self, shape, **kw):
        fan_in, _ = self.decompose_size(shape)
        return Uniform(np.sqrt(3.0 / fan_in))(shape, **kw)


class GlorotUniform(Initializer):
    def __call__(self, shape, **kw):
        fan_in, fan_out = self.decompose_size(shape)
        return Uniform(np.sqrt(6.0 / (fan_in + fan_out)))(shape, **kw)


class HeUniform(Initializer):
    def __call__(self, shape, **kw):
        fan_in, _ = self.decompose_size(shape)
        return Uniform(np.sqrt(6.0 / fan_in))(shape, **kw)


class LecunNormal(Initializer):
    def __call__(self, shape, **kw):
        fan_in, _ = self.decompose_size(shape)
        return Normal(np.sqrt(1.0 / fan_in))(shape, **kw)


class GlorotNormal(Initializer):
    def __call__(self, shape, **kw):
        fan_in, fan_out = self.decompose_size(shape)
        return Normal(np.sqrt(2.0 / (fan_in + fan_out)))(shape, **kw)


class HeNormal(Initializer):
    def __call__(self, shape, **kw):
        fan_in, _ = self.decompose_size(shape)
        return Normal(np.sqrt(2.0 / fan_in))(shape, **kw)


_BY_NAME = {
    "zeros": Zeros, "ones": Ones, "uniform": Uniform, "normal": Normal,
    "heuniform": HeUniform, "he_uniform": HeUniform, "henormal": HeNormal, "he_normal": HeNormal,
    "lecunnormal": LecunNormal, "lecun_normal": LecunNormal, "lecununiform": LecunUniform,
    "lecun_uniform": LecunUniform, "glorotnormal": GlorotNormal, "glorot_normal": GlorotNormal,
    "glorotuniform": GlorotUniform, "glorot_uniform": GlorotUniform,
}


def get_initializer(initializer):
    """xs/nn/initializers.py:189-218."""
    if isinstance(initializer, str):
        cls = _BY_NAME.get(initializer.lower())
        if cls is None:
            raise TypeError("unknown initialization type!")
        return cls()
    if isinstance(initializer, Initializer):
        return copy.deepcopy(initializer)
    raise TypeError("unknown initialization type!")


# ---- tensor factories (xs/nn/rng.py:153-195)
def tensor(data, **kw):
    return Tensor(data=data, requires_grad=kw.pop("requires_grad", False), **kw)


def zeros(*shape, **kw):
    return Zeros()(shape, requires_grad=kw.pop("requires_grad", False), **kw)


def ones(*shape, **kw):
    return Ones()(shape, requires_grad=kw.pop("requires_grad", False), **kw)


def rand(*shape, **kw):
    return Tensor(np.random.rand(*shape), requires_grad=kw.pop("requires_grad", False), **kw)


def randn(*shape, **kw):
    return Tensor(np.random.randn(*shape), requires_grad=kw.pop("requires_grad", False), **kw)


def randint(low, high=None, shape=None, **kw):
    return Tensor(np.random.randint(low=low, high=high, size=shape), requires_grad=kw.pop("requires_grad", False), **kw)


def manual_seed_all(seeds=None):
    np.random.seed(seeds)
