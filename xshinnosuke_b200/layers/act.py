"""ReLU / Activation layers (mirror of xs/layers/activators.py:4-60, 74-87)."""
import copy

from ..core import state as GLOBAL
from ..core.device import rt
from ..nn import functional as F
from ..utils.common import initialize_ops_grad
from .base import Layer


class ReLU(Layer):
    """`inplace=True` follows the reference protocol (activators.py:37-50): the producer's closure
    is stacked in `cache['grad_fn']`, the tensor's grad_fn becomes ReLUBackward, and the x<0 mask is
    recorded -- here as one BIT per element, written by the same kernel pass that rectifies in place."""

    def __init__(self, inplace=False, **kwargs):
        super().__init__(**kwargs)
        self.inplace = inplace

    def init_layer_out_tensor(self, x=None):
        x = self._in_bounds[0].data if x is None else x
        if self.inplace:
            self._data = x
        else:
            super().init_layer_out_tensor(x)

    def call(self, x, *args, **kwargs):
        if self.inplace:
            self._data = x
            track = GLOBAL.TRAINING and GLOBAL.COMPUTE_GRAD and x.requires_grad
            mask = None
            if track:
                x.cache.setdefault("grad_fn", []).append(x.grad_fn)
                mask = rt.empty((x.numel() + 7) // 8, "uint8")
                x.cache["mask"] = mask
                x.grad_fn = F.ReLUBackward
            arr = x.arr
            producer = arr.pending_op
            if track and producer is not None and getattr(producer, "fusable_relu", False):
                # the producer (BN apply / residual add) has not run yet: run it fused with this ReLU + mask
                if GLOBAL.INPUTS is None:
                    GLOBAL.INPUTS = x
                arr.pending_op = None
                bn = getattr(producer, "bn", None) if rt.fuse_bn_relu_pool else None
                if bn is not None:
                    # BatchNorm -> ReLU: stay deferred one more op -- a max-pool that follows runs BN apply + ReLU + mask
                    # + pool as ONE kernel (F.max_pool2d); anything else launches the fused BN+ReLU kernel on first touch
                    def run2(m=None, _p=producer, _mask=mask):
                        _p(_mask)
                    run2.fusable_relu = False
                    run2.bn_relu = dict(bn, mask=mask)
                    arr.pending_op = run2
                else:
                    producer(mask)
            else:
                F.relu(x, True, None, _mask=mask)
            if track:
                x.cache["inplace"] = True
                initialize_ops_grad(x)
            return x
        self._data = F.relu(x, False, self._data if (self._data is not None and self._data.is_static) else None)
        if GLOBAL.TRAINING and x.requires_grad:
            self._data.cache["inplace"] = False
            initialize_ops_grad(x)
        return self._data

    def backward(self, gradients=None):
        if gradients is not None:
            self._data.grad = gradients
        if self._data.grad_fn is not None:
            self._data.grad_fn(self._data)
            if not self.inplace:
                self._data.zero_grad()


class Activation(Layer):
    def __init__(self, act_name="relu", **kwargs):
        self.activation = get_activator(act_name)
        super().__init__(**kwargs)

    def init_layer_out_tensor(self, x=None):
        super().init_layer_out_tensor(x)
        self.activation._data = self._data

    def call(self, x, *args, **kwargs):
        self.activation._data = self._data
        self._data = self.activation.call(x)
        return self._data


class _Unary(Layer):
    _fn = None

    def call(self, x, *args, **kwargs):
        out = self._data if (self._data is not None and self._data.is_static) else None
        self._data = type(self)._fn(x, out)
        return self._data


class Sigmoid(_Unary):
    """xs/layers/activators.py:62-65."""
    _fn = staticmethod(F.sigmoid)


class Tanh(_Unary):
    """xs/layers/activators.py:68-71."""
    _fn = staticmethod(F.tanh)


def get_activator(activator):
    if isinstance(activator, str):
        known = {"relu": ReLU, "sigmoid": Sigmoid, "tanh": Tanh}
        if activator.lower() in known:
            return known[activator.lower()]()
        raise NotImplementedError(f"activation {activator!r} is not provided (relu, sigmoid, tanh)")
    if isinstance(activator, Layer):
        return copy.deepcopy(activator)
    raise TypeError("unknown activator type!")
