"""Optimizers of the hot path: one fused multi-tensor kernel over the flat parameter buffer (optimizer.py)."""
import copy

from .optimizer import Optimizer, SGD, Adam, Momentum, RMSprop, AdaGrad, AdaDelta  # noqa: F401

_BY_NAME = {cls.__name__.lower(): cls for cls in (SGD, Adam, Momentum, RMSprop, AdaGrad, AdaDelta)}


def get_optimizer(optimizer, **kwargs):
    """String or instance -> optimizer, with the reference's keyword defaults (xs/optim/__init__.py:10-30: lr 0.001,
    weight_decay 0; an instance is deep-copied so that `compile` never shares state with the caller's object)."""
    if isinstance(optimizer, Optimizer):
        return copy.deepcopy(optimizer)
    cls = _BY_NAME.get(optimizer.lower()) if isinstance(optimizer, str) else None
    if cls is None:
        raise ValueError("unknown optimizer type!")
    return cls(parameters=kwargs.pop("parameters", None), lr=kwargs.pop("lr", 0.001),
               weight_decay=kwargs.pop("weight_decay", 0.0))
