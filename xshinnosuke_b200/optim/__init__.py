import copy

from .optimizer import Optimizer, SGD, Adam, Momentum, RMSprop, AdaGrad, AdaDelta  # noqa: F401


def get_optimizer(optimizer, **kwargs):
    """xs/optim/__init__.py:10-30."""
    parameters = kwargs.pop("parameters", None)
    lr = kwargs.pop("lr", 0.001)
    weight_decay = kwargs.pop("weight_decay", 0.0)
    if isinstance(optimizer, str):
        name = optimizer.lower()
        if name == "sgd":
            return SGD(parameters=parameters, lr=lr, weight_decay=weight_decay)
        if name == "adam":
            return Adam(parameters=parameters, lr=lr, weight_decay=weight_decay)
        others = {"momentum": Momentum, "rmsprop": RMSprop, "adagrad": AdaGrad, "adadelta": AdaDelta}
        if name in others:
            return others[name](parameters=parameters, lr=lr, weight_decay=weight_decay)
        raise ValueError("unknown optimizer type!")
    if isinstance(optimizer, Optimizer):
        return copy.deepcopy(optimizer)
    raise ValueError("unknown optimizer type!")
