import copy

from .optimizer import Optimizer, SGD, Adam  # noqa: F401


def get_optimizer(optimizer, **kwargs):
    """xs/optim/__init__.py:10-30 ('sgd' and 'adam' are on the hot path)."""
    parameters = kwargs.pop("parameters", None)
    lr = kwargs.pop("lr", 0.001)
    weight_decay = kwargs.pop("weight_decay", 0.0)
    if isinstance(optimizer, str):
        name = optimizer.lower()
        if name == "sgd":
            return SGD(parameters=parameters, lr=lr, weight_decay=weight_decay)
        if name == "adam":
            return Adam(parameters=parameters, lr=lr, weight_decay=weight_decay)
        raise NotImplementedError(f"optimizer {optimizer!r}: SURVEY.md 8f rank 3 (next); sgd and adam are built")
    if isinstance(optimizer, Optimizer):
        return copy.deepcopy(optimizer)
    raise ValueError("unknown optimizer type!")
