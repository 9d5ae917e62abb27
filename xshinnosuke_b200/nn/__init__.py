from ..core import state as GLOBAL  # noqa: F401
from ..core.tensor import Parameter, Tensor  # noqa: F401
from . import functional  # noqa: F401  (must precede objectives/models: they import the layers package)
from . import grad_fn  # noqa: F401
from .objectives import MSELoss, CrossEntropyLoss  # noqa: F401
from .models import Sequential, Model, Module  # noqa: F401
