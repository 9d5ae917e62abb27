from .dense import Dense, Flatten  # noqa: F401
from .conv import Conv2D  # noqa: F401
from .pool import MaxPooling2D, AvgPooling2D  # noqa: F401
from .norm import BatchNormalization  # noqa: F401
from .act import ReLU, Activation  # noqa: F401
from .graph_ops import Input, Reshape, Add  # noqa: F401
