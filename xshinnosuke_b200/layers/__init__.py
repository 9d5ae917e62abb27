from .dense import Dense, Flatten  # noqa: F401
from .conv import Conv2D  # noqa: F401
from .pool import MaxPooling2D, AvgPooling2D  # noqa: F401
from .norm import BatchNormalization, Dropout, GroupNormalization, LayerNormalization  # noqa: F401
from .act import Activation, ReLU, Sigmoid, Tanh  # noqa: F401
from .graph_ops import Add, Input, Reshape, ZeroPadding2D  # noqa: F401
