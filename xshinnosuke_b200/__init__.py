"""xshinnosuke_b200 -- B200-native backend for the XShinnosuke ("XS") data-parallel hot path.

Drop-in for the reference's `xs` package on that path:  `import xshinnosuke_b200 as xs`.
Same `xs.layers` / `xs.nn.functional` / `Module` / `Sequential` / `Model` / `xs.optim` API, but every
tensor lives in B200 HBM and every op is a hand-written sm_100a CUDA kernel behind the C ABI of
`libxsb200.so` (include/xsb200.h). There is no CPU, NumPy or PyTorch compute fallback.
"""
import numpy as _np

from . import nn, layers, optim, utils  # noqa: F401
from .core.device import rt as runtime, set_math_mode, get_math_mode  # noqa: F401
from .nn.initializers import manual_seed_all, randn, tensor, zeros, ones, rand, randint  # noqa: F401
from .utils.toolkit import gradient_check, SummaryProfile  # noqa: F401
from .utils.common import no_grad  # noqa: F401
from .utils.graph import GraphedStep  # noqa: F401

float32 = _np.dtype(_np.float32)
float64 = _np.dtype(_np.float64)
int64 = _np.dtype(_np.int64)
int32 = _np.dtype(_np.int32)


def cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
