"""Small host helpers shared by ops and layers (mirror of xs/utils/common.py:18-32)."""
from ..core import state as GLOBAL


class no_grad:
    """xs/utils/common.py:18-26: suspend grad_fn wiring (used by gradient_check)."""

    def __init__(self):
        self.compute_grad_flag = GLOBAL.COMPUTE_GRAD

    def __enter__(self):
        GLOBAL.COMPUTE_GRAD = False

    def __exit__(self, exc_type, exc_val, exc_tb):
        GLOBAL.COMPUTE_GRAD = self.compute_grad_flag


def initialize_ops_grad(*ops):
    """xs/utils/common.py:29-32: give every grad-requiring operand a (logically) zeroed grad."""
    for op in ops:
        if op is not None and op.grad is None and op.requires_grad:
            op.zero_grad()


def pair(v):
    if isinstance(v, (int, float)):
        return int(v), int(v)
    return int(v[0]), int(v[1])
