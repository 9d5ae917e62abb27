from . import state as GLOBAL  # noqa: F401
