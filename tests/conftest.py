import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """Lazy view over tests/golden/reference_vectors.npz (made by tests/golden/make_golden.py
    from the unmodified reference)."""

    def __init__(self):
        self._z = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))

    def case(self, prefix):
        pre = prefix.rstrip("/") + "/"
        return {k[len(pre):]: self._z[k] for k in self._z.files if k.startswith(pre) and "/" not in k[len(pre):]}

    def names(self, prefix):
        pre = prefix.rstrip("/") + "/"
        return sorted({k[len(pre):].split("/")[0] for k in self._z.files if k.startswith(pre)})


@pytest.fixture(scope="session")
def golden():
    return Golden()
