"""CPU: the C-ABI shared library loads without a GPU and exports every symbol include/xsb200.h declares
(no compute calls here), and the product refuses to run without CUDA instead of falling back."""
import ctypes
import os

import pytest

from xshinnosuke_b200._lib import HEADER, LIB_PATH, parse_header


def _ensure_built():
    if not os.path.exists(LIB_PATH):
        from xshinnosuke_b200.build import build_lib
        build_lib()


def test_header_declares_the_expected_surface():
    protos = parse_header(HEADER)
    for name in ("xsb_conv2d_fwd", "xsb_conv2d_dgrad", "xsb_conv2d_wgrad", "xsb_gemm", "xsb_maxpool2d_fwd",
                 "xsb_maxpool2d_bwd", "xsb_relu_fwd", "xsb_relu_bwd", "xsb_bn_train_fwd", "xsb_bn_bwd",
                 "xsb_sgd_step", "xsb_adam_step", "xsb_softmax_xent_fwd", "xsb_nchw_to_nhwc", "xsb_last_error"):
        assert name in protos, name
    assert len(protos) >= 36


def test_library_exports_every_declared_symbol():
    _ensure_built()
    dll = ctypes.CDLL(LIB_PATH)
    missing = [n for n in parse_header(HEADER) if not hasattr(dll, n)]
    assert not missing, missing
    dll.xsb_version.restype = ctypes.c_int
    assert dll.xsb_version() == 100
    dll.xsb_last_error.restype = ctypes.c_char_p
    assert isinstance(dll.xsb_last_error(), bytes)


def test_product_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    import xshinnosuke_b200 as xs
    from xshinnosuke_b200.core.device import Runtime
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Runtime().ensure()
    assert not xs.runtime.ready
    with pytest.raises(RuntimeError):
        xs.tensor(np.zeros((2, 2), dtype=np.float32))


def test_product_never_imports_the_oracle():
    import re
    root = os.path.dirname(LIB_PATH)
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), os.path.join(dirpath, f)
                assert "fake_backend" not in src, os.path.join(dirpath, f)
