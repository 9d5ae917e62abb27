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


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): ONE JSON line with the contract's keys, the
    line's own value repeated in `cpu_baseline` and `e2e`; under torchrun only rank 0 prints, the other ranks exit 0."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "mlp", "--steps", "2", "--warmup", "1"]
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-400:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == dict(value=d["value"], unit=d["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert "workload" in d["config"] and "model" not in d["config"]
    other = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, RANK="1", WORLD_SIZE="2"), timeout=300)
    assert other.returncode == 0 and other.stdout.strip() == ""
