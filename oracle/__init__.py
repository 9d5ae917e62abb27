"""CPU oracle for the XS hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy (float64) restatement of the algorithm the reference
(E1eveNn/xshinnosuke, NumPy path) runs for conv2d / Dense / max_pool2d / avg_pool2d /
ReLU / BatchNormalization forward+backward, softmax cross-entropy, MSE, and the
SGD / Adam update. Every function cites the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / CPU baseline.
The product package ``xshinnosuke_b200`` never imports ``oracle``.

Parity pin: the oracle is checked in ``tests/test_oracle_golden.py`` against fixtures in
``tests/golden/*.npz`` that were produced by running the UNMODIFIED reference from
``/root/reference`` (script: ``tests/golden/make_golden.py``), and against the seeded
known-answer values in BASELINE.md section 4 (ResNet-18 loss trajectory etc.).
"""
