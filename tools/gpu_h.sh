#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "tf32 or fp32x3 or resnet" > $O/h_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/h_pytest.log
for f in 0 1; do
  XSB_TC_WGRAD_STRIP=$f timeout 300 python tests/op_bench.py --ops conv_wgrad --iters 20 --layers layer2,layer3,layer4 > $O/h_opbench_strip$f.log 2>&1; echo "opbench strip=$f rc=$?"; grep conv_wgrad $O/h_opbench_strip$f.log
  XSB_TC_WGRAD_STRIP=$f timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-parity --no-kernel-table > $O/h_bench_strip$f.json 2> $O/h_bench_strip$f.err; echo "bench strip=$f rc=$?"
done
python - <<'PY'
import json
for t in ("strip0","strip1"):
    try:
        d=json.load(open(f"gpurun_out/h_bench_{t}.json")); print(t, round(d["value"]), round(d["ms_per_step"],4))
    except Exception as e: print(t, "ERR", e)
PY
