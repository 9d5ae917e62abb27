#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/f_smoke.txt 2>&1; echo "smoke rc=$?"; tail -8 $O/f_smoke.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $O/f_pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 > $O/f_bench_r224.json 2> $O/f_bench_r224.err; echo "r224 rc=$?"; tail -2 $O/f_bench_r224.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/f_bench_r224.json"))
print(round(d["value"]), d["parity_checked"]); print(d["roofline"]["kernel"], d["roofline"]["achieved"], d["roofline"]["frac_of_tf32_peak"])
print(d["roofline_hbm"])
PY
