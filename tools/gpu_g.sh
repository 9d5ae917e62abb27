#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/g_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/g_pytest_gpu.log
for f in 0 1; do
  XSB_BN_FUSED=$f timeout 300 python tests/op_bench.py --ops bn_bwd --iters 20 > $O/g_opbench_bn_fused$f.log 2>&1; echo "opbench fused=$f rc=$?"
  XSB_BN_FUSED=$f timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-parity --no-kernel-table > $O/g_bench_fused$f.json 2> $O/g_bench_fused$f.err; echo "bench fused=$f rc=$?"
  XSB_BN_FUSED=$f timeout 300 python bench.py --workload r56 --steps 100 --warmup 5 --no-cpu-baseline --no-parity --no-kernel-table > $O/g_bench_r56_fused$f.json 2> $O/g_bench_r56_fused$f.err; echo "bench r56 fused=$f rc=$?"
done
paste $O/g_opbench_bn_fused0.log $O/g_opbench_bn_fused1.log | awk '{print $1,$2,$3,$4,$5, "|", $13,$14,$15}' | head -20
python - <<'PY'
import json
for t in ("fused0","fused1","r56_fused0","r56_fused1"):
    try:
        d=json.load(open(f"gpurun_out/g_bench_{t}.json")); print(t, round(d["value"]), round(d["ms_per_step"],4))
    except Exception as e: print(t, "ERR", e)
PY
