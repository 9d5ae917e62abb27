#!/bin/bash
# round-2 GPU batch A: persistent multi-job conv kernel -- parity, then A/B timing against the per-tile kernels
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/a_smi.txt 2>&1
timeout 300 python tools/tf32_probe.py > $O/a_tf32_probe.txt 2>&1; echo "probe rc=$?"
CONV="--ops conv_fwd,conv_dgrad,conv_wgrad --iters 20"
timeout 600 python tests/op_bench.py $CONV --pers 0 --json $O/a_opbench_pers0.json > $O/a_opbench_pers0.log 2>&1; echo "opbench0 rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > $O/a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/a_pytest_gpu.log
XSB_TC_PERS=2 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "tf32 or postponed" > $O/a_pytest_pers2.log 2>&1; echo "pytest pers2 rc=$?"; tail -3 $O/a_pytest_pers2.log
timeout 600 python tests/op_bench.py $CONV --pers 1 --json $O/a_opbench_pers1.json > $O/a_opbench_pers1.log 2>&1; echo "opbench1 rc=$?"
timeout 600 python tests/op_bench.py $CONV --pers 2 --json $O/a_opbench_pers2.json > $O/a_opbench_pers2.log 2>&1; echo "opbench2 rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/a_bench_pers1.json 2> $O/a_bench_pers1.err; echo "bench rc=$?"
XSB_TC_PERS=2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/a_bench_pers2.json 2> $O/a_bench_pers2.err; echo "bench2 rc=$?"
XSB_TC_PERS=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/a_bench_pers0.json 2> $O/a_bench_pers0.err; echo "bench0 rc=$?"
grep -h "conv_dgrad\|conv_fwd" $O/a_opbench_pers0.log $O/a_opbench_pers1.log $O/a_opbench_pers2.log | sort -k2,2 -k1,1 | head -80
python - <<'PY'
import json
for t in ("pers0","pers1","pers2"):
    try:
        d=json.load(open(f"gpurun_out/a_bench_{t}.json")); print(t, round(d["value"]), d["ms_per_step"])
    except Exception as e: print(t, "ERR", e)
PY
