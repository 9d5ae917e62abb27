#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python tools/parity_probe.py 32 56 tf32,fp32x3 > $O/e_parity_probe.txt 2>&1; echo "probe rc=$?"
timeout 900 python tools/parity_probe.py 2 224 tf32,fp32x3 >> $O/e_parity_probe.txt 2>&1; echo "probe rc=$?"
grep -v Warn $O/e_parity_probe.txt | tail -20
timeout 600 python -m pytest tests -m gpu -x -q > $O/e_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/e_pytest_gpu.log
timeout 600 python bench.py --workload r224 --steps 30 --warmup 5 --no-cpu-baseline --no-parity > $O/e_bench_r224.json 2> $O/e_bench_r224.err; echo "r224 rc=$?"; tail -2 $O/e_bench_r224.err
timeout 600 python bench.py --workload r224 --math fp32x3 --steps 20 --warmup 5 --no-cpu-baseline > $O/e_bench_r224_x3.json 2> $O/e_bench_r224_x3.err; echo "r224 x3 rc=$?"; tail -2 $O/e_bench_r224_x3.err
python - <<'PY'
import json
for w in ("r224","r224_x3"):
    try:
        d=json.load(open(f"gpurun_out/e_bench_{w}.json"))
        kt=d["kernel_timing"]
        print(w, round(d["value"]), "ms", round(d["ms_per_step"],4), "ksum", round(kt["kernels_sum_ms"],4), kt["how"])
        r=d["roofline"]; print("   roofline", r and (r["kernel"], round(r["achieved"],1), round(r["frac"],3), r.get("frac_of_tf32_peak"), r.get("by_entry")))
        print("   parity", d.get("parity_checked"))
    except Exception as e: print(w, "ERR", repr(e))
PY
