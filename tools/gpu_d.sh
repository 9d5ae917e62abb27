#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python tools/parity_probe.py 32 56 fp32,tf32 > $O/d_parity_probe.txt 2>&1; echo "probe rc=$?"
timeout 600 python tools/parity_probe.py 2 224 fp32,tf32 >> $O/d_parity_probe.txt 2>&1; echo "probe rc=$?"
cat $O/d_parity_probe.txt | grep -v Warn
for w in r224 mlp; do
  timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-parity > $O/d_bench_$w.json 2> $O/d_bench_$w.err; echo "$w rc=$?"; tail -2 $O/d_bench_$w.err
done
python - <<'PY'
import json
for w in ("r224","mlp"):
    try:
        d=json.load(open(f"gpurun_out/d_bench_{w}.json"))
        kt=d["kernel_timing"]
        print(w, round(d["value"]), "ms", round(d["ms_per_step"],4), "ksum", round(kt["kernels_sum_ms"],4), kt["how"])
        r=d["roofline"]; print("   roofline", r and (r["kernel"], round(r["achieved"],1), round(r["frac"],3), r.get("frac_of_tf32_peak"), r.get("by_entry")))
        r=d.get("roofline_hbm"); print("   hbm", r and (r["kernel"], round(r["achieved"],1), round(r["frac"],3)))
        print("   ", {k:(v["us_per_step"], v["achieved"]) for k,v in list(d["kernels"].items())[:12]})
    except Exception as e: print(w, "ERR", repr(e))
PY
