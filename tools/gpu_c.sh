#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
for w in r224 r56 cnn mlp; do
  timeout 600 python bench.py --workload $w --steps 30 --warmup 5 > $O/c_bench_$w.json 2> $O/c_bench_$w.err; echo "$w rc=$?"; tail -3 $O/c_bench_$w.err
done
python - <<'PY'
import json
for w in ("r224","r56","cnn","mlp"):
    try:
        d=json.load(open(f"gpurun_out/c_bench_{w}.json"))
        kt=d["kernel_timing"]
        print(w, round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), "ksum", round(kt["kernels_sum_ms"],4), kt["how"][:60])
        r=d["roofline"]; print("   roofline", r and (r["kernel"], round(r["achieved"],1), round(r["frac"],3), r.get("frac_of_tf32_peak")))
        print("   parity", d.get("parity_checked")); print("   cpu", d.get("cpu_baseline",{}).get("value"))
    except Exception as e: print(w, "ERR", e)
PY
