#!/bin/bash
# round-2 FINAL evidence on 1 GPU: smoke, the bench lines of every workload / math mode (+ per-kernel list of the default
# step), the sustained run, the reference arms, the per-op sweep of the ResNet-18 layer shapes
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
b() { name=$1; shift; timeout 900 python bench.py "$@" > $O/r02_bench_$name.json 2> $O/r02_bench_$name.err; echo "$name rc=$?"; }
b r224_tf32_n1 --steps 50 --warmup 5 --dump-kernels $O/r02_kernels_r224_tf32.tsv
b r224_tf32_n1_steps2000 --steps 2000 --warmup 10 --no-cpu-baseline --no-parity
b r224_fp32x3_n1 --math fp32x3 --steps 30 --warmup 5 --no-cpu-baseline
b r224_fp32_n1 --math fp32 --steps 5 --warmup 3 --no-cpu-baseline
b r56_tf32_n1 --workload r56 --steps 200 --warmup 10 --dump-kernels $O/r02_kernels_r56_tf32.tsv
b cnn_tf32_n1 --workload cnn --steps 500 --warmup 10
b mlp_tf32_n1 --workload mlp --steps 500 --warmup 10
b reference_arm_r224 --impl reference --steps 10 --warmup 1
timeout 900 python tests/op_bench.py --iters 20 --ops conv_fwd,conv_fwd_stats,conv_dgrad,conv_wgrad,bn_fwd,bn_bwd,relu_fwd,relu_bwd,add,maxpool_fwd,maxpool_bwd,bnpool,colsum,sgd,adam --json $O/r02_opbench_r224_tf32.json > $O/r02_opbench_r224_tf32.log 2>&1; echo "opbench rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_*_n1*.json")) + ["gpurun_out/r02_bench_reference_arm_r224.json"]:
    try:
        d=json.load(open(f)); r=d.get("roofline") or {}
        print(f.split("r02_bench_")[1][:-5], round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1),
              "| roof", r.get("kernel","")[:22], round(r.get("achieved",0),1), round(r.get("frac",0),3), r.get("frac_of_tf32_peak"), "| clocks", d.get("clocks"), "| parity", (d.get("parity_checked") or {}).get("passed"))
    except Exception as e: print(f, "ERR", e)
PY
