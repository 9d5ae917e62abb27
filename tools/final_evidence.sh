#!/bin/bash
# Round-end evidence on ONE B200 (run under gpurun): bench lines, per-op sweep, ncu launch list, ncu --set full of the
# bandwidth kernels that changed last. Reports stay in /tmp (size); only CSV / JSON / logs go to gpurun_out/.
set -u
O=gpurun_out
timeout 200 python bench.py --steps 50 --warmup 10 --workload r56 2>$O/r56_err.log > $O/r56_n1.json
timeout 200 python tests/op_bench.py --shapes r224 --math tf32 --json $O/opbench_r224.json > $O/opbench_r224.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_d.csv \
    python bench.py --steps 2 --warmup 1 --graph off --no-cpu-baseline > $O/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:"col_sums_atomic|bn_bwd_apply|maxpool|wgrad_halo" --launch-count 45 -o /tmp/bw_kernels \
    python tests/op_bench.py --shapes r224 --math tf32 --layers stem,layer1 --ops bn_bwd,maxpool_fwd,maxpool_bwd,conv_wgrad --iters 2 \
    > $O/ncu_full.log 2>&1
ncu -i /tmp/bw_kernels.ncu-rep --page raw --csv > $O/ncu_bw_kernels_raw.csv 2>>$O/ncu_full.log
ls -la $O | tail -12
python - <<'PY'
import json
d = json.load(open("gpurun_out/r56_n1.json"))
print("r56", round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]))
PY
grep -E "bn_bwd|maxpool|conv_wgrad|conv_fwd|conv_dgrad" gpurun_out/opbench_r224.log | head -40
