#!/bin/bash
# multi-GPU evidence: bench.py under torchrun at N = $1 (r224 + r56), and at N = 2 the NCCL data-parallel parity tests
set -u
N=$1
O=gpurun_out
mkdir -p $O
tr() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 "$@"; }
tr bench.py --gpus $N --steps 50 --warmup 5 > $O/r02_bench_r224_tf32_n$N.json 2> $O/r02_bench_r224_tf32_n$N.err; echo "r224 n$N rc=$?"
tr bench.py --gpus $N --steps 200 --warmup 10 --workload r56 > $O/r02_bench_r56_tf32_n$N.json 2> $O/r02_bench_r56_tf32_n$N.err; echo "r56 n$N rc=$?"
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "data_parallel" > $O/r02_pytest_dp_nccl.log 2>&1; echo "pytest dp rc=$?"; tail -3 $O/r02_pytest_dp_nccl.log
fi
python - "$N" <<'PY'
import json, sys
n=sys.argv[1]
for w in ("r224","r56"):
    try:
        d=json.load(open(f"gpurun_out/r02_bench_{w}_tf32_n{n}.json")); print(w, n, round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), d["clocks"])
    except Exception as e: print(w, n, "ERR", e)
PY
