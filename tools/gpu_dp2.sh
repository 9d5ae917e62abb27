#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
run() {
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --no-parity --no-kernel-table "$@" > $O/dp_$name.json 2> $O/dp_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/dp_{n}.json")); print(f"{n:28s} value {d['value']:9.0f} ms {d['ms_per_step']:.4f} e2e {d['e2e']['value']:9.0f}")
except Exception as e: print(n, "ERR", e)
PY
}
for b in 148 146 144 140 132; do
  XSB_SM_BUDGET=$b timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-parity --no-kernel-table > $O/dp_n1_b$b.json 2> $O/dp_n1_b$b.err
  python -c "import json; d=json.load(open('gpurun_out/dp_n1_b$b.json')); print('n1 budget $b', round(d['value']), round(d['ms_per_step'],4))"
done
run base XSB_X=0 --
run b146_c2 XSB_SM_BUDGET=146 NCCL_MAX_CTAS=2 --
run b144_c4 XSB_SM_BUDGET=144 NCCL_MAX_CTAS=4 --
run b144_c8 XSB_SM_BUDGET=144 NCCL_MAX_CTAS=8 --
run b140_c8 XSB_SM_BUDGET=140 NCCL_MAX_CTAS=8 --
run b140_def XSB_SM_BUDGET=140 --
run b132_def XSB_SM_BUDGET=132 --
run b132_c16 XSB_SM_BUDGET=132 NCCL_MAX_CTAS=16 --
