#!/bin/bash
# data-parallel tail: where do the +0.24 ms/step go when DP is on? (2 GPUs)
set -u
O=gpurun_out
mkdir -p $O
run() {  # name, env..., -- extra bench args
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
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-parity --no-kernel-table > $O/dp_n1.json 2> $O/dp_n1.err
python -c "import json; d=json.load(open('gpurun_out/dp_n1.json')); print('n1', round(d['value']), d['ms_per_step'])"
run base XSB_X=0 --
run maxctas1 NCCL_MAX_CTAS=1 --
run maxctas2 NCCL_MAX_CTAS=2 --
run maxctas4 NCCL_MAX_CTAS=4 --
run maxctas8 NCCL_MAX_CTAS=8 --
run bucket2 XSB_X=0 -- --bucket-mb 2
run bucket16 XSB_X=0 -- --bucket-mb 16
run bucket64 XSB_X=0 -- --bucket-mb 64
run nograph XSB_X=0 -- --graph off
run nvls NCCL_ALGO=NVLS --
run ring NCCL_ALGO=Ring --
run tree NCCL_ALGO=Tree --
run base_again XSB_X=0 --
