#!/bin/bash
set -u
N=$1
O=gpurun_out
mkdir -p $O
run() {
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-parity --no-kernel-table "$@" > $O/dp${N}_$name.json 2> $O/dp${N}_$name.err
  python - "$name" "$N" <<'PY'
import json,sys
n,N=sys.argv[1],sys.argv[2]
try:
    d=json.load(open(f"gpurun_out/dp{N}_{n}.json")); print(f"{n:24s} value {d['value']:9.0f} ms {d['ms_per_step']:.4f} e2e {d['e2e']['value']:9.0f}")
except Exception as e: print(n, "ERR", e)
PY
}
run b8_info NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL --
grep -i -m5 "nvls\|algo\|channels" $O/dp${N}_b8_info.err | cut -c1-200
run b64 XSB_X=0 -- --bucket-mb 64
run b24 XSB_X=0 -- --bucket-mb 24
run b8_nvls NCCL_ALGO=NVLS --
run b64_nvls NCCL_ALGO=NVLS -- --bucket-mb 64
