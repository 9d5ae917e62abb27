#!/bin/bash
# round-2 multi-GPU evidence: the default bench line at N GPUs (R224 and R56), R56 with smaller gradient buckets
set -u
N=$1
O=gpurun_out
mkdir -p $O
run() {
  name=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --no-cpu-baseline --no-parity "$@" > $O/r02_bench_${name}_n$N.json 2> $O/r02_bench_${name}_n$N.err
  python - "$name" "$N" <<'PY'
import json,sys
n,N=sys.argv[1],sys.argv[2]
try:
    d=json.load(open(f"gpurun_out/r02_bench_{n}_n{N}.json")); print(f"{n:24s} N={N} value {d['value']:9.0f} ms {d['ms_per_step']:.4f} e2e {d['e2e']['value']:9.0f} clocks {d.get('clocks')}")
except Exception as e: print(n, "ERR", e)
PY
}
run r224_tf32 --steps 50 --warmup 5
run r56_tf32 --workload r56 --steps 200 --warmup 10 --no-kernel-table
run r56_tf32_b8 --workload r56 --steps 200 --warmup 10 --no-kernel-table --bucket-mb 8
run r56_tf32_b16 --workload r56 --steps 200 --warmup 10 --no-kernel-table --bucket-mb 16
