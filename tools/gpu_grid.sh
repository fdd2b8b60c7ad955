#!/bin/bash
# N GPUs as Q query groups x (N / Q) row shards (GridIndexFlat)
N=${1:-2}; Q=${2:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --query-groups $Q --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_n${N}_q$Q.json 2> gpurun_out/bench_c4_n${N}_q$Q.err; echo "n$N q$Q rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_c4_n${N}_q$Q.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ["value","ms_per_step","e2e","ms_per_step_by_rank","shard_rows_by_rank","query_groups","parity_spot_check"]}, d["roofline"]["achieved"], d["config"]["sharding"])
except Exception as e:
    print("no result", e)
PY
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_c4_n${N}_q$Q.err | tail -5
