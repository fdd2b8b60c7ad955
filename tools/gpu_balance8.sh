#!/bin/bash
# 8-GPU C4 bench: speed-proportional shards, then equal shards, same box
N=${1:-8}
mkdir -p gpurun_out
for mode in balanced equal; do
  flag=""; [ $mode = equal ] && flag="--no-balance"
  extra="--no-e2e"; [ $mode = balanced ] && extra=""
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline $extra $flag > gpurun_out/bench_c4_n${N}_$mode.json 2> gpurun_out/bench_c4_n${N}_$mode.err; echo "n$N $mode rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_c4_n${N}_$mode.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ["value","ms_per_step","e2e","ms_per_step_by_rank","shard_rows_by_rank","rank_speed_weights","phases_ms_rank0","parity_spot_check"]}, d["roofline"]["achieved"])
except Exception as e:
    print("no result", e)
PY
  grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_c4_n${N}_$mode.err | tail -3
done
