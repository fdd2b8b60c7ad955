#!/bin/bash
# N-GPU C4 bench (torchrun) with all legs, under its own timeout
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; echo "n$N rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c4_n$N.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ["value","ms_per_step","e2e","ms_per_step_by_rank","shard_rows_by_rank","rank_speed_weights","phases_ms_rank0","parity_spot_check"]}, d["roofline"]["achieved"])
PY
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_c4_n$N.err | tail -3
