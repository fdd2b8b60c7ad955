#!/bin/bash
# 8-GPU session: C4 bench (peer-memory merge) and the C5 configuration (100M x 1024 bf16 rows, 1M queries, k = 1000)
N=${1:-8}
mkdir -p gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; echo "$2 rc=$?"; cut -c1-2200 gpurun_out/$2.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/$2.err | tail -5; }
run 300 bench_c4_n${N}_v2 --steps 3 --warmup 3 --no-cpu-baseline
run 400 bench_c5_n${N} --nb 100000000 --nq 1000000 --k 1000 --bf16-storage --steps 1 --warmup 1 --no-e2e --no-phases --no-cpu-baseline
