#!/bin/bash
# 8-GPU (or N-GPU) C4 bench only, under its own timeout
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; echo "n$N rc=$?"
cat gpurun_out/bench_c4_n$N.json | cut -c1-1800; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_c4_n$N.err | tail -5
