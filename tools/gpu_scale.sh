#!/bin/bash
# scaling check: N GPUs (torchrun) then 1 GPU on the same box; every command under its own timeout
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; echo "n$N rc=$?"
cat gpurun_out/bench_c4_n$N.json | cut -c1-1800; tail -5 gpurun_out/bench_c4_n$N.err
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err; echo "n1 rc=$?"
cat gpurun_out/bench_c4_n1.json | cut -c1-1800; tail -3 gpurun_out/bench_c4_n1.err
