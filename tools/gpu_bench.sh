#!/bin/bash
# tests + small bench + full C4 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/t_gpu.log
tail -3 gpurun_out/t_gpu.log
timeout 600 python bench.py --nb 1000000 --nq 16384 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "small rc=$?"
cat gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
timeout 1500 python bench.py > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"
cat gpurun_out/bench_c4.json; tail -5 gpurun_out/bench_c4.err
