#!/bin/bash
# single-GPU session: tests, C4 bench, post-processing bench, smoke, ncu evidence for the new kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c4_n1_v2.json 2> gpurun_out/bench_c4_n1_v2.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/bench_c4_n1_v2.json
timeout 300 python tools/bench_postproc.py > gpurun_out/postproc_bench_k100.jsonl 2> gpurun_out/postproc_bench.err; echo "postproc rc=$?"; cat gpurun_out/postproc_bench_k100.jsonl; tail -3 gpurun_out/postproc_bench.err
timeout 300 python tools/bench_postproc.py --nq 30000 --k 1000 --cpu-sample 300 > gpurun_out/postproc_bench_k1000.jsonl 2>> gpurun_out/postproc_bench.err; cat gpurun_out/postproc_bench_k1000.jsonl
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_postproc.csv \
   python tools/bench_postproc.py --cpu-sample 10 > gpurun_out/ncu_postproc.log 2>&1; echo "ncu postproc rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:rerank -c 1 -f -o gpurun_out/prof_rerank_bf16 \
   python tools/prof_k.py 4000000 16384 1000 bf16 > gpurun_out/ncu_rerank.log 2>&1; echo "ncu rerank rc=$?"
