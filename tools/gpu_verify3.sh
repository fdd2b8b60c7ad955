#!/bin/bash
# verification of HEAD on one B200: GPU tests, smoke, small-batch check, default bench (both arms)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 200 python tools/exp_small_batch.py > gpurun_out/small_v4.jsonl 2> gpurun_out/small_v4.err; echo "small rc=$?"; grep tensor_cg2 gpurun_out/small_v4.jsonl
timeout 600 python bench.py > gpurun_out/bench_c4_default.json 2> gpurun_out/bench_c4_default.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/bench_c4_default.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c4_reference.json 2>> gpurun_out/bench_c4_default.err; echo "ref rc=$?"
cut -c1-300 gpurun_out/bench_c4_reference.json
