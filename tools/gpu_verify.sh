#!/bin/bash
# verification of HEAD on one B200: GPU tests, smoke, default bench (both arms), launch list of the C4 bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench_c4_default.json 2> gpurun_out/bench_c4_default.err; echo "bench rc=$?"
cut -c1-2500 gpurun_out/bench_c4_default.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c4_reference.json 2>> gpurun_out/bench_c4_default.err; echo "ref rc=$?"
cut -c1-800 gpurun_out/bench_c4_reference.json
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c4_full.csv $CMD > gpurun_out/ncu_launches_c4.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/ncu_launches_c4.log | cut -c1-600
