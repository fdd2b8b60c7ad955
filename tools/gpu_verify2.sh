#!/bin/bash
# GPU tests, default bench, DRAM traffic of every GEMM launch of one C4 step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench_c4_default.json 2> gpurun_out/bench_c4_default.err; echo "bench rc=$?"
cut -c1-2600 gpurun_out/bench_c4_default.json
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_filter -s 98 -c 98 --csv --log-file gpurun_out/gemm_traffic_c4.csv $CMD > gpurun_out/ncu_traffic_c4.log 2>&1
echo "traffic rc=$?"; tail -3 gpurun_out/gemm_traffic_c4.csv | cut -c1-300
