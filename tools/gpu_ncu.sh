#!/bin/bash
# ncu evidence for the bench's dominant kernel: launch list + one full capture (1 GPU, small config)
mkdir -p gpurun_out
CMD="python bench.py --nb 1000000 --nq 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/ncu_plain2.json 2>> gpurun_out/ncu_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:gemm_filter -s 27 -c 9 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
