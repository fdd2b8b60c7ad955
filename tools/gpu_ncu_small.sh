#!/bin/bash
# ncu evidence for the small-batch kernels: few-queries variant at nq = 8, main kernel at nq = 256 (largest panel of each)
mkdir -p gpurun_out
python tools/prof_small.py 4000000 8 > gpurun_out/prof_small_plain.log 2>&1 && python tools/prof_small.py 4000000 256 >> gpurun_out/prof_small_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_ -f -o gpurun_out/prof_stream_nq8 python tools/prof_small.py 4000000 8 > gpurun_out/ncu_stream.log 2>&1
echo "ncu stream rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_ -f -o gpurun_out/prof_main_nq256 python tools/prof_small.py 4000000 256 > gpurun_out/ncu_main256.log 2>&1
echo "ncu main rc=$?"
ls -la gpurun_out/*.ncu-rep
