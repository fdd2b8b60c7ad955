#!/bin/bash
# launch lists of the k = 1000 configurations
mkdir -p gpurun_out
for cfg in "4000000 16384 1000 bf16" "4000000 16384 1000" "14433 14433 1000" "300000 32768 1000"; do
  tag=$(echo $cfg | tr ' ' '_')
  python tools/prof_k.py $cfg > gpurun_out/prof_k_$tag.log 2>&1 || { echo plain run failed $cfg; tail -5 gpurun_out/prof_k_$tag.log; continue; }
  cat gpurun_out/prof_k_$tag.log
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
     --log-file gpurun_out/launches_k_$tag.csv python tools/prof_k.py $cfg > gpurun_out/ncu_k_$tag.log 2>&1
  echo ncu rc=$?
done
