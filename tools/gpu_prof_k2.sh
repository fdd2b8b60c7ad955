#!/bin/bash
# launch lists of the rescoring-heavy stand-ins with the automatic operand format (fp16)
mkdir -p gpurun_out
for cfg in "14433 14433 1000" "300000 32768 1000"; do
  tag=$(echo $cfg | tr ' ' '_')
  python tools/prof_k.py $cfg > gpurun_out/prof_k_$tag.log 2>&1 || { echo plain run failed $cfg; tail -5 gpurun_out/prof_k_$tag.log; continue; }
  cat gpurun_out/prof_k_$tag.log
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
     --log-file gpurun_out/launches_k_${tag}_fp16.csv python tools/prof_k.py $cfg > gpurun_out/ncu_k_$tag.log 2>&1
  echo ncu rc=$?
done
