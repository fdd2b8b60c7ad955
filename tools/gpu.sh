#!/bin/bash
# One parameterised GPU-box script (replaces the one-shot gpu_*.sh of round 1).  Run through gpurun from the repo root:
#   gpurun --timeout 1500 -- 'bash tools/gpu.sh <tag> <step> [<step> ...]'
# Every step writes into gpurun_out/<tag>_* and prints a short tail.  Steps:
#   micro            co-residency micro-test (tools/micro/overlap_test)
#   tests            pytest -m gpu (all GPU tests)
#   tests:<expr>     pytest -m gpu -k <expr>
#   smoke            __graft_entry__.smoke()
#   bench[:args]     python bench.py <args>              (args with '+' for spaces, e.g. bench:--steps+5+--warmup+3)
#   mbench:N[:args]  torchrun --nproc-per-node N bench.py --gpus N <args>
#   ref[:args]       bench.py --impl reference <args>
#   launches[:args]  ncu launch list (gpu__time_duration) of bench.py <args>
#   ncu:<kernel-regex>[:args]  one ncu --set full capture of the first launches matching the regex
#   py:<script>[:args]  python <script> <args>
#   prof:<name>:<kernel-regex>:<script+args>   ncu --set full --profile-from-start off of the launches matching the regex
#                        in a script that brackets ONE search with cudaProfilerStart/Stop (tools/prof_k.py, prof_small.py)
#   proflist:<name>:<script+args>              the launch list (gpu__time_duration) of the same bracketed region
tag=$1; shift
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for step in "$@"; do
  name=${step%%:*}; rest=""; [[ "$step" == *:* ]] && rest=${step#*:}
  args=${rest//+/ }
  echo "=== [$tag] $step"
  case $name in
    micro)
      timeout 120 tools/micro/overlap_test > gpurun_out/${tag}_micro.txt 2>&1; echo "rc=$?"; cat gpurun_out/${tag}_micro.txt ;;
    tests)
      if [ -n "$rest" ]; then timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x -k "$args" > gpurun_out/${tag}_tests.txt 2>&1
      else timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${tag}_tests.txt 2>&1; fi
      echo "rc=$?"; tail -15 gpurun_out/${tag}_tests.txt ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 ;;
    bench)
      f=gpurun_out/${tag}_bench$(echo "$rest" | tr -c 'a-zA-Z0-9\n' '_').json
      timeout 900 python bench.py $args > $f 2> ${f%.json}.err; echo "rc=$?"; cut -c1-600 $f; tail -3 ${f%.json}.err ;;
    mbench)
      n=${rest%%:*}; margs=""; [[ "$rest" == *:* ]] && margs=${rest#*:}; margs=${margs//+/ }
      f=gpurun_out/${tag}_bench_${n}gpu$(echo "$margs" | tr -c 'a-zA-Z0-9\n' '_').json
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $n $margs > $f 2> ${f%.json}.err; echo "rc=$?"; cut -c1-600 $f; tail -3 ${f%.json}.err ;;
    ref)
      f=gpurun_out/${tag}_reference.json
      timeout 600 python bench.py --impl reference $args > $f 2> ${f%.json}.err; echo "rc=$?"; cut -c1-300 $f ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${tag}_launches.csv \
        python bench.py $args > gpurun_out/${tag}_launches.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/${tag}_launches.log | cut -c1-200 ;;
    ncu)
      kre=${rest%%:*}; nargs=""; [[ "$rest" == *:* ]] && nargs=${rest#*:}; nargs=${nargs//+/ }
      timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$kre" -s 6 -c 3 -o gpurun_out/${tag}_ncu_$(echo "$kre" | tr -c 'a-zA-Z0-9\n' '_') -f \
        python bench.py $nargs > gpurun_out/${tag}_ncu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/${tag}_ncu.log | cut -c1-200 ;;
    prof)
      pname=${rest%%:*}; r2=${rest#*:}; kre=${r2%%:*}; pargs=${r2#*:}; pargs=${pargs//+/ }
      timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$kre" -c 12 \
        -o gpurun_out/${tag}_ncu_${pname} -f python $pargs > gpurun_out/${tag}_ncu_${pname}.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/${tag}_ncu_${pname}.log | cut -c1-200 ;;
    proflist)
      pname=${rest%%:*}; pargs=${rest#*:}; pargs=${pargs//+/ }
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/${tag}_launches_${pname}.csv python $pargs > gpurun_out/${tag}_launches_${pname}.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/${tag}_launches_${pname}.log | cut -c1-200 ;;
    py)
      script=${rest%%:*}; pargs=""; [[ "$rest" == *:* ]] && pargs=${rest#*:}; pargs=${pargs//+/ }
      f=gpurun_out/${tag}_$(basename $script .py).txt
      timeout 900 python $script $pargs > $f 2>&1; echo "rc=$?"; tail -25 $f ;;
    *) echo "unknown step $step" ;;
  esac
done
nvidia-smi --query-gpu=index,clocks.sm,power.draw,temperature.gpu --format=csv,noheader | head -8
