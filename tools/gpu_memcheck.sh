#!/bin/bash
# compute-sanitizer memcheck (one tool per call) over small searches of every kernel family
mkdir -p gpurun_out
timeout 200 python tools/sanitize_cases.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -3 gpurun_out/sanitize_plain.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck.log python tools/sanitize_cases.py > gpurun_out/sanitize_run.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/sanitize_run.log; tail -12 gpurun_out/memcheck.log
