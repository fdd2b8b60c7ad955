#!/bin/bash
# 16-bit operand formats: GPU tests, C4 bench with bf16 operands of 7/6/5/4 mantissa bits on the same box, stand-ins with the automatic choice
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -8
for mb in 7 6 5 4 7; do
  timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --shadow-fmt 1 --mantissa-bits $mb > gpurun_out/bench_c4_mb$mb.json 2> gpurun_out/bench_c4_mb$mb.err; echo "bench mbits=$mb rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c4_mb$mb.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ["value","ms_per_step","dtype","parity_spot_check"]}, d["roofline"]["achieved"], d["roofline"]["kernel_ms_per_step"], d["clocks"])
PY
done
timeout 300 python tools/exp_configs.py C2 C3 C5 > gpurun_out/configs_auto.jsonl 2> gpurun_out/configs_auto.err; echo "configs auto rc=$?"; cut -c1-330 gpurun_out/configs_auto.jsonl
