#!/bin/bash
# tests, then cta_group 1 vs 2 on the small and the full config
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/t_gpu.log
tail -15 gpurun_out/t_gpu.log
for cg in 1 2; do
  timeout 600 python bench.py --cta-group $cg --nb 1000000 --nq 16384 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_small_cg$cg.json 2> gpurun_out/bench_small_cg$cg.err; echo "small cg$cg rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/bench_small_cg$cg.json'));print(d['value'],d['ms_per_step'],d['roofline']['achieved'],d['roofline']['kernel_ms_per_step'],d['parity_spot_check'],d['clocks'])"
  tail -3 gpurun_out/bench_small_cg$cg.err
done
for cg in 2 1; do
  timeout 900 python bench.py --cta-group $cg --no-cpu-baseline > gpurun_out/bench_c4_cg$cg.json 2> gpurun_out/bench_c4_cg$cg.err; echo "c4 cg$cg rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/bench_c4_cg$cg.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['achieved'],d['roofline']['kernel_ms_per_step'],d['parity_spot_check'],d['clocks'])"
  tail -3 gpurun_out/bench_c4_cg$cg.err
done
