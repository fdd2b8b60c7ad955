#!/bin/bash
# first GPU contact: exact path, then tensor path, each under its own timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 180 \
  -k "not tensor and not scale and not overflow and not bf16 and not torch_device and not ties" \
  > gpurun_out/t_exact.log 2>&1; echo "exact rc=$?" >> gpurun_out/t_exact.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 180 \
  -k "tensor or scale or overflow or bf16 or torch_device or ties" \
  > gpurun_out/t_tensor.log 2>&1; echo "tensor rc=$?" >> gpurun_out/t_tensor.log
tail -5 gpurun_out/t_exact.log; tail -30 gpurun_out/t_tensor.log
