#!/usr/bin/env python3
"""GPU experiment: search / GEMM-kernel throughput for runtime variants on one fixed problem
(same GPU, interleaved so that thermal drift averages out), bracketed by a cuBLAS bf16 GEMM of
the same K as the at-this-GPU, at-this-power reference.

    python tools/exp_gemm.py [nb] ['[{"cta_group":2,...}, ...]']
"""
import json
import sys
import time

sys.path.insert(0, "knn-for-homology_b200")
import torch

import knn_b200

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
nq, d, k = (int(sys.argv[3]) if len(sys.argv) > 3 else 32768), 1024, 100
dev = torch.device("cuda:0")
variants = [dict(cta_group=2), dict(cta_group=1)]  # debug_skip_epilogue needs a library built with -DKNN_EXPERIMENTS
if len(sys.argv) > 2:
    variants = json.loads(sys.argv[2])


def cublas_reference():
    """torch.matmul (cuBLAS) on a GEMM of the same K, ~3 s sustained."""
    a = torch.randn(16384, d, device=dev, dtype=torch.bfloat16)
    b = torch.randn(262144, d, device=dev, dtype=torch.bfloat16)
    c = torch.empty(16384, 262144, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        torch.matmul(a, b.T, out=c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 300
    e0.record()
    for _ in range(n):
        torch.matmul(a, b.T, out=c)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps(dict(cublas_tflops=n * 2.0 * 16384 * 262144 * d / (ms / 1e3) / 1e12, seconds=ms / 1e3)), flush=True)


idx = knn_b200.IndexFlat(d, 0)
idx.reserve(nb)
for blk in range(0, nb, 65536):
    g = torch.Generator(device=dev).manual_seed(1234 + blk)
    rows = torch.randn(min(65536, nb - blk), d, device=dev, generator=g)
    knn_b200.normalize_L2(rows)
    idx.add(rows)
g = torch.Generator(device=dev).manual_seed(4321)
xq = torch.randn(nq, d, device=dev, generator=g)
knn_b200.normalize_L2(xq)

cublas_reference()
STEPS = 6
DEFAULTS = dict(cta_group=2, gemm_stages=0, panel_ratio=0, query_batch=16384)
for rep in range(3):
    for v in variants:
        for name, val in dict(DEFAULTS, **v).items():
            idx.set_param(name, val)
        prof = True
        idx.set_param("profile", 1 if prof else 0)
        idx.search(xq, k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gm = 0.0
        for _ in range(STEPS):
            idx.search(xq, k)
            gm += idx.stat("gemm_ms") if prof else 0.0
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        flop = STEPS * 2.0 * nq * nb * d
        print(json.dumps(dict(v, rep=rep, qps=STEPS * nq / dt, gemm_tflops=(flop / (gm / 1e3) / 1e12) if gm else None,
                              search_tflops=flop / dt / 1e12, gemm_ms=gm / STEPS, step_ms=dt / STEPS * 1e3)), flush=True)
cublas_reference()
