#!/usr/bin/env python3
"""GPU experiment: GEMM-kernel throughput for runtime variants on one fixed problem (same GPU,
interleaved so that thermal drift averages out).  Prints one line per (variant, repeat)."""
import json, sys, time
sys.path.insert(0, "knn-for-homology_b200")
import torch
import knn_b200

nb, nq, d, k = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000, 16384, 1024, 100
dev = torch.device("cuda:0")
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
variants = [dict(cta_group=1, l2_hints=0), dict(cta_group=1, l2_hints=1), dict(cta_group=2, l2_hints=0), dict(cta_group=2, l2_hints=1)]
idx.set_param("profile", 1)
for rep in range(3):
    for v in variants:
        for name, val in v.items():
            idx.set_param(name, val)
        idx.search(xq, k)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); gm = 0.0
        for _ in range(6):
            idx.search(xq, k); gm += idx.stat("gemm_ms")
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps(dict(v, rep=rep, qps=6 * nq / dt, gemm_tflops=6 * 2.0 * nq * nb * d / (gm / 1e3) / 1e12,
                              gemm_ms=gm / 6, step_ms=dt / 6 * 1e3)), flush=True)
