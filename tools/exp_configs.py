#!/usr/bin/env python3
"""GPU run of the other BASELINE.json configurations at single-GPU scale (stand-ins with synthetic
data of the same shape), with a sampled parity check against the exact fp32
path of the same engine.  One JSON line per configuration.

  C2  CATH20-like: 14,433 x 1024 all-vs-all, k = 11 (reference) and k = 1000 (BASELINE.json), IP and L2
  C3  Pfam20-like: 300,000 x 1024 all-vs-all, k = 1000, clustered rows (families)
  C5  UniRef90-like shard: bf16-only storage, 4M x 1024 rows (1/25 of a 100M database), 32,768 queries, k = 1000
"""
import json
import sys
import time

sys.path.insert(0, "knn-for-homology_b200")
import torch

import knn_b200

dev = torch.device("cuda:0")
# 16-bit operand format: "fmt=1" forces bf16 shadow rows, "fmt=2" fp16 (0 = automatic); "mbits=5" keeps 5 mantissa bits in bf16
OPTS = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
SHADOW_FMT, MBITS = int(OPTS.get("fmt", 0)), int(OPTS.get("mbits", 0))


def rows(n, d, seed, clustered=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(n, d, device=dev, generator=g)
    if clustered:
        c = torch.randn(clustered, d, device=dev, generator=g)
        x = c[torch.randint(0, clustered, (n,), device=dev, generator=g)] + 0.45 * x
    return x


def run(name, xb, xq, k, metric, normalize=True, bf16_storage=False, reps=3):
    if normalize:
        knn_b200.normalize_L2(xb)
        if xq.data_ptr() != xb.data_ptr():
            knn_b200.normalize_L2(xq)
    t0 = time.perf_counter()
    idx = knn_b200.IndexFlat(xb.shape[1], metric, bf16_storage=bf16_storage)
    if SHADOW_FMT and not bf16_storage:
        idx.set_param("shadow_fmt", SHADOW_FMT)
    idx.set_param("mantissa_bits", MBITS)
    idx.reserve(xb.shape[0])
    for i in range(0, xb.shape[0], 1 << 20):
        idx.add(xb[i:i + (1 << 20)])
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    D, I = idx.search(xq, k)
    torch.cuda.synchronize()
    path, overflow = int(idx.stat("path")), int(idx.stat("overflow_batches"))
    t0 = time.perf_counter()
    for _ in range(reps):
        D, I = idx.search(xq, k)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    sample = torch.arange(0, xq.shape[0], max(1, xq.shape[0] // 48), device=dev)[:48]
    idx.set_param("path", 1)
    D1, I1 = idx.search(xq[sample].contiguous(), k)
    same = bool(torch.equal(I[sample], I1) and torch.equal(D[sample], D1))
    flop = 2.0 * xq.shape[0] * xb.shape[0] * xb.shape[1]
    print(json.dumps(dict(config=name, rows=xb.shape[0], queries=xq.shape[0], k=k, metric="IP" if metric == 0 else "L2",
                          ms=round(ms, 2), qps=round(xq.shape[0] / ms * 1e3, 1), tflops_equiv=round(flop / ms / 1e9, 1),
                          path=path, overflow_batches=overflow, shadow_fmt=int(idx.stat("shadow_fmt")), mantissa_bits=int(idx.stat("mantissa_bits")), identical_to_exact_path_on_sample=same,
                          build_s=round(build_s, 3))), flush=True)
    del idx


which = [a for a in sys.argv[1:] if "=" not in a] or ["C2", "C3", "C5"]
if "C2" in which:
    x = rows(14433, 1024, 1, clustered=5125)
    run("C2 k=11 cosine", x.clone(), None or x.clone(), 11, 0)
    xb = x.clone(); run("C2 k=11 cosine all-vs-all", xb, xb, 11, 0)
    xb = x.clone(); run("C2 k=11 euclidean all-vs-all", xb, xb, 11, 1, normalize=False)
    xb = x.clone(); run("C2 k=1000 cosine all-vs-all", xb, xb, 1000, 0)
    del x, xb
if "C3" in which:
    xb = rows(300_000, 1024, 2, clustered=3000)
    run("C3 k=1000 cosine all-vs-all", xb, xb, 1000, 0, reps=1)
    del xb
if "C5" in which:
    xb = rows(4_000_000, 1024, 3)
    xq = rows(32768, 1024, 4)
    run("C5 shard (bf16 storage) k=1000", xb, xq, 1000, 0, bf16_storage=True, reps=1)
    run("C5 shard (bf16 storage) k=100", xb, xq, 100, 0, normalize=False, bf16_storage=True, reps=1)
