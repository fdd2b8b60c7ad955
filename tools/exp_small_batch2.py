#!/usr/bin/env python3
"""GPU experiment: the tensor path as the bandwidth kernel of small query batches - the main kernel (6-stage ring
for few query tiles) against the few-queries variant.  Prints whole-call ms, the GEMM launches' ms and GB/s on the
16-bit database bytes."""
import json
import sys

sys.path.insert(0, "knn-for-homology_b200")
import torch

import knn_b200

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
d, k = 1024, 100
dev = torch.device("cuda:0")
idx = knn_b200.IndexFlat(d, 0)
idx.reserve(nb)
for blk in range(0, nb, 65536):
    g = torch.Generator(device=dev).manual_seed(1234 + blk)
    rows = torch.randn(min(65536, nb - blk), d, device=dev, generator=g)
    knn_b200.normalize_L2(rows)
    idx.add(rows)
g = torch.Generator(device=dev).manual_seed(4321)
xq_all = torch.randn(1024, d, device=dev, generator=g)
knn_b200.normalize_L2(xq_all)
idx.set_param("path", 1)
ref = {nq: idx.search(xq_all[:nq].contiguous(), k) for nq in (1, 64, 256, 1024)}
idx.set_param("path", 2)
for nq in (1, 16, 32, 64, 256, 1024):
    xq = xq_all[:nq].contiguous()
    if nq not in ref:
        idx.set_param("path", 1)
        ref[nq] = idx.search(xq, k)
        idx.set_param("path", 2)
    for stream in (0, 1):
        if stream and nq > 64:
            continue
        idx.set_param("stream_kernel", stream)
        idx.set_param("profile", 0)
        D, I = idx.search(xq, k)
        same = bool(torch.equal(ref[nq][1], I) and torch.equal(ref[nq][0], D))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            idx.search(xq, k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        idx.set_param("profile", 1)
        idx.search(xq, k)
        gemm_ms = idx.stat("gemm_ms")
        print(json.dumps(dict(nq=nq, kernel="few-queries (rows as M, queries resident)" if stream else "main (6-stage ring)", ms=round(ms, 3),
                              gemm_ms=round(gemm_ms, 3), rerank_ms=round(idx.stat("rerank_ms"), 3), launches=int(idx.stat("launches")),
                              GBps_call=round(nb * d * 2 / ms / 1e6, 1), GBps_gemm=round(nb * d * 2 / gemm_ms / 1e6, 1),
                              identical_to_exact=same, shadow_fmt=int(idx.stat("shadow_fmt")))), flush=True)
