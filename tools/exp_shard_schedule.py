#!/usr/bin/env python3
"""GPU experiment: panel growth ratio at the size of one shard of an 8-GPU C4 run (1.25M x 1024 rows, 100k queries,
k = 100) - fewer, larger panels trade tighten launches for candidates.  One GPU as a proxy for a rank."""
import json
import sys

sys.path.insert(0, "knn-for-homology_b200")
import torch

import knn_b200

N, NQ, K = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000, 100_000, 100
dev = torch.device("cuda:0")
idx = knn_b200.IndexFlat(1024, 0)
idx.reserve(N)
for blk in range(0, N, 65536):
    g = torch.Generator(device=dev).manual_seed(1234 + blk)
    x = torch.randn(min(65536, N - blk), 1024, device=dev, generator=g)
    knn_b200.normalize_L2(x)
    idx.add(x)
g = torch.Generator(device=dev).manual_seed(4321)
xq = torch.randn(NQ, 1024, device=dev, generator=g)
knn_b200.normalize_L2(xq)
ref = None
for rep in range(2):
    for ratio in (0, 3, 4, 8):
        idx.set_param("panel_ratio", ratio)
        idx.set_param("profile", 0)
        for _ in range(2):
            D, I = idx.search(xq, K)
        torch.cuda.synchronize()
        if ref is None:
            ref = (D.clone(), I.clone())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            idx.search(xq, K)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 4
        idx.set_param("profile", 1)
        idx.search(xq, K)
        print(json.dumps(dict(rows=N, panel_ratio=ratio or "auto (2)", ms=round(ms, 2), gemm_ms=round(idx.stat("gemm_ms"), 2),
                              rerank_ms=round(idx.stat("rerank_ms"), 2), gemm_launches=int(idx.stat("gemm_launches")),
                              launches=int(idx.stat("launches")), identical=bool(torch.equal(ref[1], I) and torch.equal(ref[0], D)))), flush=True)
