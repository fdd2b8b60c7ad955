#!/usr/bin/env python3
"""GPU experiment: small query batches.  For nq in {1, 8, 16, 64, 256} times the exact fp32 scan
path (streams the fp32 rows once per 8 queries) against the tensor path (streams the bf16 rows once
through the tcgen05 kernel, cta_group 1 or 2) and prints achieved HBM GB/s on the algorithmic bytes."""
import json
import sys
import time

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
xq_all = torch.randn(256, d, device=dev, generator=g)
knn_b200.normalize_L2(xq_all)
ref = {}
for nq in (1, 8, 16, 64, 256):
    xq = xq_all[:nq].contiguous()
    for name, params in [("exact", dict(path=1)), ("tensor_cg1", dict(path=2, cta_group=1)), ("tensor_cg2", dict(path=2, cta_group=2))]:
        for p, v in params.items():
            idx.set_param(p, v)
        D, I = idx.search(xq, k)
        torch.cuda.synchronize()
        if name == "exact":
            ref[nq] = (D, I)
        same = bool(torch.equal(ref[nq][1], I) and torch.equal(ref[nq][0], D))
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            idx.search(xq, k)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / reps * 1e3
        passes = (nq + 7) // 8 if name == "exact" else 1
        alg_bytes = nb * d * (4 * passes if name == "exact" else 2)
        print(json.dumps(dict(nq=nq, path=name, ms=round(ms, 3), qps=round(nq / ms * 1e3, 1), identical_to_exact=same,
                              db_bytes_streamed_GB=round(alg_bytes / 1e9, 2), achieved_GBps=round(alg_bytes / ms / 1e6, 1))), flush=True)
