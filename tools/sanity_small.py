#!/usr/bin/env python3
"""Tiny searches through every kernel family and entry point (host, device, batch-wise two-phase), each compared with
the exact scan.  (Written for compute-sanitizer, which is closed on this pool; useful as a 5-second smoke of all variants.)"""
import sys

sys.path.insert(0, "knn-for-homology_b200")
import numpy as np
import torch

import knn_b200

rng = np.random.default_rng(0)
xb = rng.standard_normal((20011, 256)).astype(np.float32)
for metric in (0, 1):
    exact = knn_b200.IndexFlat(256, metric)
    exact.set_param("path", 1)
    exact.add(xb)
    for nq, k, params in [(1, 10, {}), (40, 100, {}), (100, 10, {}), (128, 300, {}), (200, 10, {"stream_quad": 1}), (256, 50, {"stream_quad": 1}),
                          (700, 20, {"query_batch": 256}), (700, 300, {"query_batch": 256, "l2_blocked_rerank": 1}), (300, 5, {"cta_group": 1})]:
        xq = rng.standard_normal((nq, 256)).astype(np.float32)
        idx = knn_b200.IndexFlat(256, metric)
        idx.set_param("path", 2)
        for name, v in params.items():
            idx.set_param(name, v)
        idx.add(xb[:7000])
        idx.add(xb[7000:])
        D, I = idx.search(xq, k)                       # host path (pipelined)
        Dd, Id = idx.search(torch.from_numpy(xq).cuda(), k)  # device path
        D1, I1 = exact.search(xq, k)
        assert np.array_equal(I, I1) and np.array_equal(D, D1), (metric, nq, k, params)
        assert np.array_equal(Id.cpu().numpy(), I1) and np.array_equal(Dd.cpu().numpy(), D1), (metric, nq, k, params)
        # two-phase, batch-wise
        tq = torch.from_numpy(xq).cuda()
        nb, rows = idx.search_begin(tq, k)
        bounds = torch.zeros((nb, 2, rows), dtype=torch.float32, device="cuda")
        D2 = torch.empty((nq, k), dtype=torch.float32, device="cuda")
        I2 = torch.empty((nq, k), dtype=torch.int64, device="cuda")
        for b in range(nb):
            idx.search_filter_batch(b, k, bounds)  # one shard: j = ceil(k / 1)
            idx.search_finish_batch(b, bounds, D2, I2)
        idx.search_end(D2, I2)
        assert np.array_equal(I2.cpu().numpy(), I1) and np.array_equal(D2.cpu().numpy(), D1), (metric, nq, k, "two-phase")
    print("metric", metric, "ok", flush=True)
x = torch.randn(1000, 100, device="cuda")
knn_b200.normalize_L2(x)
print("done", flush=True)
