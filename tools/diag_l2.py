#!/usr/bin/env python3
"""Diagnostic: one all-vs-all search of clustered 14,433 x 1024 rows (the cath.search flow) per metric, with the engine's
own statistics - where does the time of the euclidean pass go?"""
import json
import sys
import time

sys.path.insert(0, "knn-for-homology_b200")
import numpy as np
import torch

import knn_b200

rng = np.random.default_rng(0)
cent = rng.standard_normal((5125, 1024)).astype(np.float32)
x = cent[rng.integers(0, 5125, 14433)] + 0.45 * rng.standard_normal((14433, 1024)).astype(np.float32)
for name, metric in [("cosine", 0), ("euclidean", 1)]:
    xd = torch.from_numpy(x).cuda()
    if metric == 0:
        knn_b200.normalize_L2(xd)
    for params in ({}, {"overlap_finish": 0}, {"shadow_fmt": 1}, {"path": 1}):
        idx = knn_b200.IndexFlat(1024, metric)
        for k_, v in params.items():
            idx.set_param(k_, v)
        idx.add(xd)
        idx.search(xd, 11)
        torch.cuda.synchronize()
        idx.set_param("profile", 1)
        t0 = time.perf_counter()
        D, I = idx.search(xd, 11)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        print(json.dumps(dict(metric=name, params=params, ms=round(ms, 2), path=idx.stat("path"), launches=idx.stat("launches"),
                              gemm_ms=round(idx.stat("gemm_ms"), 2), rerank_ms=round(idx.stat("rerank_ms"), 2),
                              overflow_queries=idx.stat("overflow_queries"), shadow_fmt=idx.stat("shadow_fmt"),
                              conversions=idx.stat("shadow_conversions"))), flush=True)
        del idx
