#!/usr/bin/env python3
"""Wall-clock of the reference's driver-level flows on host (pageable numpy) arrays - what a user of the
reference's scripts experiences after switching: uploads, normalisation, index build, search, download included.
  cath.search.search              14,433 x 1024 all-vs-all, hits = 10, cosine and euclidean   (C2, cath/search.py:13-26)
  seqvec_search.main.faiss_search 300,000 x 1024 database, 30,000 queries, hits = 1000        (C3-like, main.py:22-50)
Both through knn_b200.drivers (one upload) and through the step-by-step faiss-style API over numpy."""
import json
import sys
import time

sys.path.insert(0, "knn-for-homology_b200")
sys.path.insert(0, ".")
import numpy as np
import torch

import knn_b200
from knn_b200 import drivers

rng = np.random.default_rng(0)


def clustered(n, d, c):
    cent = rng.standard_normal((c, d)).astype(np.float32)
    return cent[rng.integers(0, c, n)] + 0.45 * rng.standard_normal((n, d)).astype(np.float32)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


x = clustered(14433, 1024, 5125)
for name, metric in [("cosine", knn_b200.METRIC_INNER_PRODUCT), ("euclidean", knn_b200.METRIC_L2)]:
    t_drv, (hits, scores) = timed(lambda: drivers.search(x, hits=10, metric=metric))

    def stepwise():
        e = x.copy()
        if metric == knn_b200.METRIC_INNER_PRODUCT:
            knn_b200.normalize_L2(e)
        index = knn_b200.IndexFlat(e.shape[1], metric)
        index.add(e)
        s, r = index.search(e, 11)
        return r[:, 1:], s[:, 1:]

    t_step, (hits2, scores2) = timed(stepwise)
    print(json.dumps(dict(flow="cath.search.search", rows=14433, hits=10, metric=name, driver_ms=round(t_drv * 1e3, 1),
                          stepwise_numpy_api_ms=round(t_step * 1e3, 1), identical=bool(np.array_equal(hits, hits2) and np.array_equal(scores, scores2)))), flush=True)

hay = clustered(300_000, 1024, 3000)
qry = clustered(30_000, 1024, 3000)
drivers.faiss_search(hay[:70000].copy(), qry[:2000].copy(), hits=1000)  # warm-up: pinned bounce buffers, workspaces
hay_in, qry_in = hay.copy(), qry.copy()  # the copies are not part of the flow
torch.cuda.synchronize()
t0 = time.perf_counter()
ids, sc, search_s = drivers.faiss_search(hay_in, qry_in, hits=1000)
torch.cuda.synchronize()
total = time.perf_counter() - t0
print(json.dumps(dict(flow="seqvec_search.main.faiss_search", database_rows=300_000, queries=30_000, hits=1000, total_s=round(total, 3),
                      search_s_as_the_driver_reports_it=round(search_s, 3), result_bytes=int(ids.nbytes + sc.nbytes))), flush=True)
