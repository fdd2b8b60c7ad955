"""CPU flat-search baseline for bench.py (TEST/BENCH INFRASTRUCTURE ONLY - never on the product path).

The reference runs this path in faiss-cpu 1.7.2 (not installable offline; `import faiss` is
still tried first so that a real faiss is used if one ever appears).  Otherwise this is the
oracle's algorithm (oracle/flat_oracle.py) with the heavy lifting moved to multithreaded BLAS,
i.e. what faiss does: blocked fp32 sgemm (query block 4096) + per-block top-k + merge, on all
host cores.  Flat search costs exactly nq*N*d multiply-adds, so a bounded sample is timed and
the queries/s figure is scaled by N_sample/N to the full database size.
"""
from __future__ import annotations

import os
import time

import numpy as np


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _real_faiss():
    try:
        import faiss  # noqa

        if hasattr(faiss, "omp_get_max_threads") and hasattr(faiss, "IndexFlat"):
            return faiss
    except Exception:
        pass
    return None


def flat_search_blas(xq: np.ndarray, xb: np.ndarray, k: int, threads: int):
    """Blocked sgemm + top-k on `threads` cores (torch MKL/OpenBLAS).  IP metric."""
    import torch

    torch.set_num_threads(threads)
    tq = torch.from_numpy(xq)
    tb = torch.from_numpy(xb)
    nq, nb = tq.shape[0], tb.shape[0]
    D = torch.empty((nq, k), dtype=torch.float32)
    I = torch.empty((nq, k), dtype=torch.int64)
    qbs, dbs = 4096, 1024 * 64
    for i0 in range(0, nq, qbs):
        i1 = min(nq, i0 + qbs)
        best_v = best_i = None
        for j0 in range(0, nb, dbs):
            j1 = min(nb, j0 + dbs)
            ip = tq[i0:i1] @ tb[j0:j1].T
            v, idx = torch.topk(ip, min(k, j1 - j0), dim=1)
            idx += j0
            if best_v is None:
                best_v, best_i = v, idx
            else:
                cv = torch.cat([best_v, v], dim=1)
                ci = torch.cat([best_i, idx], dim=1)
                best_v, sel = torch.topk(cv, min(k, cv.shape[1]), dim=1)
                best_i = torch.gather(ci, 1, sel)
        D[i0:i1, : best_v.shape[1]] = best_v
        I[i0:i1, : best_i.shape[1]] = best_i
    return D.numpy(), I.numpy()


def time_sample(n_full: int, d: int, k: int, *, n_sample: int, nq_sample: int, seed: int = 99):
    """Times one flat search on a bounded sample; returns a dict for bench.py's cpu_baseline."""
    cores = host_cores()
    rng = np.random.default_rng(seed)
    n_sample = min(n_sample, n_full)
    xb = rng.standard_normal((n_sample, d), dtype=np.float32)
    xq = rng.standard_normal((nq_sample, d), dtype=np.float32)
    xb /= np.linalg.norm(xb, axis=1, keepdims=True)
    xq /= np.linalg.norm(xq, axis=1, keepdims=True)
    faiss = _real_faiss()
    t0 = time.perf_counter()
    if faiss is not None:
        index = faiss.IndexFlat(d, faiss.METRIC_INNER_PRODUCT)
        index.add(xb)
        t0 = time.perf_counter()
        index.search(xq, k)
        kind, cores_used = "reference", faiss.omp_get_max_threads()
    else:
        flat_search_blas(xq[:8], xb[: max(k, 1024)], k, cores)  # import + thread-pool warm-up, untimed
        t0 = time.perf_counter()
        flat_search_blas(xq, xb, k, cores)
        kind, cores_used = "port", cores
    dt = time.perf_counter() - t0
    qps_sample = nq_sample / dt
    scale = n_sample / n_full
    return {
        "value": qps_sample * scale,
        "unit": "queries/s",
        "cores": int(cores_used),
        "kind": kind,
        "seconds": dt,
        "sample": f"{nq_sample} queries x {n_sample} rows x {d}-d, k={k}, fp32 sgemm+top-k on {cores_used} threads; "
                  f"queries/s scaled by {scale:.6g} (=N_sample/N) to N={n_full}",
    }
