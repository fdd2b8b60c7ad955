"""CPU oracle for the flat (exact) kNN path of konstin/knn-for-homology.

TEST INFRASTRUCTURE ONLY.  Nothing under ``knn-for-homology_b200/`` imports this
module; it is used by ``tests/``, by ``__graft_entry__.smoke()`` and by the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the checker and
the CPU baseline, never as the shipped path.

What it restates
----------------
The reference executes this path inside the third-party wheel ``faiss-cpu``
(pinned 1.7.2, /root/reference/poetry.lock:100-101), whose source is NOT in
/root/reference and which cannot be installed here (no network).  The semantics
are therefore restated from the reference's call sites

* ``faiss.normalize_L2(x)``            cath/search.py:19, pfam/proteins_search.py:22,
                                       seqvec_search/main.py:31,34
* ``faiss.IndexFlat(d, metric)``       cath/search.py:20, pfam/proteins_search.py:24,
                                       seqvec_search/main.py:35
* ``index.train / index.add``          pfam/proteins_search.py:35-37, seqvec_search/main.py:37-39
* ``index.search(xq, k) -> (D, I)``    cath/search.py:24, pfam/proteins_search.py:49,
                                       seqvec_search/main.py:45

and from faiss 1.7.2's published algorithm (upstream knowledge, see SURVEY.md
section 3.5): fp32 everywhere; ``x[i] *= 1/sqrt(sum x[i]^2)`` with zero rows left
untouched; blocked ``sgemm`` (query block 4096 x database block 1024) feeding a
per-query top-k; squared L2 computed as ``|x|^2 + |y|^2 - 2<x,y>`` clamped at 0;
results sorted best first; int64 labels; when ``k > ntotal`` the tail is padded
with label -1 and distance ``-FLT_MAX`` (IP) / ``+FLT_MAX`` (L2), the heap's
neutral element.

Parity status
-------------
PINNED (by the reference's own tests, reproduced in tests/test_oracle.py):
the IP neighbour ids and their order on ``test-data/small-random`` (k=5,
tests/test_main.py:17-18) and ``test-data/pfam-20-10`` (k=10,
tests/test_main.py:26-27) through the reference's unmodified
``faiss_search``/``evaluate_faiss`` (see tests/golden/make_golden.py).
UNPINNED (no reference test, no runnable faiss): the distance values ``D``, the
L2 metric, ``k > ntotal`` padding values, the order inside exact-score ties
(this oracle breaks ties by the lower id) and k >= 100.  For those the oracle is
the published algorithm restated, cross-checked against an fp64 evaluation and
against the plain-C heap restatement in ``flat_oracle.c``.
"""
from __future__ import annotations

import numpy as np

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1

FLT_MAX = np.float32(np.finfo(np.float32).max)

# faiss/utils/distances.cpp block sizes (upstream knowledge): distance_compute_blas_query_bs /
# distance_compute_blas_database_bs.
QUERY_BS = 4096
DATABASE_BS = 1024


def _check_matrix(x: np.ndarray, d: int | None = None) -> None:
    if not isinstance(x, np.ndarray) or x.ndim != 2:
        raise ValueError("expected a 2-d numpy array")
    if x.dtype != np.float32:
        raise TypeError("expected float32, got %s" % x.dtype)
    if not x.flags.c_contiguous:
        raise ValueError("expected a C-contiguous array")
    if d is not None and x.shape[1] != d:
        raise ValueError("expected %d columns, got %d" % (d, x.shape[1]))


def normalize_L2(x: np.ndarray) -> None:
    """In-place row normalisation (faiss.normalize_L2; call sites cath/search.py:19,
    seqvec_search/main.py:31,34).  fp32; rows with zero norm are left as they are."""
    _check_matrix(x)
    nr = np.einsum("ij,ij->i", x, x, dtype=np.float32)
    nz = nr > 0
    inv = np.ones_like(nr)
    inv[nz] = np.float32(1.0) / np.sqrt(nr[nz], dtype=np.float32)
    x *= inv[:, None]


def _topk_sorted(scores: np.ndarray, base: int, k: int, largest: bool):
    """Per-row best-k of a dense block.  Returns (vals, ids) sorted best first with ties
    broken by the lower id (stable sort on ids that are already ascending)."""
    key = -scores if largest else scores
    n = key.shape[1]
    kk = min(k, n)
    if kk < n:
        # argpartition then a stable sort on (key, id) restricted to the kept part; ties at the
        # k-th value are resolved explicitly so that the lower id wins.
        part = np.argpartition(key, kk - 1, axis=1)[:, :kk]
        kth = np.take_along_axis(key, part, axis=1).max(axis=1)
        out_ids = np.empty((key.shape[0], kk), dtype=np.int64)
        for r in range(key.shape[0]):
            row = key[r]
            strictly = np.flatnonzero(row < kth[r])
            ties = np.flatnonzero(row == kth[r])[: kk - strictly.size]
            cand = np.concatenate([strictly, ties])
            order = np.lexsort((cand, row[cand]))
            out_ids[r] = cand[order]
    else:
        out_ids = np.argsort(key, axis=1, kind="stable").astype(np.int64)
    vals = np.take_along_axis(scores, out_ids, axis=1)
    return vals, out_ids + base


def _merge(vals_a, ids_a, vals_b, ids_b, k: int, largest: bool):
    vals = np.concatenate([vals_a, vals_b], axis=1)
    ids = np.concatenate([ids_a, ids_b], axis=1)
    key = -vals if largest else vals
    out_v = np.empty((vals.shape[0], min(k, vals.shape[1])), dtype=np.float32)
    out_i = np.empty(out_v.shape, dtype=np.int64)
    for r in range(vals.shape[0]):
        order = np.lexsort((ids[r], key[r]))[: out_v.shape[1]]
        out_v[r] = vals[r, order]
        out_i[r] = ids[r, order]
    return out_v, out_i


def knn_flat(xq: np.ndarray, xb: np.ndarray, k: int, metric: int, *, want: int | None = None):
    """Blocked fp32 flat search (faiss knn_inner_product / knn_L2sqr, BLAS path).

    Returns ``(D float32 (nq, k), I int64 (nq, k))`` sorted best first, padded with
    (-/+FLT_MAX, -1) when ``k > ntotal``.  ``want`` lets the parity checker ask for k+1.
    """
    nq, d = xq.shape
    nb = xb.shape[0]
    kk = k if want is None else want
    largest = metric == METRIC_INNER_PRODUCT
    pad_val = -FLT_MAX if largest else FLT_MAX
    D = np.full((nq, kk), pad_val, dtype=np.float32)
    I = np.full((nq, kk), -1, dtype=np.int64)
    if nq == 0 or nb == 0:
        return D, I
    if metric == METRIC_L2:
        xq_n = np.einsum("ij,ij->i", xq, xq, dtype=np.float32)
        xb_n = np.einsum("ij,ij->i", xb, xb, dtype=np.float32)
    for i0 in range(0, nq, QUERY_BS):
        i1 = min(nq, i0 + QUERY_BS)
        best_v = np.empty((i1 - i0, 0), dtype=np.float32)
        best_i = np.empty((i1 - i0, 0), dtype=np.int64)
        # database blocks are visited in multiples of DATABASE_BS; a larger stride only
        # changes how often the running top-k is merged, not the arithmetic of a score.
        step = DATABASE_BS * 64
        for j0 in range(0, nb, step):
            j1 = min(nb, j0 + step)
            ip = xq[i0:i1] @ xb[j0:j1].T  # sgemm, fp32
            if metric == METRIC_L2:
                dis = xq_n[i0:i1, None] + xb_n[None, j0:j1] - np.float32(2.0) * ip
                np.maximum(dis, np.float32(0.0), out=dis)  # faiss: "if (dis < 0) dis = 0"
                blk = dis
            else:
                blk = ip
            v, ids = _topk_sorted(blk, j0, kk, largest)
            best_v, best_i = _merge(best_v, best_i, v, ids, kk, largest)
        m = best_v.shape[1]
        D[i0:i1, :m] = best_v
        I[i0:i1, :m] = best_i
    return D, I


class IndexFlat:
    """faiss.IndexFlat(d, metric) restated (cath/search.py:20, pfam/proteins_search.py:24)."""

    def __init__(self, d: int, metric: int = METRIC_L2):
        if metric not in (METRIC_INNER_PRODUCT, METRIC_L2):
            raise ValueError("unsupported metric %r" % (metric,))
        self.d = int(d)
        self.metric_type = int(metric)
        self.is_trained = True
        self._xb = np.empty((0, self.d), dtype=np.float32)

    @property
    def ntotal(self) -> int:
        return self._xb.shape[0]

    def train(self, x: np.ndarray) -> None:  # no-op for a flat index
        _check_matrix(x, self.d)

    def add(self, x: np.ndarray) -> None:
        _check_matrix(x, self.d)
        self._xb = np.concatenate([self._xb, x.copy()], axis=0)  # add copies (faiss semantics)

    def reset(self) -> None:
        self._xb = np.empty((0, self.d), dtype=np.float32)

    def search(self, x: np.ndarray, k: int):
        _check_matrix(x, self.d)
        if k <= 0:
            raise ValueError("k must be positive")
        return knn_flat(x, self._xb, int(k), self.metric_type)


def IndexFlatIP(d: int) -> IndexFlat:
    return IndexFlat(d, METRIC_INNER_PRODUCT)


def IndexFlatL2(d: int) -> IndexFlat:
    return IndexFlat(d, METRIC_L2)


class IndexLSH:  # name must exist: seqvec_search/main.py:23 evaluates faiss.IndexLSH at import
    def __init__(self, *a, **kw):
        raise NotImplementedError("IndexLSH is outside the flat-search path")


class IndexHNSWFlat:
    def __init__(self, *a, **kw):
        raise NotImplementedError("IndexHNSWFlat is outside the flat-search path")


def scores_fp64(xq: np.ndarray, xb: np.ndarray, metric: int) -> np.ndarray:
    """Exact-arithmetic stand-in used by the parity checker to arbitrate near-ties."""
    q = xq.astype(np.float64)
    b = xb.astype(np.float64)
    ip = q @ b.T
    if metric == METRIC_INNER_PRODUCT:
        return ip
    return np.maximum((q * q).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2.0 * ip, 0.0)
