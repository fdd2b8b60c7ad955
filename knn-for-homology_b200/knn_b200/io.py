"""faiss-compatible persistence of a flat index.

``faiss.write_index(index, path)`` is called on the flat index by pfam/proteins_search.py:39-40
and ``faiss.read_index`` by seqvec_search/main.py:131-132.  File layout of faiss 1.7.x
``write_index`` for IndexFlat (upstream knowledge - faiss is not installable here, so the
byte-compatibility is unverified against a real faiss):

    fourcc   4 bytes  "IxFI" (inner product) | "IxF2" (L2)
    d        int32
    ntotal   int64
    dummy    int64 = 1 << 20, twice
    trained  uint8
    metric   int32
    count    uint64 = ntotal * d, then count float32 values (row-major vectors)
"""
from __future__ import annotations

import struct

import numpy as np

from .index import METRIC_INNER_PRODUCT, METRIC_L2, IndexFlat

_HEADER = struct.Struct("<4siqqqBi")


def write_index(index: IndexFlat, path: str) -> None:
    """Any flat index works that has ``d``, ``ntotal``, ``metric_type`` and ``reconstruct_n(i0, n)`` (the CPU test of
    the byte layout passes a host-side stand-in; the product passes an IndexFlat)."""
    if not all(hasattr(index, a) for a in ("d", "ntotal", "metric_type", "reconstruct_n")):
        raise TypeError("write_index: only flat indexes are supported")
    if index.metric_type not in (METRIC_INNER_PRODUCT, METRIC_L2):
        raise TypeError("write_index: only flat indexes are supported")
    fourcc = b"IxFI" if index.metric_type == METRIC_INNER_PRODUCT else b"IxF2"
    n = index.ntotal
    with open(path, "wb") as f:
        f.write(_HEADER.pack(fourcc, index.d, n, 1 << 20, 1 << 20, 1, index.metric_type))
        f.write(struct.pack("<Q", n * index.d))
        step = max(1, (64 << 20) // (4 * index.d))
        for i0 in range(0, n, step):
            np.ascontiguousarray(index.reconstruct_n(i0, min(step, n - i0)), dtype="<f4").tofile(f)


def read_index(path: str, device: int | None = None, index_factory=None) -> IndexFlat:
    """``index_factory(d, metric)``: what to build instead of a device IndexFlat (CPU tests of the byte layout)."""
    with open(path, "rb") as f:
        head = f.read(_HEADER.size)
        if len(head) != _HEADER.size:
            raise ValueError("read_index: truncated header")
        fourcc, d, n, _d1, _d2, _trained, metric = _HEADER.unpack(head)
        if fourcc not in (b"IxFI", b"IxF2", b"IxFl"):
            raise NotImplementedError("read_index: %r is not a flat index" % fourcc)
        if metric not in (METRIC_INNER_PRODUCT, METRIC_L2):
            raise NotImplementedError("read_index: metric %d" % metric)
        (count,) = struct.unpack("<Q", f.read(8))
        if count != n * d:
            raise ValueError("read_index: vector count %d != ntotal*d %d" % (count, n * d))
        if index_factory is None:
            index = IndexFlat(d, metric, device=device)
            index.reserve(n)
        else:
            index = index_factory(d, metric)
        step = max(1, (64 << 20) // (4 * d))
        for i0 in range(0, n, step):
            m = min(step, n - i0)
            rows = np.fromfile(f, dtype=np.float32, count=m * d)
            if rows.size != m * d:
                raise ValueError("read_index: truncated vectors")
            index.add(rows.reshape(m, d))
    return index
