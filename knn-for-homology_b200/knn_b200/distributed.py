"""Row-sharded flat index over several GPUs of one box (one process per GPU).

New relative to the reference (which is single-process): the database is
split by rows, every rank scores ALL queries against its shard with the single-GPU engine,
the per-shard (D, I) are exchanged with one NCCL all-gather over NVLink and merged per query
on the device (knn_merge_topk_dev).  Ids stay global row numbers, so (D, I) is still a
drop-in for the reference's consumers.  Tie rule is shard-count invariant: equal score ->
lower global id.

``index_factory`` / ``merge_fn`` exist so that the host logic (row split, id mapping,
collective, merge call) can be exercised under gloo on CPU in tests.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, world: int):
    """Contiguous near-equal row ranges: rank r owns [b[r], b[r+1])."""
    return [(n * r) // world for r in range(world + 1)]


class ShardedIndexFlat:
    TWO_PHASE_MAX_QUERIES = 131072  # limit of knn_index_search_filter_dev (candidate lists stay resident)

    def __init__(self, d: int, metric: int, group=None, device=None, index_factory=None, merge_fn=None,
                 exchange_bounds: bool = True, **index_kw):
        import torch.distributed as dist

        self._dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.d = int(d)
        self.metric_type = int(metric)
        self.is_trained = True
        if index_factory is None:
            from .index import IndexFlat

            index_factory = lambda: IndexFlat(d, metric, device=device, **index_kw)  # noqa: E731
        if merge_fn is None:
            from .index import merge_topk

            merge_fn = merge_topk
        self.exchange_bounds = exchange_bounds
        self.profile_phases = False
        self.last_phases_ms = None
        self.local = index_factory()
        self._merge = merge_fn
        # one (global_start, local_start, count) triple per add() call
        self._segments = []
        self._ntotal = 0
        self._nlocal = 0

    @property
    def ntotal(self) -> int:
        return self._ntotal

    def train(self, x) -> None:
        pass

    def add(self, x) -> None:
        """Every rank passes the same rows (as the reference's drivers would); rank r keeps its
        contiguous slice.  Global id of a row = its position in the concatenation of all adds."""
        n = x.shape[0]
        b = shard_bounds(n, self.world)
        lo, hi = b[self.rank], b[self.rank + 1]
        self.add_local(x[lo:hi], global_start=self._ntotal + lo, n_global=n)

    def add_local(self, x_local, global_start: int, n_global: int) -> None:
        """For data generated per shard: this rank's rows are x_local, their global ids start at
        global_start; n_global rows are being added across all ranks."""
        cnt = x_local.shape[0]
        if cnt:
            self.local.add(x_local)
        self._segments.append((int(global_start), self._nlocal, cnt))
        self._nlocal += cnt
        self._ntotal += int(n_global)

    def adopt_local(self, global_start: int, n_global: int) -> None:
        """Book-keeping for rows that were added straight into ``self.local`` (e.g. generated
        block by block on the device): they become one segment starting at global_start."""
        cnt = self.local.ntotal - self._nlocal
        self._segments.append((int(global_start), self._nlocal, cnt))
        self._nlocal += cnt
        self._ntotal += int(n_global)

    def _to_global(self, I):
        """local row numbers -> global row numbers, -1 kept."""
        import torch

        if len(self._segments) == 1:
            g0 = self._segments[0][0]
            return torch.where(I >= 0, I + g0, I)
        starts = torch.tensor([s[1] for s in self._segments], device=I.device, dtype=I.dtype)
        offs = torch.tensor([s[0] - s[1] for s in self._segments], device=I.device, dtype=I.dtype)
        seg = torch.bucketize(I.clamp(min=0), starts, right=True) - 1
        return torch.where(I >= 0, I + offs[seg], I)

    def search(self, x, k: int):
        """Returns the merged (D, I) on every rank (torch tensors on the local index's device, or
        numpy arrays when x is numpy)."""
        import torch

        as_numpy = isinstance(x, np.ndarray)
        marks = []

        def mark(name):
            if self.profile_phases:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))

        mark("start")
        two_phase = (self.world > 1 and self.exchange_bounds and hasattr(self.local, "search_filter")
                     and x.shape[0] <= self.TWO_PHASE_MAX_QUERIES)
        if two_phase:
            # filter on every shard -> all-reduce(MAX) of the per-query lower bounds of the k-th best score ->
            # each shard rescoring only what can still be in the global top-k
            xd = torch.from_numpy(x).to(torch.device("cuda", self.local.device)) if as_numpy else x
            # two bounds on the global k-th best true score: the best shard's own k-th, and - every shard holding
            # j = ceil(k/G) rows at or above its own j-th - the smallest j-th over the shards
            j = -(-k // self.world)
            lower, lower_j = self.local.search_filter(xd, k, j)
            mark("filter")
            self._dist.all_reduce(lower, op=self._dist.ReduceOp.MAX, group=self.group)
            self._dist.all_reduce(lower_j, op=self._dist.ReduceOp.MIN, group=self.group)
            lower = torch.maximum(lower, lower_j)
            mark("all_reduce_bounds")
            D, I = self.local.search_finish(lower, k)
            mark("finish")
        else:
            D, I = self.local.search(x, k)
            if as_numpy:
                D, I = torch.from_numpy(D), torch.from_numpy(I)
                backend = self._dist.get_backend(self.group) if self._dist.is_initialized() else "gloo"
                if backend == "nccl":
                    dev = torch.device("cuda", self.local.device)
                    D, I = D.to(dev), I.to(dev)
        I = self._to_global(I)
        mark("to_global")
        if self.world > 1:
            nq, kk = D.shape
            Dg = torch.empty((self.world * nq, kk), dtype=D.dtype, device=D.device)
            Ig = torch.empty((self.world * nq, kk), dtype=I.dtype, device=I.device)
            self._dist.all_gather_into_tensor(Dg, D.contiguous(), group=self.group)
            self._dist.all_gather_into_tensor(Ig, I.contiguous(), group=self.group)
            mark("all_gather")
            D, I = self._merge(Dg.view(self.world, nq, kk), Ig.view(self.world, nq, kk), self.metric_type)
            mark("merge")
        if marks:
            torch.cuda.synchronize()
            self.last_phases_ms = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(marks[:-1], marks[1:])}
        if as_numpy:
            return D.cpu().numpy(), I.cpu().numpy()
        return D, I
