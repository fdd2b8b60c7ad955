"""GPU tests of the exchange-fused merge over peer memory (knn_merge_topk_peer_dev, PeerExchange).

1. the kernel alone, G virtual ranks on one device, against the all-gather merge kernel (bit-identical);
2. the whole IPC path: two processes sharing cuda:0 (gloo for the host-side handshakes, CUDA IPC for the
   buffers) run ShardedIndexFlat with peer_merge and must return what one unsharded index returns."""
import ctypes
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("G,nq,k,metric", [(8, 1000, 100, 0), (2, 333, 1000, 0), (3, 50, 7, 1), (16, 64, 100, 1), (5, 17, 1, 0),
                                           (8, 40, 2048, 0)])
def test_peer_merge_kernel_equals_allgather_merge(G, nq, k, metric):
    import torch

    import knn_b200
    from knn_b200 import _lib

    lib = _lib.load()
    gen = torch.Generator(device="cuda").manual_seed(G * nq + k)
    # sorted per-shard lists with global ids, some shards short (padding -1) and many equal scores across shards
    D = torch.randint(0, 50, (G, nq, k), device="cuda", generator=gen).float() / 7
    D, _ = torch.sort(D, dim=2, descending=(metric == 0))
    ids = torch.stack([torch.randperm(G * k * 4, device="cuda", generator=gen)[:G * k] for _ in range(nq)])  # distinct per query
    I = ids.view(nq, G, k).permute(1, 0, 2).contiguous()
    # equal scores inside a list must be ordered by id for the list to be a valid search result
    key = D.double() * (1 if metric else -1) * 1e6 + I.double() / (G * k * 8)
    order = torch.argsort(key, dim=2)
    D, I = torch.gather(D, 2, order), torch.gather(I, 2, order)
    short = k // 3
    if short:
        I[0, :, k - short:] = -1
        D[0, :, k - short:] = -3.4028234663852886e38 if metric == 0 else 3.4028234663852886e38
    D_ref, I_ref = knn_b200.merge_topk(D, I, metric)
    outs_D = [torch.full((nq, k), 7.0, device="cuda") for _ in range(G)]
    outs_I = [torch.full((nq, k), 7, dtype=torch.int64, device="cuda") for _ in range(G)]
    arr = lambda ts: (ctypes.c_void_p * G)(*[t.data_ptr() for t in ts])  # noqa: E731
    b = [nq * r // G for r in range(G + 1)]
    for r in range(G):  # every virtual rank merges its slice into everybody's output
        _lib.check(lib.knn_merge_topk_peer_dev(metric, nq, k, G, b[r], b[r + 1], arr(list(D)), arr(list(I)), arr(outs_D), arr(outs_I),
                                               torch.cuda.current_stream().cuda_stream))
    for r in range(G):
        assert torch.equal(outs_I[r], I_ref) and torch.equal(outs_D[r], D_ref), r


def test_peer_merge_limits():
    from knn_b200 import _lib

    lib = _lib.load()
    p = (ctypes.c_void_p * 17)(*[1] * 17)
    assert lib.knn_merge_topk_peer_dev(0, 10, 5, 17, 0, 10, p, p, p, p, None) == -4
    assert lib.knn_merge_topk_peer_dev(0, 10, 5, 2, 0, 11, p, p, p, p, None) == -1


def _worker(rank, world, port, out):
    sys.path.insert(0, str(REPO))
    sys.path.insert(0, str(REPO / "knn-for-homology_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    import knn_b200
    from knn_b200.distributed import ShardedIndexFlat

    rng = np.random.default_rng(0)
    xb = rng.standard_normal((20011, 64)).astype(np.float32)
    xq = rng.standard_normal((301, 64)).astype(np.float32)
    for metric in (0, 1):
        # gloo moves no CUDA tensors here: the bound exchange (an all-reduce of device tensors) stays off
        # metric 1 also runs with speed-proportional (unequal) shards: placement must not change the result
        index = ShardedIndexFlat(64, metric, device=0, exchange_bounds=False, peer_merge=True,
                                 shard_weights=[0.7, 1.3] if metric == 1 else None)
        index.add(xb[:7000])
        index.add(xb[7000:])
        if metric == 1:
            assert abs(index.local.ntotal - 20011 * (0.35 if rank == 0 else 0.65)) <= 2
        single = knn_b200.IndexFlat(64, metric, device=0)
        single.add(xb)
        xq_d = torch.from_numpy(xq).cuda()
        for k in (10, 100, 1000):  # the second and third searches regrow / reuse the mapped buffers
            D, I = index.search(xq_d, k)
            assert index._exchange is not None and index._exchange.capacity >= 24 * 301 * k
            D1, I1 = single.search(xq_d, k)
            assert torch.equal(I, I1) and torch.equal(D, D1), (rank, metric, k)
        index._exchange.close()
    if rank == 0:
        Path(out).write_text("ok")
    dist.destroy_process_group()


def test_sharded_search_over_ipc_peer_memory_two_processes_one_gpu(tmp_path):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = tmp_path / "ok"
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"
