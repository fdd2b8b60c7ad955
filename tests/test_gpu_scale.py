"""Exactness at scale without a CPU oracle (the CPU port needs minutes at these sizes): an fp64 matmul on the GPU
(torch, checker only) gives the true scores of a query sample against EVERY database row; the engine's (D, I)
must be a valid top-k of them up to fp32 accumulation noise (tau of oracle/parity.py).  Plus size-independent
properties: sortedness, uniqueness, batch independence, power-of-two scaling, idempotence."""
import numpy as np
import pytest

from oracle.parity import tie_tolerance

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def knn():
    import knn_b200

    assert knn_b200._lib.load().knn_device_count() >= 1
    return knn_b200


def _check_against_fp64(xq_s, xb, D_s, I_s, metric, k):
    """xq_s: (s, d) sample queries, D_s/I_s the engine's rows for them."""
    import torch

    d = xb.shape[1]
    q64 = xq_s.double()
    S = torch.empty((xq_s.shape[0], xb.shape[0]), dtype=torch.float64, device=xb.device)
    step = 1 << 18
    for j in range(0, xb.shape[0], step):
        y = xb[j:j + step].double()
        ip = q64 @ y.T
        S[:, j:j + step] = ip if metric == 0 else ((q64 * q64).sum(1, keepdim=True) + (y * y).sum(1)[None, :] - 2 * ip).clamp_(min=0)
    qn = q64.norm(dim=1)
    bn = torch.stack([xb[j:j + step].double().norm(dim=1).max() for j in range(0, xb.shape[0], step)]).max()
    largest = metric == 0
    tau = tie_tolerance(d) * (qn * bn if largest else qn * qn + bn * bn)
    top = torch.topk(S, k, dim=1, largest=largest)
    kth, exact = top.values[:, -1], top.indices
    ours = torch.gather(S, 1, I_s)
    if largest:
        assert (ours >= (kth - tau)[:, None]).all(), "an id outside the true top-k (beyond fp32 noise) was returned"
    else:
        assert (ours <= (kth + tau)[:, None]).all(), "an id outside the true top-k (beyond fp32 noise) was returned"
    # distances: 1e-5 relative (+ tau near zero)
    assert ((D_s.double() - ours).abs() <= 1e-5 * ours.abs() + tau[:, None]).all()
    # near-tie swaps are rare: the id SETS agree except for a handful of boundary exchanges
    same = sum(len(set(a) & set(b)) for a, b in zip(I_s.tolist(), exact.tolist()))
    assert same >= 0.999 * I_s.numel(), (same, I_s.numel())


def _basic_properties(D, I, n, largest):
    import torch

    diff = D[:, 1:] - D[:, :-1]
    assert (diff <= 0).all() if largest else (diff >= 0).all()
    assert int(I.min()) >= 0 and int(I.max()) < n
    srt, _ = torch.sort(I, dim=1)
    assert (srt[:, 1:] != srt[:, :-1]).all(), "duplicate id in a row"


def test_c4_shape_slice_ip_exact_against_fp64(knn):
    """C4 shape at 1/5 scale: 2M x 1024 normalised rows, 20,000 queries (two query batches, the second ragged), k=100."""
    import torch

    g = torch.Generator(device="cuda").manual_seed(99)
    n, nq, k = 2_000_000, 20_000, 100
    idx = knn.IndexFlat(1024, 0)
    idx.reserve(n)
    blocks = []
    for i in range(0, n, 1 << 19):
        x = torch.randn(min(1 << 19, n - i), 1024, device="cuda", generator=g)
        knn.normalize_L2(x)
        idx.add(x)
        blocks.append(x)
    xb = torch.cat(blocks)
    del blocks
    xq = torch.randn(nq, 1024, device="cuda", generator=g)
    knn.normalize_L2(xq)
    D, I = idx.search(xq, k)
    assert idx.stat("path") == 2 and idx.stat("overflow_batches") == 0
    _basic_properties(D, I, n, True)
    sample = torch.arange(0, nq, 79, device="cuda")[:254]
    _check_against_fp64(xq[sample], xb, D[sample], I[sample], 0, k)
    # idempotent, and independent of what else is in the batch (bit-identical)
    D2, I2 = idx.search(xq, k)
    assert torch.equal(D, D2) and torch.equal(I, I2)
    rows = torch.tensor([0, 1, 4097, 16383, 16384, 19999], device="cuda")
    D3, I3 = idx.search(xq[rows].contiguous(), k)
    assert torch.equal(D3, D[rows]) and torch.equal(I3, I[rows])
    # scaling the queries by a power of two scales every score exactly and keeps the ranking
    D4, I4 = idx.search((xq[:4096] * 4.0).contiguous(), k)
    assert torch.equal(I4, I[:4096]) and torch.equal(D4, D[:4096] * 4.0)
    # the exact fp32 scan agrees bit for bit on a sample
    idx.set_param("path", 1)
    D5, I5 = idx.search(xq[sample[:32]].contiguous(), k)
    assert torch.equal(D5, D[sample[:32]]) and torch.equal(I5, I[sample[:32]])


def test_l2_unnormalised_odd_width_exact_against_fp64(knn):
    """Euclidean pass of cath/search.py:30-32 at scale: un-normalised rows of width 1000 (padded to 1024), k=11."""
    import torch

    g = torch.Generator(device="cuda").manual_seed(7)
    n, nq, k = 600_000, 5000, 11
    xb = torch.randn(n, 1000, device="cuda", generator=g) * (1 + torch.rand(n, 1, device="cuda", generator=g))
    xq = xb[torch.randint(0, n, (nq,), device="cuda", generator=g)] + 0.05 * torch.randn(nq, 1000, device="cuda", generator=g)
    idx = knn.IndexFlat(1000, 1)
    idx.add(xb)
    D, I = idx.search(xq, k)
    assert idx.stat("path") == 2 and idx.stat("overflow_batches") == 0
    _basic_properties(D, I, n, False)
    assert (D >= 0).all()
    sample = torch.arange(0, nq, 20, device="cuda")[:250]
    _check_against_fp64(xq[sample], xb, D[sample], I[sample], 1, k)


def test_k1000_bf16_storage_slice(knn):
    """C5 shape slice: bf16-only storage, k = 1000: the stored (bf16) values are the database, results must be the exact
    top-k of THOSE values."""
    import torch

    g = torch.Generator(device="cuda").manual_seed(11)
    n, nq, k = 500_000, 3000, 1000
    x = torch.randn(n, 1024, device="cuda", generator=g)
    knn.normalize_L2(x)
    idx = knn.IndexFlat(1024, 0, bf16_storage=True)
    idx.add(x)
    xb = x.to(torch.bfloat16).float()  # what the index holds (round to nearest even)
    xq = torch.randn(nq, 1024, device="cuda", generator=g)
    knn.normalize_L2(xq)
    D, I = idx.search(xq, k)
    assert idx.stat("path") == 2 and idx.stat("overflow_batches") == 0
    _basic_properties(D, I, n, True)
    sample = torch.arange(0, nq, 24, device="cuda")[:125]
    _check_against_fp64(xq[sample], xb, D[sample], I[sample], 0, k)
