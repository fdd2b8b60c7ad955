"""Seeded randomised GPU tests: many random shapes and parameter combinations, every one compared bit for bit
between the tensor-core path (filter + exact rescoring), the two-phase row-sharded path and the exact fp32 scan
of the same engine, and a sample of them against the CPU oracle.  The fixed-shape tests in test_gpu_parity.py
pin the known edge cases; this one looks for the ones nobody thought of (ragged tiles, panel boundaries,
k against list capacities, clustered scores, mixed norms)."""
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "knn-for-homology_b200"))

from oracle import flat_oracle as fo  # noqa: E402
from oracle.parity import check_parity  # noqa: E402

pytestmark = pytest.mark.gpu
IP, L2 = fo.METRIC_INNER_PRODUCT, fo.METRIC_L2


def _case(rng):
    d = int(rng.choice([8, 20, 64, 96, 100, 128, 256, 300, 512, 1024, 1280]))
    nb = int(rng.choice([257, 1000, 1024, 1025, 4095, 4097, 8192, 12345, 20000, 33333, 65537, 131073, 200000]))
    if nb > 65537:
        d = min(d, 256)  # keeps the host-side data generation of the big cases short
    nq = int(rng.choice([1, 2, 31, 127, 128, 129, 255, 256, 257, 400, 700]))
    k = int(rng.choice([1, 2, 5, 10, 11, 13, 64, 100, 255, 256, 257, 500, 1000, 2000]))
    k = min(k, nb)
    metric = int(rng.choice([IP, L2]))
    kind = str(rng.choice(["gauss", "normalised", "clustered", "mixed_norms", "duplicates"]))
    params = dict(cta_group=int(rng.choice([1, 2])), shadow_fmt=int(rng.choice([0, 1, 2])),
                  query_batch=int(rng.choice([128, 256, 16384])))
    return d, nb, nq, k, metric, kind, params


def _rows(rng, n, d, kind, centres=None):
    x = rng.standard_normal((n, d)).astype(np.float32)
    if kind == "clustered":
        x = centres[rng.integers(0, len(centres), n)] + 0.3 * x
    if kind == "mixed_norms":
        x *= rng.choice([0.01, 1.0, 30.0], size=(n, 1)).astype(np.float32)
    if kind in ("normalised", "clustered"):
        fo.normalize_L2(x)
    return np.ascontiguousarray(x, dtype=np.float32)


@pytest.mark.parametrize("seed", range(100))
def test_random_shapes_all_device_paths_agree(seed):
    import torch

    import knn_b200

    rng = np.random.default_rng(1000 + seed)
    d, nb, nq, k, metric, kind, params = _case(rng)
    centres = rng.standard_normal((37, d)).astype(np.float32)
    xb, xq = _rows(rng, nb, d, kind, centres), _rows(rng, nq, d, kind, centres)
    if kind == "duplicates":  # exact ties between database rows, and queries that are database rows
        xb[nb // 2:] = xb[:nb - nb // 2]
        xq[: min(nq, 50)] = xb[: min(nq, 50)]
    what = dict(seed=seed, d=d, nb=nb, nq=nq, k=k, metric=metric, kind=kind, **params)

    exact = knn_b200.IndexFlat(d, metric)
    exact.set_param("path", 1)
    exact.add(xb)
    D1, I1 = exact.search(xq, k)

    tensor = knn_b200.IndexFlat(d, metric)
    tensor.set_param("path", 2)
    for name, v in params.items():
        tensor.set_param(name, v)
    tensor.add(xb[: nb // 3])  # incremental adds, ragged
    tensor.add(xb[nb // 3:])
    D2, I2 = tensor.search(xq, k)
    assert tensor.stat("path") == 2, what
    assert np.array_equal(I1, I2), what
    assert np.array_equal(D1, D2), what

    # two-phase row-sharded search over three ragged shards (bound exchange emulated in one process)
    dev = torch.device("cuda:0")
    tq = torch.from_numpy(xq).to(dev)
    bounds = [0, max(1, nb // 5), max(2, (2 * nb) // 3), nb]
    shards = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        sh = knn_b200.IndexFlat(d, metric)
        sh.set_param("cta_group", params["cta_group"])
        if seed % 2:  # small shards on the tensor path too (automatic: the exact scan below 8192 rows)
            sh.set_param("path", 2)
        sh.add(torch.from_numpy(xb[a:b]).to(dev))
        shards.append(sh)
    j = -(-k // len(shards))
    both = [sh.search_filter(tq, k, j) for sh in shards]
    lower = torch.maximum(torch.stack([b[0] for b in both]).max(dim=0).values,
                          torch.stack([b[1] for b in both]).min(dim=0).values)
    Ds, Is = zip(*[sh.search_finish(lower, k, id_base=a) for sh, a in zip(shards, bounds[:-1])])
    Dm, Im = knn_b200.merge_topk(torch.stack(Ds), torch.stack(Is), metric)
    assert np.array_equal(Im.cpu().numpy(), I1), what
    assert np.array_equal(Dm.cpu().numpy(), D1), what

    if seed % 3 == 0 and kind in ("gauss", "normalised", "clustered"):  # and against the CPU restatement of the reference
        D_ref, I_ref = fo.knn_flat(xq, xb, k, metric)
        check_parity(D2, I2, D_ref, I_ref, xq, xb, metric, max_excused_frac=2e-2)
