#!/usr/bin/env python3
"""Small searches touching every kernel family (for compute-sanitizer): exact scan, tcgen05 filter with bf16 and fp16
operands, both cta_group variants, L2, k = 1000, bf16 storage, ragged sizes, post-processing, merge."""
import sys

sys.path.insert(0, "knn-for-homology_b200")
import numpy as np
import torch

import knn_b200

rng = np.random.default_rng(0)


def data(n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    knn_b200.normalize_L2(x)
    return x


for (nb, nq, d, k, metric, params, kw) in [
    (5000, 37, 1024, 10, 0, dict(path=1), {}),
    (9000, 300, 1024, 100, 0, dict(path=2, shadow_fmt=1), {}),
    (9000, 300, 1024, 100, 1, dict(path=2, shadow_fmt=2), {}),
    (12345, 77, 96, 50, 0, dict(path=2, cta_group=1), {}),
    (30000, 150, 256, 1000, 0, dict(path=2), {}),
    (20000, 64, 128, 20, 0, dict(path=2), dict(bf16_storage=True)),
    (8192, 1, 1024, 5, 0, dict(path=2, query_batch=128), {}),
]:
    xb, xq = data(nb, d), data(nq, d)
    idx = knn_b200.IndexFlat(d, metric, **kw)
    for name, v in params.items():
        idx.set_param(name, v)
    idx.add(xb)
    D, I = idx.search(xq, k)
    ref = knn_b200.IndexFlat(d, metric, **kw)
    ref.set_param("path", 1)
    ref.add(xb)
    D1, I1 = ref.search(xq, k)
    assert np.array_equal(I, I1) and np.array_equal(D, D1), (nb, nq, d, k)
    print("ok", nb, nq, d, k, metric, params, kw, flush=True)

I_dev = torch.from_numpy(I).cuda()
fam_db = rng.integers(0, 50, xb.shape[0]).astype(np.int32)
fam_q = rng.integers(0, 50, xq.shape[0]).astype(np.int32)
knn_b200.evaluate_ids(I_dev, fam_q, fam_db)
ident = np.arange(xb.shape[0])
knn_b200.format_prefilter_db(I_dev, ident[:len(I)], torch.from_numpy(D).cuda(), ident, ident)
Dl = torch.sort(torch.rand(3, 40, 16, device="cuda"), dim=2, descending=True)[0]
Il = torch.stack([torch.stack([torch.randperm(1000, device="cuda")[:16] for _ in range(40)]) for _ in range(3)])
knn_b200.merge_topk(Dl, Il, 0)
torch.cuda.synchronize()
print("all ok")
