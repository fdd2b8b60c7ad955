#!/usr/bin/env python3
"""One profiled small-batch search (for `ncu --profile-from-start off`): N x 1024 rows, NQ queries, k = 100.
  python tools/prof_small.py N NQ [main]     ("main": force the main kernel instead of the few-queries variant)"""
import sys

sys.path.insert(0, "knn-for-homology_b200")
import torch

import knn_b200

N, NQ = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(3)
idx = knn_b200.IndexFlat(1024, 0)
idx.reserve(N)
for i in range(0, N, 1 << 20):
    x = torch.randn(min(1 << 20, N - i), 1024, device=dev, generator=g)
    knn_b200.normalize_L2(x)
    idx.add(x)
if "main" in sys.argv[3:]:
    idx.set_param("stream_kernel", 0)
xq = torch.randn(NQ, 1024, device=dev, generator=g)
knn_b200.normalize_L2(xq)
for _ in range(3):
    idx.search(xq, 100)
torch.cuda.synchronize()
torch.cuda.profiler.start()
idx.search(xq, 100)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", flush=True)
