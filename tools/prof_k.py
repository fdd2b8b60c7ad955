#!/usr/bin/env python3
"""One profiled search of a k-heavy configuration (for `ncu --profile-from-start off` launch lists).
  python tools/prof_k.py N NQ K [bf16] [self]"""
import sys

sys.path.insert(0, "knn-for-homology_b200")
import torch

import knn_b200

N, NQ, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
bf16 = "bf16" in sys.argv[4:]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(3)
idx = knn_b200.IndexFlat(1024, 0, bf16_storage=bf16)
idx.reserve(N)
for i in range(0, N, 1 << 20):
    x = torch.randn(min(1 << 20, N - i), 1024, device=dev, generator=g)
    knn_b200.normalize_L2(x)
    idx.add(x)
if "self" in sys.argv[4:]:
    xq = torch.from_numpy(idx.reconstruct_n(0, NQ)).to(dev) if hasattr(idx, "reconstruct_n") else x[:NQ].contiguous()
else:
    xq = torch.randn(NQ, 1024, device=dev, generator=g)
    knn_b200.normalize_L2(xq)
idx.search(xq, K)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
idx.search(xq, K)
e1.record()
torch.cuda.synchronize()
print("search ms", e0.elapsed_time(e1), flush=True)
torch.cuda.profiler.start()
idx.search(xq, K)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
