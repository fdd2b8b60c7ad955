#!/usr/bin/env python3
"""Throughput of the post-search kernels at the C4 result shape
(100k x 100) and a k = 1000 shape, against the HBM roofline, with the reference's Python loops
(oracle/postproc_oracle.py restatement) timed on a bounded sample of the same data beside them.
One JSON line per kernel.  python tools/bench_postproc.py [--nq 100000 --k 100]"""
import argparse
import json
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "knn-for-homology_b200"))
import numpy as np
import torch

import knn_b200
from knn_b200 import postproc as pp
from oracle import postproc_oracle as po

ap = argparse.ArgumentParser()
ap.add_argument("--nq", type=int, default=100_000)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--n-db", type=int, default=10_000_000)
ap.add_argument("--cpu-sample", type=int, default=2000)
args = ap.parse_args()
nq, k, n_db = args.nq, args.k, args.n_db
peaks = json.loads((REPO / "MEASURED_PEAKS.json").read_text()) if (REPO / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(5)
I = torch.randint(0, n_db, (nq, k), device=dev, generator=gen)
I[:, 0] = torch.arange(nq, device=dev)  # all-vs-all: self hit first
I[::13, 0], I[::13, 3 % k] = I[::13, 3 % k].clone(), I[::13, 0].clone()
D, _ = torch.sort(torch.rand((nq, k), device=dev, generator=gen), dim=1, descending=True)
fam_db = torch.randint(0, 5000, (n_db,), device=dev, generator=gen, dtype=torch.int32)
fam_q = torch.randint(0, 5000, (nq,), device=dev, generator=gen, dtype=torch.int32)
ident = torch.arange(n_db, device=dev)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # inputs (>= 120 MB per pass at the default shape, rotating buffers are not needed: I + D + outputs exceed nothing
    # cached between reps only when larger than L2; the k = 100 shape is 120 MB ~ L2, so flush explicitly)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def cpu(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


def line(name, ms, bytes_, cpu_s, cpu_hits, note):
    hits = nq * k
    print(json.dumps({
        "kernel": name, "shape": [nq, k], "ms": round(ms, 4), "hits_per_s": round(hits / ms * 1e3, 1),
        "roofline": {"bound": "hbm", "achieved": round(bytes_ / ms / 1e6, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": round(bytes_ / ms / 1e6 / peaks["hbm_gbs"], 4), "algorithmic_bytes": int(bytes_)},
        "cpu_baseline": {"value": round(cpu_hits / cpu_s, 1), "unit": "hits/s", "cores": 1, "kind": "port",
                         "sample": f"{cpu_hits // k} queries of the same data, reference's Python loop"},
        "note": note}), flush=True)


S = min(args.cpu_sample, nq)
I_h, D_h = I[:S].cpu().numpy(), D[:S].cpu().numpy()
fam_db_h, fam_q_h = fam_db.cpu().numpy(), fam_q.cpu().numpy()
lib = knn_b200._lib.load()
stream = torch.cuda.current_stream().cuda_stream

# ---- evaluate (main.py:53-82) -------------------------------------------------------------------
lead = torch.empty(nq, dtype=torch.int32, device=dev)
tp = torch.empty(nq, dtype=torch.int32, device=dev)
err = torch.zeros(1, dtype=torch.int32, device=dev)
ms = timed(lambda: lib.knn_eval_family_dev(nq, k, I.data_ptr(), fam_q.data_ptr(), fam_db.data_ptr(), n_db, lead.data_ptr(),
                                           tp.data_ptr(), err.data_ptr(), stream))
c = cpu(lambda: po.evaluate_counts(I_h, fam_q_h[:S], fam_db_h))
line("eval_family_kernel", ms, nq * k * (8 + 4) + nq * 12, c, S * k, "I read once + one 4-byte label gather per hit (32-byte sectors: 8x amplification on a 40 MB table)")

# ---- remove_self_hit (proteins.py:85-122) -------------------------------------------------------
I2, D2 = I.clone(), D.clone()
missing = torch.zeros(1, dtype=torch.int64, device=dev)


def rsh():
    I2.copy_(I)
    D2.copy_(D)
    lib.knn_remove_self_hit_dev(nq, k, I2.data_ptr(), D2.data_ptr(), None, missing.data_ptr(), stream)


def rsh_copy_only():
    I2.copy_(I)
    D2.copy_(D)


ms = timed(rsh) - timed(rsh_copy_only)
Ih2, Dh2 = I_h.copy(), D_h.copy()
c = cpu(lambda: po.remove_self_hit(Ih2, Dh2, np.arange(S)))
line("remove_self_hit_kernel", ms, nq * 8 + (nq // 13) * 4 * 12, c, S * k, "reads column 0 of every row; rotates the 1/13 of rows whose self hit is not first")

# ---- prefilter writer (_write_prefilter_db.py:52-97) --------------------------------------------
sec = torch.empty(nq + 1, dtype=torch.int64, device=dev)
idx = torch.empty(nq + 1, dtype=torch.int64, device=dev)
a = (nq, k, I.data_ptr(), D.data_ptr(), None, ident.data_ptr(), n_db, ident.data_ptr(), n_db, 1)
lib.knn_prefilter_measure_dev(*a, sec.data_ptr(), idx.data_ptr(), err.data_ptr(), stream)
nbytes, ibytes = int(sec[-1]), int(idx[-1])
data = torch.empty(nbytes, dtype=torch.uint8, device=dev)
index = torch.empty(ibytes, dtype=torch.uint8, device=dev)
ms_m = timed(lambda: lib.knn_prefilter_measure_dev(*a, sec.data_ptr(), idx.data_ptr(), err.data_ptr(), stream))
ms_e = timed(lambda: lib.knn_prefilter_emit_dev(*a, sec.data_ptr(), idx.data_ptr(), data.data_ptr(), index.data_ptr(), err.data_ptr(), stream))
ident_h = np.arange(n_db)
c = cpu(lambda: po.write_prefilter_db(I_h, np.arange(S), D_h, ident_h, ident_h))
line("prefilter_measure+emit", ms_m + ms_e, 2 * nq * k * (12 + 8) + nbytes + ibytes, c, S * k,
     f"measure {ms_m:.3f} ms + emit {ms_e:.3f} ms; output {nbytes} + {ibytes} bytes of text; (D, I) and the id map are read by both passes")
assert int(err.item()) == 0
