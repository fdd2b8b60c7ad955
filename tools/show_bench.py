#!/usr/bin/env python3
"""Prints the interesting keys of bench.py JSON lines (files may hold other stdout lines, e.g. NCCL's version banner)."""
import json
import sys

for f in sys.argv[1:]:
    line = [l for l in open(f) if l.startswith("{")]
    if not line:
        print(f, "no JSON line")
        continue
    d = json.loads(line[-1])
    print(f.split("/")[-1][-70:], "| overlap", d.get("overlap_finish"), "| groups", d.get("query_groups"))
    for k in ["value", "ms_per_step", "e2e", "e2e_pageable", "parity_vs_fp64", "phases_ms_rank0", "ms_per_step_by_rank",
              "shard_rows_by_rank", "clocks", "gpu_launches", "parity_spot_check"]:
        v = d.get(k)
        if isinstance(v, dict):
            v = {a: b for a, b in v.items() if a not in ("reference", "unit", "host_memory", "h2d_bytes_per_step", "d2h_bytes_per_step", "tau", "rtol_D")}
        if v is not None:
            print("  ", k, json.dumps(v)[:420])
    r = d.get("roofline") or {}
    print("  ", "roofline", {k: round(r[k], 3) for k in ["achieved", "frac", "frac_of_burst", "kernel_ms_per_step", "whole_step_frac_of_burst"] if r.get(k) is not None})
    if d.get("small_batch"):
        print("  ", "small", [(x["nq"], round(x["ms_per_call"], 3), round(x["frac_of_hbm"], 3), x["identical_to_large_batch_result"]) for x in d["small_batch"]])
