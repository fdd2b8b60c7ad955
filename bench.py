#!/usr/bin/env python3
"""Benchmark of the flat-search hot path (BASELINE.json: queries/s at 1024-d, k=100).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config C4 of BASELINE.json / SURVEY.md section 8d): synthetic 10M x 1024 fp32
database of L2-normalised N(0,1) rows (seeded per 65536-row block, so any GPU count builds
the same database), 100k normalised queries, k = 100, inner product.  One "step" = one full
pass of index.search over all queries.  With N GPUs the database is row-sharded (strong
scaling: total work fixed; shard sizes proportional to each GPU's measured speed unless
--no-balance), every rank scores all queries against its shard, the shards exchange a per-query
bound on the k-th best score (one NCCL all-reduce) before the exact rescoring, and the per-shard
results are exchanged and merged in one kernel over NVLink peer memory.

Legs, all in one JSON line printed by rank 0:
  value     queries/s with the queries already resident in HBM (device-pointer C-ABI call)
  e2e       same from pinned host queries to (D, I) in pinned host memory: N = 1 the host-pointer C-ABI call
            (H2D -> search -> D2H inside it); N > 1 every rank uploads 1/N of the queries, all-gather, search,
            rank 0 reads the result back
  roofline  the tcgen05 GEMM kernel: 2*nq*N*d flop / CUDA-event time of its launches
  cpu_baseline  blocked sgemm + top-k on the host cores (bounded sample, scaled by N)
--impl reference times the CPU implementation only (see oracle/cpu_baseline.py).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
for p in (str(REPO), str(REPO / "knn-for-homology_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

D_DIM = 1024
BLOCK_ROWS = 65536


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nb", type=int, default=10_000_000, help="database rows (C4: 10M)")
    ap.add_argument("--nq", type=int, default=100_000, help="queries per step (C4: 100k)")
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--cta-group", type=int, default=2, choices=[1, 2], help="tcgen05 cta_group of the GEMM kernel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-phases", action="store_true", help="skip the extra untimed pass that times the phases of the sharded search")
    ap.add_argument("--bf16-storage", action="store_true", help="the bf16 values ARE the database (config C5): no fp32 master rows")
    ap.add_argument("--shadow-fmt", type=int, default=0, choices=[0, 1, 2], help="16-bit format of the tensor-core copy of the database: 0 automatic, 1 bf16, 2 fp16")
    ap.add_argument("--query-groups", type=int, default=1, help="N > 1: rows x query-groups grid of ranks (GridIndexFlat; not the default: "
                    "measured on 2 GPUs only, the 8-GPU shape 2 x 4 is unmeasured)")
    ap.add_argument("--no-balance", action="store_true", help="N > 1: equal row shards instead of shards proportional to each GPU's measured speed")
    ap.add_argument("--mantissa-bits", type=int, default=0, help="mantissa bits kept in bf16 tensor-core operands: 0 automatic, 2..7")
    return ap.parse_args()


def measured_peaks():
    f = REPO / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return {"tflops_burst": d.get("bf16_tflops"), "tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic():
    """DRAM bytes per launch of the GEMM kernel from the committed `ncu --set full` capture (profiles/)."""
    f = REPO / "profiles" / "gemm_traffic.json"
    if not f.exists():
        return None, "no ncu capture committed"
    d = json.loads(f.read_text())
    return d["dram_bytes_per_launch"], d["note"]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def reference_arm(args, rank, world):
    """CPU implementation of the path on the host cores, bounded sample per step."""
    if rank != 0:
        return
    from oracle import cpu_baseline

    n_sample, nq_sample = min(args.nb, 400_000), min(args.nq, 2048)
    res = None
    times = []
    for i in range(args.warmup + args.steps):
        res = cpu_baseline.time_sample(args.nb, D_DIM, args.k, n_sample=n_sample, nq_sample=nq_sample, seed=99 + i)
        if i >= args.warmup:
            times.append(res)
    qps = statistics.mean(r["value"] for r in times)
    line = {
        "impl": "reference", "metric": "queries/s at 1024-d, k=%d, exact flat inner-product search" % args.k,
        "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * statistics.mean(r["seconds"] for r in times), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": res["cores"], "kind": res["kind"],
                         "sample": res["sample"]},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    name = "C4" if (args.nb, args.nq, args.k) == (10_000_000, 100_000, 100) else "custom"
    if (args.nb, args.nq, args.k, args.bf16_storage) == (100_000_000, 1_000_000, 1000, True):
        name = "C5"
    store = "bf16" if args.bf16_storage else "fp32"
    return {"workload": f"{name}: synthetic normalised {args.nb}x{D_DIM} {store} database, {args.nq} queries, k={args.k}, "
                        f"inner product, exact (ids = fp32 IndexFlatIP on the stored values)",
            "database_rows": args.nb, "queries_per_step": args.nq, "k": args.k, "d": D_DIM,
            "sharding": (f"rows over {world} GPU(s)" if getattr(args, "query_groups", 1) <= 1 or world == 1 else
                         f"{args.query_groups} query groups x rows over {world // args.query_groups} GPU(s)"), "l2": "inputs larger than L2 (bf16 database shard >> 126 MB)"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import knn_b200
    from knn_b200.distributed import GridIndexFlat, ShardedIndexFlat, measured_rank_speeds, shard_bounds

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- build the database shard on the device (index build is reported, not timed as search) ----
    # A step ends when the slowest shard is done, and under the board power cap the GPUs of one box do not clock
    # alike: shards are sized in proportion to each GPU's measured speed on the engine's own GEMM (2 s, untimed, part
    # of the index build like any placement decision).
    weights = None
    if world > 1 and not args.no_balance:
        weights = measured_rank_speeds(D_DIM, local_rank, seconds=2.0)
    t_build = time.perf_counter()
    Q = args.query_groups if world > 1 else 1
    if Q > 1:  # R = world / Q row shards per query group; every group holds the whole database
        index = GridIndexFlat(D_DIM, knn_b200.METRIC_INNER_PRODUCT, query_groups=Q, device=local_rank,
                              bf16_storage=args.bf16_storage, shard_weights=weights)
        R = world // Q
        b = shard_bounds(args.nb, R, weights[(rank // R) * R:(rank // R + 1) * R] if weights else None)
        lo, hi = b[rank % R], b[rank % R + 1]
    else:
        index = ShardedIndexFlat(D_DIM, knn_b200.METRIC_INNER_PRODUCT, device=local_rank, bf16_storage=args.bf16_storage,
                                 shard_weights=weights)
        b = shard_bounds(args.nb, world, weights)
        lo, hi = b[rank], b[rank + 1]
    index.local.set_param("cta_group", args.cta_group)
    if args.shadow_fmt and not args.bf16_storage:
        index.local.set_param("shadow_fmt", args.shadow_fmt)
    index.local.set_param("mantissa_bits", args.mantissa_bits)
    index.local.reserve(hi - lo)
    for blk in range(lo // BLOCK_ROWS, (hi + BLOCK_ROWS - 1) // BLOCK_ROWS):
        r0, r1 = blk * BLOCK_ROWS, min((blk + 1) * BLOCK_ROWS, args.nb)
        g = torch.Generator(device=dev).manual_seed(1234 + blk)
        rows = torch.randn(r1 - r0, D_DIM, device=dev, generator=g)
        knn_b200.normalize_L2(rows)
        s0, s1 = max(lo, r0), min(hi, r1)
        index.local.add(rows[s0 - r0:s1 - r0])
    index.adopt_local(global_start=lo, n_global=args.nb)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t_build

    g = torch.Generator(device=dev).manual_seed(4321)
    xq_dev = torch.randn(args.nq, D_DIM, device=dev, generator=g)
    knn_b200.normalize_L2(xq_dev)
    if not args.no_e2e:
        xq_host = torch.empty((args.nq, D_DIM), dtype=torch.float32, pin_memory=True)
        xq_host.copy_(xq_dev)
        D_host = torch.empty((args.nq, args.k), dtype=torch.float32, pin_memory=True)
        I_host = torch.empty((args.nq, args.k), dtype=torch.int64, pin_memory=True)
    torch.cuda.synchronize()

    lib = knn_b200._lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    last = {}

    def step_device():
        last["DI"] = index.search(xq_dev, args.k)
        return last["DI"]

    def step_e2e():
        if world == 1:
            # the raw host-pointer C-ABI call: H2D of the queries, search, D2H of (D, I)
            index.local.search_into(xq_host.data_ptr(), args.nq, args.k, D_host.data_ptr(), I_host.data_ptr())
        else:
            # every rank holds the host queries: each uploads 1/N of them, one all-gather over NVLink does the rest
            xq = index.upload_queries(xq_host) if Q == 1 else xq_host.to(dev, non_blocking=True)
            D, I = index.search(xq, args.k)
            if rank == 0:
                D_host.copy_(D, non_blocking=True)
                I_host.copy_(I, non_blocking=True)
            torch.cuda.synchronize()

    by_rank = {}

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        index.local.set_param("profile", 1 if profile else 0)
        barrier()
        launches0 = lib.knn_kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gemm_ms, gemm_launches = 0.0, 0
        e0.record()
        for _ in range(steps):
            fn()
            if profile:
                gemm_ms += index.last_stats["gemm_ms"]
                gemm_launches += int(index.last_stats["gemm_launches"])
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.knn_kernel_launches() - launches0
        index.local.set_param("profile", 0)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            per_rank = torch.zeros((world, 2), device=dev, dtype=torch.float64)
            dist.all_gather_into_tensor(per_rank, torch.tensor([ms, gemm_ms], device=dev, dtype=torch.float64))
            by_rank["step_ms"] = [round(v / steps, 3) for v in per_rank[:, 0].tolist()]
            by_rank["gemm_ms"] = [round(v / steps, 3) for v in per_rank[:, 1].tolist()]
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches, gemm_ms, gemm_launches

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, launches, gemm_ms, gemm_launches = timed(step_device, args.steps, args.warmup, profile=True)
    clocks = sampler.stop() if rank == 0 else None
    value = args.nq * args.steps / (ms / 1e3)
    by_rank_value = dict(by_rank)  # per-rank step / GEMM time of the `value` leg (chip-to-chip spread under the power cap)

    e2e = None
    if not args.no_e2e:
        ms_e, _, _, _ = timed(step_e2e, args.steps, 1)
        h2d = args.nq * D_DIM * 4
        d2h = args.nq * args.k * 12
        e2e = {"value": args.nq * args.steps / (ms_e / 1e3), "unit": "queries/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": ms_e / args.steps}

    # parity spot check inside the bench: a few queries rescored exhaustively by the exact fp32 path
    D, I = last["DI"]  # result of the last timed step
    search_path = int(index.local.stat("path"))
    fmt_names = {1: "bf16", 2: "fp16"}
    fmt = fmt_names[int(index.local.stat("shadow_fmt"))]
    mbits = int(index.local.stat("mantissa_bits"))
    chk = torch.arange(0, args.nq, max(1, args.nq // 64), device=dev)[:64]
    index.local.set_param("path", 1)
    D1, I1 = index.search(xq_dev[chk].contiguous(), args.k)
    index.local.set_param("path", 0)
    parity_ok = bool(torch.equal(I[chk], I1) and torch.equal(D[chk], D1))

    phases = None
    if world > 1 and not args.no_phases:  # one extra, untimed step with CUDA events around the phases of the sharded search
        index.profile_phases = True
        step_device()
        index.profile_phases = False
        phases = index.last_phases_ms
    if rank == 0:
        peaks = measured_peaks()
        n_shard = hi - lo
        nq_rank = args.nq
        if Q > 1:
            q_lo, q_hi = index.query_slice(args.nq)
            nq_rank = q_hi - q_lo
        flops_per_step = 2.0 * nq_rank * n_shard * D_DIM  # this rank's GEMM work: its rows x its query group's queries
        achieved = flops_per_step * args.steps / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
        peak = peaks["tflops_sustained"]
        traffic, traffic_note = ncu_traffic()
        roofline = {"bound": "tensor", "kernel": "gemm_filter_kernel (tcgen05 kind::f16, fused threshold filter)",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": (achieved / peak) if achieved else None, "peak_kind": "sustained, " + peaks["source"],
                    "frac_of_burst": (achieved / peaks["tflops_burst"]) if achieved else None,
                    "traffic": traffic, "traffic_note": traffic_note, "launches": gemm_launches, "kernel_ms_per_step": gemm_ms / args.steps,
                    "algorithmic_flop_per_launch_avg": flops_per_step * args.steps / max(1, gemm_launches)}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import cpu_baseline

            cpu = cpu_baseline.time_sample(args.nb, D_DIM, args.k, n_sample=min(args.nb, 400_000),
                                           nq_sample=min(args.nq, 4096))
            cpu = {k: cpu[k] for k in ["value", "unit", "cores", "kind", "sample"]}
        line = {
            "metric": "queries/s at 1024-d, k=%d, exact flat inner-product search" % args.k,
            "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": fmt, "dtype_note": f"{fmt} x {fmt} ({mbits} mantissa bits) -> f32 tensor-core filter (tcgen05 kind::f16), then exact f32 "
                                        "rescoring: results equal f32 IndexFlatIP",
            "data": "synthetic",
            "config": workload_config(args, world), "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": cpu, "clocks": clocks, "index_build_s": build_s, "parity_spot_check": parity_ok,
            "search_path": search_path, "cta_group": args.cta_group, "phases_ms_rank0": phases, "ms_per_step_by_rank": by_rank_value or None,
            "shard_rows_by_rank": [b[r + 1] - b[r] for r in range(len(b) - 1)], "query_groups": Q,
            "rank_speed_weights": [round(w, 4) for w in weights] if weights else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
