#!/usr/bin/env python
"""Headline benchmark: pandrs groupby-aggregate (BASELINE.json configs[1]) on N B200s, one process per GPU.

  python bench.py --gpus 1 --steps K --warmup W              # the CUDA path (libpandrs_b200.so)
  python bench.py --impl reference --steps K --warmup W      # the reference's CPU algorithm (oracle port) on host cores

A step = one groupby(key).agg(sum, mean, min, max, count, std) over `rows` synthetic rows per GPU
(i64 key with 1000 distinct values, f64 value, 5% NULL values).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_ROW = 16.125          # 8 B key + 8 B value + 1 bit NULL bitmap, read once
NULL_PER_MILLION = 50_000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000_000, help="rows per GPU")
    ap.add_argument("--groups", type=int, default=1000)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-rows", type=int, default=4_000_000, help="rows of the bounded CPU-baseline sample")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled every 20 ms from before the warm-up; the summary uses the samples that fall inside the
    timed region (mark_begin / mark_end), or, if the region was shorter than one sample, the samples under load."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.samples, self.proc, self.index, self.t0, self.t1 = [], None, index, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()

        def parse(rows):
            sm, mx, reasons = [], 0, set()
            for _, s in rows:
                f = [x.strip() for x in s.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0])); mx = max(mx, float(f[1]))
                except ValueError:
                    continue
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons
        inside = [x for x in self.samples if self.t0 is not None and self.t1 is not None and self.t0 <= x[0] <= self.t1 + 0.03]
        sm, mx, reasons = parse(inside)
        where = "timed region"
        if not sm:
            sm, mx, reasons = parse(self.samples)
            where = "whole run (timed region shorter than one sample)"
        sm.sort()
        hi = [x for x in sm if x > 0.5 * mx] or sm
        return {"sm_mhz": hi[len(hi) // 2] if hi else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm), "window": where}


# ---------------------------------------------------------------- the reference arm / CPU baseline
def cpu_reference_step(orc, n, groups, threads, seed=42):
    """One pass of the reference algorithm (string keys, HashMap<Vec<String>, Vec<usize>>, per-group per-aggregate
    gathers; par_aggregate's thread pool over groups) on n synthetic rows.  Returns seconds."""
    k = orc.synth_keys(n, seed=seed, card=groups)
    v = orc.synth_vals(n, seed=seed)
    vn = orc.synth_nulls(n, seed=seed, per_million=NULL_PER_MILLION)
    ops = [orc.SUM, orc.MEAN, orc.MIN, orc.MAX, orc.COUNT, orc.STD]
    t0 = time.perf_counter()
    r = orc.groupby([orc.Col(orc.I64, k)], [orc.Col(orc.F64, v, vn)], [(0, op) for op in ops], mode=orc.MODE_PAR_AGGREGATE, nthreads=threads,
                    want_key_strings=False)
    dt = time.perf_counter() - t0
    assert r["n_groups"] == min(groups, n) or n < 50 * groups
    return dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as orc
    orc.build()
    threads = os.cpu_count() or 1
    n = args.cpu_rows
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_reference_step(orc, min(n, 200_000), args.groups, threads)
    ts = [cpu_reference_step(orc, n, args.groups, threads) for _ in range(args.steps)]
    dt = sum(ts) / len(ts)
    val = n / dt
    sample = f"{n} rows per step (same generator / key cardinality / 5% NULLs as the GPU arm), grouping serial like grouping.rs:62-104, aggregation over groups on {threads} threads like aggregation.rs:81"
    print(json.dumps({
        "impl": "reference", "metric": "groupby_agg_rows_per_s", "value": val, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, n),
        "cpu_baseline": {"value": val, "unit": "rows/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, rows):
    return {"workload": f"groupby-agg {rows} rows/GPU, i64 key {args.groups} distinct, sum/mean/min/max/count/std of f64, 5% nulls (BASELINE.json configs[1])",
            "rows_per_gpu": rows, "groups": args.groups, "aggs": "sum,mean,min,max,count,std", "null_fraction": 0.05,
            "cache": "inputs (16 GB per GPU) larger than L2; no flush needed"}


# ---------------------------------------------------------------- the CUDA arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch

    import pandrs_b200 as pb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream()
    ctx = pb.Context(device=local, stream=stream.cuda_stream)
    n = args.rows
    ops = [pb.SUM, pb.MEAN, pb.MIN, pb.MAX, pb.COUNT, pb.STD]
    aggs = [(0, op) for op in ops]

    keys = ctx.synth_keys(n, card=args.groups, row0=rank * n)
    vals = ctx.synth_vals(n, row0=rank * n, null_per_million=NULL_PER_MILLION)
    ctx.sync()

    if world > 1:
        from pandrs_b200.dist import DistGroupBy
        dgb = DistGroupBy(ctx, dist)

        def step():
            r = dgb.groupby_agg_lowcard([keys], [vals], aggs)
            r.close()
    else:
        def step():
            r = ctx.groupby_agg([keys], [vals], aggs)
            r.close()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        step()
    launches0 = ctx.stats()["kernel_launches"]
    barrier()
    clocks.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    e0.record(stream)
    for _ in range(args.steps):
        step()
        kernel_ms.append(ctx.stats()["main_kernel_ms"])
    e1.record(stream)
    barrier()
    clocks.mark_end()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clk = clocks.stop() if rank == 0 else None
    launches = (ctx.stats()["kernel_launches"] - launches0) // max(args.steps, 1)
    ms_per_step = ms / args.steps
    value = n * world / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (gb_tsort_kernel at 1K groups): algorithmic bytes / CUDA-event duration of that launch
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    k_ms = sum(kernel_ms) / len(kernel_ms)
    achieved = ALG_BYTES_PER_ROW * n / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": {pb.GB_TILESORT: "gb_tsort_kernel", pb.GB_SHARED: "gb_shared_kernel"}.get(ctx.stats()["groupby_algo_used"], "gb_global_kernel"),
                "kernel_ms": k_ms, "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6.65 TB/s (of fallback)",
                "alg_bytes_per_row": ALG_BYTES_PER_ROW}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            tr = json.load(open(traffic_file))
            roofline["traffic"] = tr.get(roofline["kernel"] + "_bytes_per_row", 0) * n or None
        except Exception:
            pass

    out = {"metric": "groupby_agg_rows_per_s", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args, n), "roofline": roofline, "gpu_launches": int(launches), "clocks": clk}

    # ---- end to end through the C ABI with HOST (pinned) buffers: H2D of the inputs and D2H of the result inside the timed region
    if not args.no_e2e:
        nb = (n + 7) // 8
        hk, hv, hn = ctx.host_alloc(n * 8), ctx.host_alloc(n * 8), ctx.host_alloc(nb)
        ctx.memcpy(hk, keys.ptr, n * 8, 1)
        ctx.memcpy(hv, vals.ptr, n * 8, 1)
        ctx.memcpy(hn, vals.nulls_ptr, nb, 1)
        hkeys = pb.Column(pb.I64, device_ptr=hk, length=n)
        hvals = pb.Column(pb.F64, device_ptr=hv, nulls_ptr=hn, null_len=nb, length=n)
        hkeys.mem = hvals.mem = pb.MEM_HOST
        d2h = 0

        def e2e_step():
            nonlocal d2h
            r = ctx.groupby_agg([hkeys], [hvals], aggs)
            k, _ = r.key(0)
            cols = [r.agg(a) for a in range(len(aggs))]
            d2h = k.nbytes + r.n_groups + sum(c.nbytes for c in cols)
            r.close()
        e2e_step()
        barrier()
        e0.record(stream)
        for _ in range(args.e2e_steps):
            e2e_step()
        e1.record(stream)
        barrier()
        ems = e0.elapsed_time(e1) / args.e2e_steps
        if dist is not None:
            t = torch.tensor([ems], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        out["e2e"] = {"value": n * world / (ems * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": n * 16 + nb, "d2h_bytes_per_step": int(d2h),
                      "ms_per_step": ems, "steps": args.e2e_steps, "note": "pdrs_groupby_agg on pinned host columns; PCIe-bound"}
        for p in (hk, hv, hn):
            ctx.host_free(p)

    # ---- the other configs, a few steps each (explanatory; not the headline)
    if not args.no_extras and world == 1:
        out["extras"] = extras(ctx, pb, args, n, keys, vals, peak)

    # ---- the reference's CPU algorithm on this box's host cores, bounded sample
    if rank == 0 and not args.no_cpu:
        import oracle as orc
        orc.build()
        threads = os.cpu_count() or 1
        cpu_reference_step(orc, 200_000, args.groups, threads)
        dt = cpu_reference_step(orc, args.cpu_rows, args.groups, threads)
        out["cpu_baseline"] = {"value": args.cpu_rows / dt, "unit": "rows/s", "cores": threads, "kind": "port",
                               "sample": f"{args.cpu_rows} rows, same generator/cardinality/NULLs; oracle restatement of grouping.rs + aggregation.rs (string keys, serial grouping, {threads}-thread aggregation)"}
    if rank == 0:
        print(json.dumps(out))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def extras(ctx, pb, args, n, keys, vals, peak):
    """Sum-only groupby, the 10M-group groupby and the inner / left join of BASELINE.json configs[1..2]."""
    ex = {}

    def timed(fn, reps=3):
        fn()
        best, kms = 1e30, 0.0
        for _ in range(reps):
            ctx.timer_begin()
            fn()
            ms = ctx.timer_end()
            if ms < best:
                best, kms = ms, ctx.stats()["main_kernel_ms"]
        return best, kms

    def gb(k, v, aggs):
        r = ctx.groupby_agg([k], [v], aggs)
        r.close()
    ms, kms = timed(lambda: gb(keys, vals, [(0, pb.SUM)]))
    ex["groupby_sum_1k"] = {"rows_per_s": n / (ms * 1e-3), "ms": ms, "kernel_ms": kms, "roofline_frac": ALG_BYTES_PER_ROW * n / (kms * 1e-3) / 1e9 / peak}
    try:
        k10 = ctx.synth_keys(n, card=10_000_000, seed=7)
        ms, kms = timed(lambda: gb(k10, vals, [(0, op) for op in (pb.SUM, pb.MEAN, pb.MIN, pb.MAX, pb.COUNT, pb.STD)]), reps=2)
        ex["groupby_all6_10m"] = {"rows_per_s": n / (ms * 1e-3), "ms": ms, "kernel_ms": kms, "roofline_frac": (ALG_BYTES_PER_ROW * n + 56 * 1e7) / (ms * 1e-3) / 1e9 / peak}
        del k10
    except Exception as e:  # noqa: BLE001
        ex["groupby_all6_10m"] = {"error": str(e)[:200]}
    try:
        nb_, np_ = n // 10, n
        build = ctx.synth_join_keys(nb_, unique=True)
        probe = ctx.synth_join_keys(np_, domain=2 * nb_)
        for how, name in ((pb.INNER, "join_inner"), (pb.LEFT, "join_left")):
            m = [0]

            def jn():
                j = ctx.join_pairs(probe, build, how)
                m[0] = j.n
                j.close()
            ms, kms = timed(jn, reps=2)
            alg = 8 * (np_ + nb_) + 16 * m[0]      # index-pairs variant of SURVEY.md §8(d)
            ex[name] = {"rows_per_s": (np_ + nb_) / (ms * 1e-3), "ms": ms, "probe_kernel_ms": kms, "pairs": m[0], "roofline_frac": alg / (ms * 1e-3) / 1e9 / peak}
    except Exception as e:  # noqa: BLE001
        ex["join"] = {"error": str(e)[:200]}
    return ex


if __name__ == "__main__":
    main()
