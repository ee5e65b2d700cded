#!/usr/bin/env python
"""Benchmarks of the pandrs groupby-aggregate / hash-join hot path on N B200s, one process per GPU.

  python bench.py --gpus N --steps K --warmup W                     # headline: BASELINE.json configs[1], 1K groups
  python bench.py --metric join ...                                 # configs[2]: 1e9 x 1e8 inner join (index pairs)
  python bench.py --workload c5 --gpus 8                            # configs[4]: Q1-style filter -> groupby, 7.5e8 rows per GPU
  python bench.py --scaling strong --gpus 8                         # the same total rows split over the GPUs
  python bench.py --impl reference [--metric join]                  # the reference's CPU algorithm (oracle port) on host cores

Prints ONE JSON line (rank 0).  A step = one pass of the operator over one batch of synthetic rows resident in HBM;
`e2e` = the same call on HOST columns (pageable memory, copies inside the timed region).  For N > 1 the step is the collective
operator behind the C ABI (pdrs_groupby_agg_dist / pdrs_join_pairs_dist: NCCL inside the library); torch.distributed only
launches the ranks, carries the 128-byte NCCL id once, and provides the timing barrier.
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
    ap.add_argument("--metric", default="groupby", choices=["groupby", "join"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"], help="groupby metric: configs[1] (1K groups) or configs[4] (Q1-style, 5 value columns)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU (weak) or in total (strong); 0 = the configuration's size")
    ap.add_argument("--groups", type=int, default=1000)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the bounded CPU-baseline sample (0 = default per metric)")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--extras-scale", type=float, default=1.0, help="scales the row counts of the configs[3..4] extras (debugging)")
    ap.add_argument("--dist-join", action="store_true", help="groupby metric, N > 1: also time the sharded inner join")
    return ap.parse_args()


def log(msg):
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


# ---------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock, power and throttle reasons sampled through NVML every 25 ms from before the warm-up (a thread of this process;
    `nvidia-smi -lms 20` as a child process was measured to stall every driver call of an allocation-heavy step by one polling
    period).  The summary uses the samples inside the timed region (mark_begin / mark_end), or, if the region was shorter than
    one sample, the samples under load.  Falls back to nvidia-smi at 200 ms when NVML cannot be loaded."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index=0):
        self.samples, self.proc, self.index, self.t0, self.t1 = [], None, index, None, None
        self.stop_flag, self.thread, self.how = False, None, None
        self.period = 0.025          # seconds between samples; timed_region() stretches it to 10 steps for long, driver-call-heavy steps

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                # An NVML query takes 1 - 11 ms on these boxes and holds a driver lock that CUDA calls of this process (allocations,
                # launches) wait for: the sampler keeps its duty cycle under 5% and reports its slowest query
                while not self.stop_flag:
                    t = time.perf_counter()
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        rs = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        self.samples.append((time.perf_counter(), sm, mx, {k for k, b in self.BITS.items() if rs & b}))
                    except Exception:
                        pass
                    took = time.perf_counter() - t
                    self.query_ms = max(getattr(self, "query_ms", 0.0), took * 1e3)
                    time.sleep(max(self.period, 20.0 * took))
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            self.how = "NVML, 25 ms"
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            self.how = "nvidia-smi, 200 ms"
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) < 7:
                continue
            try:
                sm, mx = float(f[0]), float(f[1])
            except ValueError:
                continue
            rs = {name for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]) if v.lower().startswith("active")}
            self.samples.append((time.perf_counter(), sm, mx, rs))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self.stop_flag = True
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()

        def parse(rows):
            sm, mx, reasons = [], 0, set()
            for _, s, m, rs in rows:
                sm.append(s); mx = max(mx, m); reasons |= rs
            return sm, mx, reasons
        inside = [x for x in self.samples if self.t0 is not None and self.t1 is not None and self.t0 <= x[0] <= self.t1 + 0.03]
        sm, mx, reasons = parse(inside)
        where = "timed region"
        if not sm:
            sm, mx, reasons = parse(self.samples)
            where = "whole run (timed region shorter than one sample)"
        sm.sort()
        hi = [x for x in sm if x > 0.5 * mx] or sm
        return {"sm_mhz": hi[len(hi) // 2] if hi else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm), "window": where, "sampler": (self.how or "").replace("25 ms", "%d ms" % round(self.period * 1e3)),
                "sampler_query_ms_max": round(getattr(self, "query_ms", 0.0), 3)}


# ---------------------------------------------------------------- the reference arm / CPU baselines (oracle/ = the checker, timed here only)
def cpu_groupby_step(orc, n, groups, threads, seed=42):
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


def cpu_groupby_idealised(orc, n, groups, threads, seed=42):
    """SURVEY.md §8(d) 'idealised CPU' line: typed i64 keys in per-thread flat hash tables, merged at the end - NOT the
    reference's algorithm, reported beside it so that the comparison does not only flatter the GPU.  rows/s."""
    k = orc.synth_keys(n, seed=seed, card=groups)
    v = orc.synth_vals(n, seed=seed)
    vn = orc.synth_nulls(n, seed=seed, per_million=NULL_PER_MILLION)
    orc.ideal_groupby(k[: n // 8], v[: n // 8], vn, nthreads=threads)
    t0 = time.perf_counter()
    _, ng = orc.ideal_groupby(k, v, vn, nthreads=threads)
    dt = time.perf_counter() - t0
    assert ng == min(groups, n) or n < 50 * groups
    return n / dt


def cpu_join_step(orc, n_probe, n_build, seed=42):
    """join_impl restated (join.rs:107-208: to_string() keys, HashMap<String, Vec<usize>>, serial build and probe) on a
    configs[2]-shaped sample: unique build keys, ~50% hits.  Returns (seconds, pairs)."""
    rk = orc.Col(orc.I64, orc.synth_join_keys(n_build, seed=seed, unique=True))
    lk = orc.Col(orc.I64, orc.synth_join_keys(n_probe, seed=seed, domain=2 * n_build))
    t0 = time.perf_counter()
    li, _ = orc.join(lk, rk, orc.INNER)
    return time.perf_counter() - t0, len(li)


def cpu_join_idealised(orc, n_probe, n_build, threads):
    """typed keys, per-thread build tables, all cores (oracle/typed_oracle.cpp) - not the reference's algorithm.  rows/s."""
    t0 = time.perf_counter()
    orc.typed_join(left_synth=dict(n=n_probe, domain=2 * n_build), right_synth=dict(n=n_build, unique=True), how=orc.INNER, nthreads=threads)
    return (n_probe + n_build) / (time.perf_counter() - t0)


def cpu_baseline_groupby(orc, args, threads, sample_rows):
    """The reference port at BASELINE.json configs[0] exactly (1M rows, 1K keys) and at 1e7 rows, measured; 1e8 / 1e9 rows
    extrapolated linearly from the largest measured size and labelled so."""
    cpu_groupby_step(orc, 200_000, args.groups, threads)
    sizes = {}
    for n in sorted({1_000_000, sample_rows, 10_000_000}):
        if n > sample_rows and n > 10_000_000:
            continue
        dt = cpu_groupby_step(orc, n, args.groups, threads)
        sizes[str(n)] = {"rows_per_s": n / dt, "seconds": dt, "how": "measured"}
    top = max(int(k) for k in sizes)
    for n in (100_000_000, 1_000_000_000):
        sizes[str(n)] = {"rows_per_s": sizes[str(top)]["rows_per_s"], "seconds": n / sizes[str(top)]["rows_per_s"], "how": f"extrapolated linearly from {top} rows"}
    main = sizes[str(sample_rows)]
    return {"value": main["rows_per_s"], "unit": "rows/s", "cores": threads, "kind": "port",
            "sample": f"{sample_rows} rows, same generator / cardinality / 5% NULLs as the GPU arm; oracle restatement of grouping.rs + aggregation.rs (string keys, serial grouping, {threads}-thread aggregation over groups like par_aggregate)",
            "sizes": sizes,
            # not the reference's algorithm: typed keys, per-thread flat tables (SURVEY.md §8d "idealised CPU"), 8x the sample
            "idealised_typed_key_rows_per_s": cpu_groupby_idealised(orc, 8 * sample_rows, args.groups, threads)}


def cpu_baseline_join(orc, threads, n_probe):
    cpu_join_step(orc, 100_000, 10_000)
    sizes = {}
    for npr in sorted({1_000_000, n_probe, 10_000_000}):
        if npr > max(n_probe, 10_000_000):
            continue
        dt, m = cpu_join_step(orc, npr, npr // 10)
        sizes[f"{npr}x{npr // 10}"] = {"rows_per_s": (npr + npr // 10) / dt, "seconds": dt, "pairs": m, "how": "measured"}
    top = max(sizes, key=lambda k: int(k.split("x")[0]))
    for npr in (100_000_000, 1_000_000_000):
        sizes[f"{npr}x{npr // 10}"] = {"rows_per_s": sizes[top]["rows_per_s"], "seconds": 1.1 * npr / sizes[top]["rows_per_s"], "how": f"extrapolated linearly from {top}"}
    main = sizes[f"{n_probe}x{n_probe // 10}"]
    return {"value": main["rows_per_s"], "unit": "rows/s", "cores": 1, "kind": "port",
            "sample": f"{n_probe} probe x {n_probe // 10} build rows (unique i64 keys, ~50% hits): oracle restatement of join_impl (join.rs:107-208): to_string() keys, HashMap<String, Vec<usize>>, serial build and probe - the reference join is single-threaded",
            "sizes": sizes,
            "idealised_typed_key_rows_per_s": cpu_join_idealised(orc, 8 * n_probe, 8 * n_probe // 10, threads), "idealised_cores": threads}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as orc
    orc.build()
    threads = os.cpu_count() or 1
    if args.metric == "join":
        n = args.cpu_rows or 4_000_000
        cpu_join_step(orc, 100_000, 10_000)
        ts = [cpu_join_step(orc, n, n // 10)[0] for _ in range(max(1, args.steps))]
        dt = sum(ts) / len(ts)
        val = (n + n // 10) / dt
        base = cpu_baseline_join(orc, threads, n)
        base["value"] = val
        cfg = join_config(n, n // 10, 1, "weak")
        metric = "join_rows_per_s"
    else:
        n = args.cpu_rows or 4_000_000
        for _ in range(max(1, min(args.warmup, 1))):
            cpu_groupby_step(orc, min(n, 200_000), args.groups, threads)
        ts = [cpu_groupby_step(orc, n, args.groups, threads) for _ in range(max(1, args.steps))]
        dt = sum(ts) / len(ts)
        val = n / dt
        base = {"value": val, "unit": "rows/s", "cores": threads, "kind": "port",
                "sample": f"{n} rows per step (same generator / key cardinality / 5% NULLs as the GPU arm), grouping serial like grouping.rs:62-104, aggregation over groups on {threads} threads like aggregation.rs:81",
                "idealised_typed_key_rows_per_s": cpu_groupby_idealised(orc, 8 * n, args.groups, threads)}
        cfg = groupby_config(args, n, "weak")
        metric = "groupby_agg_rows_per_s"
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": val, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if args.metric == "groupby" else "int64",
        "data": "synthetic", "config": cfg, "cpu_baseline": base,
        "e2e": {"value": val, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def groupby_config(args, rows, scaling):
    return {"workload": f"groupby-agg {rows} rows/GPU, i64 key {args.groups} distinct, sum/mean/min/max/count/std of f64, 5% nulls (BASELINE.json configs[1])",
            "rows_per_gpu": rows, "groups": args.groups, "aggs": "sum,mean,min,max,count,std", "null_fraction": 0.05, "scaling": scaling,
            "cache": "inputs (16 B/row) larger than L2; no flush needed"}


def join_config(n_probe, n_build, world, scaling):
    return {"workload": f"inner hash join, {n_probe} probe x {n_build} build rows per GPU, unique i64 keys, ~50% hits, index pairs (BASELINE.json configs[2])",
            "probe_rows_per_gpu": n_probe, "build_rows_per_gpu": n_build, "scaling": scaling,
            "cache": "inputs larger than L2; no flush needed"}


def c5_config(rows):
    return {"workload": f"Q1-style filter -> groupby(returnflag, linestatus), {rows} rows/GPU, 2 dictionary keys, 5 f64 value columns, Boolean mask 98% true: 4 sums, 3 means, count (BASELINE.json configs[4])",
            "rows_per_gpu": rows, "groups": 6, "cache": "inputs (40.125 B/row) larger than L2; no flush needed"}


# ---------------------------------------------------------------- helpers of the CUDA arm
class Env:
    pass


def timed_region(env, step, warmup, steps, stats_key="main_kernel_ms"):
    """W warm-up steps, then exactly K steps between barrier + synchronize, CUDA events on the launching stream, max over
    ranks.  Returns (ms per step, kernel ms per step [stats], launches per step, clocks)."""
    import torch
    clocks = ClockSampler(env.local)
    if env.rank == 0 and not os.environ.get("PDRS_BENCH_NO_CLOCKS"):
        clocks.start()
    tw = 0.0
    for _ in range(warmup):
        tw = time.perf_counter()
        step()
        tw = time.perf_counter() - tw
    clocks.period = min(0.25, max(0.025, 10.0 * tw))       # about one sample per 10 steps: a host-bound step can run into the query's driver lock
    launches0 = env.ctx.stats()["kernel_launches"]
    env.barrier()
    clocks.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms, walls = [], []
    e0.record(env.stream)
    for _ in range(steps):
        tw = time.perf_counter()
        step()
        kms.append(env.ctx.stats()[stats_key])
        walls.append((time.perf_counter() - tw) * 1e3)
    e1.record(env.stream)
    env.barrier()
    clocks.mark_end()
    ms = env.max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop() if env.rank == 0 else None
    launches = (env.ctx.stats()["kernel_launches"] - launches0) // max(steps, 1)
    env.step_walls = sorted(walls)
    if walls:
        log("library CUDA-event time per step (ms): " + " ".join("%.1f" % k for k in kms[:24]))
        log("host wall clock per step (ms): min %.3f  median %.3f  max %.3f   [%s]" % (min(walls), sorted(walls)[len(walls) // 2], max(walls), " ".join("%.1f" % w for w in walls[:24])))
    return ms / steps, sum(kms) / max(len(kms), 1), int(launches), clk


def masked_sum_on_device(ctx, torch, vals_ptr, nulls_ptr, n, chunk=1 << 27):
    """sum of the non-NULL f64 values (+ their count) with torch reductions over device-to-device copies of the column: an
    independent check of the aggregation result at full size."""
    dev = torch.device("cuda", ctx.device)
    tot, cnt = 0.0, 0
    bit = torch.arange(8, device=dev, dtype=torch.uint8)
    for off in range(0, n, chunk):
        m = min(chunk, n - off)
        v = torch.empty(m, dtype=torch.float64, device=dev)
        torch.cuda.synchronize(dev)
        ctx.memcpy(v.data_ptr(), vals_ptr + 8 * off, 8 * m, 2)
        if nulls_ptr:
            nb = (m + 7) // 8
            b = torch.empty(nb, dtype=torch.uint8, device=dev)
            ctx.memcpy(b.data_ptr(), nulls_ptr + off // 8, nb, 2)
            keep = (((b[:, None] >> bit[None, :]) & 1) == 0).reshape(-1)[:m]
            tot += float((v * keep).sum().item())
            cnt += int(keep.sum().item())
        else:
            tot += float(v.sum().item())
            cnt += m
    return tot, cnt


# ---------------------------------------------------------------- groupby headline (configs[1])
def bench_groupby(env, args, pb, out_extra):
    import numpy as np
    import torch
    ctx, world, rank = env.ctx, env.world, env.rank
    total = args.rows or 1_000_000_000
    n = total // world if args.scaling == "strong" else total
    ops = [pb.SUM, pb.MEAN, pb.MIN, pb.MAX, pb.COUNT, pb.STD]
    aggs = [(0, op) for op in ops]
    keys = ctx.synth_keys(n, card=args.groups, row0=rank * n)
    vals = ctx.synth_vals(n, row0=rank * n, null_per_million=NULL_PER_MILLION)
    ctx.sync()

    # up to groups_cap (4096) groups per rank every rank returns ALL groups (one all-gather of the states); beyond that the states are
    # exchanged by hash(key) mod ranks (one all-to-all) and every rank returns the groups it owns
    sharded = world > 1 and args.groups > 4000
    env.sharded_result = sharded

    def run(k=keys, v=vals):
        if world > 1:
            return env.comm.groupby_agg([k], [v], aggs, result_mode=pb.Comm.SHARDED if sharded else pb.Comm.REPLICATED)
        return ctx.groupby_agg([k], [v], aggs)

    def step():
        run().close()
    ms_per_step, k_ms, launches, clk = timed_region(env, step, args.warmup, args.steps)
    value = n * world / (ms_per_step * 1e-3)
    peak, peaks = env.peak, env.peaks
    achieved = ALG_BYTES_PER_ROW * n / (k_ms * 1e-3) / 1e9
    algo = ctx.stats()["groupby_algo_used"]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": {pb.GB_TILESORT: "gb_tsort_kernel", pb.GB_SHARED: "gb_shared_kernel", pb.GB_FEW: "gb_few_kernel"}.get(algo, "gb_global_kernel"),
                "kernel_ms": k_ms, "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6.65 TB/s (of fallback)",
                "alg_bytes_per_row": ALG_BYTES_PER_ROW}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            tr = json.load(open(traffic_file))
            roofline["traffic"] = tr.get(roofline["kernel"] + "_bytes_per_row", 0) * n or None
            roofline["traffic_source"] = "ncu dram__bytes_read + write of this kernel from profiles/ (not measured in this run)"
        except Exception:
            pass
    out = {"metric": "groupby_agg_rows_per_s", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": groupby_config(args, n, args.scaling), "roofline": roofline, "gpu_launches": launches, "clocks": clk}
    if world > 1:
        ex_ms, ex_bytes = env.comm.last_exchange()
        nvl = float(env.peaks.get("nvlink_gbs", 770.0))
        out["exchange"] = {"kind": ("all-to-all (ncclSend / ncclRecv) of per-group states by hash(key) mod ranks + merge by the owner" if sharded else
                                    "ncclAllGather of per-group states (fixed size) + merge kernel") + ", inside pdrs_groupby_agg_dist",
                           "ms": ex_ms, "bytes_to_peers_per_gpu": ex_bytes, "nvlink_gbs_achieved": ex_bytes / (ex_ms * 1e-3) / 1e9 if ex_ms else None, "nvlink_peak_gbs": nvl}
    log(f"headline: {ms_per_step:.3f} ms/step (kernel {k_ms:.3f} ms)")

    # ---- the result of one more step, checked outside the timed region
    if not args.no_parity:
        out["parity_check"] = parity_groupby(env, pb, run, keys, vals, n, args, ops)
        log(f"parity_check: {out['parity_check'].get('ok')}")

    # ---- end to end through the C ABI with HOST buffers: H2D of the inputs and D2H of the result inside the timed region
    if not args.no_e2e:
        out["e2e"] = e2e_groupby(env, pb, keys, vals, n, aggs, args)
        log("e2e done")
    if not args.no_extras and world == 1:
        out["extras"] = extras(ctx, pb, args, n, keys, vals, peak)
    if args.dist_join and world > 1:
        try:
            out["dist_join"] = dist_join(env, pb, n)
        except Exception as e:  # noqa: BLE001
            out["dist_join"] = {"error": str(e)[:300]}
    return out


def parity_groupby(env, pb, run, keys, vals, n, args, ops):
    """(1) invariants of the full-size result: group count, sum of the group sizes = N * n, valid counts = non-NULL rows;
    (2) sum over the groups' sums == a torch reduction over all value columns (independent code), all ranks;
    (3) full oracle parity (typed oracle, bit-exact keys / counts / min / max, 1e-12 relative on sum / mean / std) of the same
        operator on the first 4M rows of rank 0's shard."""
    import numpy as np
    import torch
    import oracle as orc
    ctx, world, rank = env.ctx, env.world, env.rank
    rec = {}
    r = run()
    try:
        rows = r.group_rows()
        nv = r.valid_n(0)
        sums = r.agg(0)
        G = r.n_groups
    finally:
        r.close()
    ref_sum, ref_cnt = masked_sum_on_device(ctx, torch, vals.ptr, vals.nulls_ptr, n)
    t = torch.tensor([ref_sum, float(ref_cnt)], dtype=torch.float64, device="cuda")
    if env.dist is not None:
        env.dist.all_reduce(t)
    ref_sum, ref_cnt = float(t[0].item()), int(t[1].item())
    G, rows_total, valid_total, sum_total = int(G), int(rows.sum()), int(nv.sum()), float(sums.sum())
    if getattr(env, "sharded_result", False):      # every rank holds the groups it owns: the invariants are sums over the ranks
        ti = torch.tensor([G, rows_total, valid_total], dtype=torch.int64, device="cuda")
        tf = torch.tensor([sum_total], dtype=torch.float64, device="cuda")
        env.dist.all_reduce(ti)
        env.dist.all_reduce(tf)
        G, rows_total, valid_total, sum_total = int(ti[0].item()), int(ti[1].item()), int(ti[2].item()), float(tf[0].item())
        rec["result"] = "sharded by hash(key) mod ranks"
    rec["groups"] = G
    rec["rows_total"] = rows_total
    rec["valid_total"] = valid_total
    rec["sum_rel_err_vs_torch_reduction"] = abs(sum_total - ref_sum) / max(abs(ref_sum), 1e-300)
    ok = G == min(args.groups, n * world) and rec["rows_total"] == n * world and rec["valid_total"] == ref_cnt and rec["sum_rel_err_vs_torch_reduction"] < 1e-9
    m = min(n, 4_000_000)
    if rank == 0:
        orc.build()
        kp = pb.Column(pb.I64, device_ptr=keys.ptr, length=m, owner=keys)
        vp = pb.Column(pb.F64, device_ptr=vals.ptr, nulls_ptr=vals.nulls_ptr, null_len=(m + 7) // 8, length=m, owner=vals)
        rr = ctx.groupby_agg([kp], [vp], [(0, op) for op in ops])
        try:
            gk, _ = rr.key(0)
            order = np.argsort(gk)
            got = {"rows": rr.group_rows()[order], "aggs": [rr.agg(a)[order] for a in range(len(ops))]}
        finally:
            rr.close()
        tg = orc.typed_groupby_synth(m, card=args.groups, null_per_million=NULL_PER_MILLION, row0=0)
        to = np.argsort(tg["keys"][0][0].view(np.int64))
        names = {pb.SUM: "sum", pb.MEAN: "mean", pb.MIN: "min", pb.MAX: "max", pb.STD: "std"}
        pok = np.array_equal(gk[order], tg["keys"][0][0].view(np.int64)[to]) and np.array_equal(got["rows"], tg["group_rows"][to])
        worst = 0.0
        for a, op in enumerate(ops):
            if op == pb.COUNT:
                pok = pok and np.array_equal(got["aggs"][a], tg["group_rows"][to].astype(np.float64))
                continue
            w = tg[names[op]][to]
            if op in (pb.MIN, pb.MAX):
                pok = pok and np.array_equal(got["aggs"][a], w)
            else:
                rel = float((np.abs(got["aggs"][a] - w) / np.maximum(np.abs(w), 1e-300)).max())
                worst = max(worst, rel)
                pok = pok and rel <= 1e-12
        rec["oracle_prefix"] = {"rows": m, "ok": bool(pok), "worst_rel_err_f64": worst, "what": "typed oracle (bit-identical to the reference restatement): keys / counts / min / max bit-exact, sum / mean / std within 1e-12 relative"}
        ok = ok and pok
    rec["ok"] = bool(ok)
    return rec


def e2e_groupby(env, pb, keys, vals, n, aggs, args):
    """pdrs_groupby_agg on HOST columns in ordinary pageable memory (what a Rust Arc<[i64]> is), every step: H2D of the inputs +
    the operator + D2H of the result.  The pinned-memory variant is reported beside it."""
    import numpy as np
    import torch
    ctx, world = env.ctx, env.world
    nb = (n + 7) // 8
    hk = np.empty(n, np.int64); hv = np.empty(n, np.float64); hn = np.empty(nb, np.uint8)
    ctx.memcpy(hk.ctypes.data, keys.ptr, n * 8, 1)
    ctx.memcpy(hv.ctypes.data, vals.ptr, n * 8, 1)
    ctx.memcpy(hn.ctypes.data, vals.nulls_ptr, nb, 1)
    res = {}
    try:
        avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
    except Exception:
        avail = 1 << 62
    if avail < 2.5 * world * (n * 16 + nb):      # every rank of this box holds a host copy of its shard
        return {"value": None, "unit": "rows/s", "h2d_bytes_per_step": n * 16 + nb, "d2h_bytes_per_step": 0, "note": f"skipped: {avail >> 30} GiB of host memory available for {world} ranks x {(n * 16 + nb) >> 30} GiB"}
    for kind in (("pageable", "pinned") if world <= 2 else ("pageable",)):
        if kind == "pageable":
            hkeys, hvals, held = pb.Column(pb.I64, hk), pb.Column(pb.F64, hv, hn), None
        else:
            pk, pv, pn = ctx.host_alloc(n * 8), ctx.host_alloc(n * 8), ctx.host_alloc(nb)
            ctx.memcpy(pk, keys.ptr, n * 8, 1); ctx.memcpy(pv, vals.ptr, n * 8, 1); ctx.memcpy(pn, vals.nulls_ptr, nb, 1)
            hkeys = pb.Column(pb.I64, device_ptr=pk, length=n)
            hvals = pb.Column(pb.F64, device_ptr=pv, nulls_ptr=pn, null_len=nb, length=n)
            hkeys.mem = hvals.mem = pb.MEM_HOST
            held = (pk, pv, pn)
        d2h = [0]

        def e2e_step():
            r = env.comm.groupby_agg([hkeys], [hvals], aggs, result_mode=pb.Comm.REPLICATED) if world > 1 else ctx.groupby_agg([hkeys], [hvals], aggs)
            k, _ = r.key(0)
            cols = [r.agg(a) for a in range(len(aggs))]
            d2h[0] = k.nbytes + r.n_groups + sum(c.nbytes for c in cols)
            r.close()
        e2e_step()
        env.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(env.stream)
        for _ in range(args.e2e_steps):
            e2e_step()
        e1.record(env.stream)
        env.barrier()
        wall = (time.perf_counter() - t0) * 1e3 / args.e2e_steps
        ems = env.max_over_ranks(max(e0.elapsed_time(e1) / args.e2e_steps, wall))     # pageable copies block the host: the wall clock is the honest one
        res[kind] = {"value": n * world / (ems * 1e-3), "ms_per_step": ems}
        if held:
            for p in held:
                ctx.host_free(p)
    return {"value": res["pageable"]["value"], "unit": "rows/s", "h2d_bytes_per_step": n * 16 + nb, "d2h_bytes_per_step": int(d2h[0]),
            "ms_per_step": res["pageable"]["ms_per_step"], "steps": args.e2e_steps, "host_memory": "pageable (numpy arrays = what a Rust Arc<[T]> is)",
            "pinned": res.get("pinned"), "note": "pdrs_groupby_agg on host columns, chunked: 2^26-row chunks travel through the staging engine (8 threads, pinned 8 MB slots; pinned sources by direct DMA) while the previous chunk is aggregated; PCIe-bound"}


# ---------------------------------------------------------------- join metric (configs[2])
def bench_join(env, args, pb):
    import torch
    ctx, world, rank = env.ctx, env.world, env.rank
    total = args.rows or 1_000_000_000
    np_ = total // world if args.scaling == "strong" else total
    nb_ = np_ // 10
    build = ctx.synth_join_keys(nb_, unique=True, row0=rank * nb_)
    probe = ctx.synth_join_keys(np_, domain=2 * nb_ * world, row0=rank * np_)
    ctx.sync()
    pairs = [0]

    def run(how=pb.INNER):
        if world > 1:
            return env.comm.join_pairs(probe, build, how, rank * np_, rank * nb_, np_, nb_, nb_ * world)
        return ctx.join_pairs(probe, build, how)

    def step():
        j = run()
        pairs[0] = j.n
        j.close()
    ms_per_step, _, launches, clk = timed_region(env, step, args.warmup, args.steps, stats_key="total_ms")
    m_total = env.sum_over_ranks(pairs[0])
    value = (np_ + nb_) * world / (ms_per_step * 1e-3)
    alg = 8.0 * (np_ + nb_) + 16.0 * m_total / world        # per GPU: keys of both sides read once + i64 index pairs written once
    peak = env.peak
    achieved = alg / (ms_per_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "kernel": "the whole pdrs_join_pairs call: jpart1 (build side, probe side) + join_build + join_probe_emit - no single dominant kernel",
                "kernel_ms": ms_per_step, "alg_bytes": alg, "peak_source": "MEASURED_PEAKS.json (of measured)" if env.peaks else "fallback 6.65 TB/s (of fallback)",
                "note": "a radix-partitioned join moves ~40 GB for 16.8 GB of algorithmic traffic: its ceiling against this denominator is ~42% (DESIGN.md)"}
    out = {"metric": "join_rows_per_s", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "int64", "data": "synthetic",
           "config": join_config(np_, nb_, world, args.scaling), "roofline": roofline, "gpu_launches": launches, "clocks": clk, "pairs_per_step": int(m_total)}
    if world > 1:
        ex_ms, ex_bytes = env.comm.last_exchange()
        nvl = float(env.peaks.get("nvlink_gbs", 770.0))
        out["exchange"] = {"kind": "partition kernel storing (key, row) runs straight into the peers' receive areas (CUDA IPC over NVLink)", "ms": ex_ms,
                           "bytes_to_peers_per_gpu": ex_bytes, "nvlink_gbs_achieved": ex_bytes / (ex_ms * 1e-3) / 1e9 if ex_ms else None, "nvlink_peak_gbs": nvl,
                           "nvlink_frac": ex_bytes / (ex_ms * 1e-3) / 1e9 / nvl if ex_ms else None}
    log(f"join headline: {ms_per_step:.3f} ms/step")
    if not args.no_parity:
        out["parity_check"] = parity_join(env, pb, run, np_, nb_)
        log(f"parity_check: {out['parity_check'].get('ok')}")
    if not args.no_e2e and world == 1:
        out["e2e"] = e2e_join(env, pb, probe, build, np_, nb_, args)
    if not args.no_extras and world == 1:
        ex = {}
        try:
            for name, fn in (("left_pairs", lambda: run(pb.LEFT)), ("inner_2payload", None)):
                if fn is None:
                    p1, p2 = ctx.synth_keys(nb_, card=1 << 40, seed=9), ctx.synth_vals(nb_, seed=9)
                    fn = lambda: ctx.join_gather(probe, build, pb.INNER, [p1, p2])      # noqa: E731
                fn().close()
                best, m = 1e30, 0
                for _ in range(2):
                    ctx.timer_begin()
                    j = fn()
                    m = j.n
                    j.close()
                    best = min(best, ctx.timer_end())
                algb = 8.0 * (np_ + nb_) + 16.0 * m if name != "inner_2payload" else 8.0 * np_ + nb_ * 24.0 + m * 32.0
                ex[name] = {"ms": best, "pairs": m, "rows_per_s": (np_ + nb_) / (best * 1e-3), "alg_bytes": algb, "roofline_frac": algb / (best * 1e-3) / 1e9 / peak}
        except Exception as e:  # noqa: BLE001
            ex["error"] = str(e)[:200]
        out["extras"] = ex
    return out


def parity_join(env, pb, run, np_, nb_):
    """The pairs of one more step against the typed oracle (oracle/typed_oracle.cpp) on the same generator-backed keys:
    count, order-independent 64-bit checksum of the (left, right) multiset, index sums - Inner and Left.  N > 1: the ranks'
    shards of the result are summed (every pair is returned by exactly one rank)."""
    import torch
    import oracle as orc
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _util import device_pair_stats
    ctx, world, rank = env.ctx, env.world, env.rank
    mask = (1 << 64) - 1
    got = {}
    for how, name in ((pb.INNER, "inner"), (pb.LEFT, "left")):
        j = run(how)
        try:
            n, cs, sl, sr, un = device_pair_stats(ctx, j)
        finally:
            j.close()
        t = torch.tensor([n, cs - (1 << 64) if cs >= (1 << 63) else cs, sl, sr, un], dtype=torch.int64, device="cuda")
        if env.dist is not None:
            env.dist.all_reduce(t)
        v = [int(x) for x in t.tolist()]
        got[name] = (v[0], v[1] & mask, v[2], v[3], v[4])
    rec = {"inner_pairs": got["inner"][0], "left_pairs": got["left"][0]}
    ok = True
    if rank == 0 and np_ * world > 2_500_000_000:
        ok = got["left"][0] == np_ * world and got["inner"][4] == 0 and got["left"][0] - got["left"][4] == got["inner"][0]
        rec["what"] = "invariants only (|left| = probe rows, inner = matched left pairs): the host oracle would need minutes for this many probe rows; tests/test_gpu_scale.py and the N = 1 run compare with the oracle"
    elif rank == 0:
        orc.build()
        want = orc.typed_join(left_synth=dict(n=np_ * world, domain=2 * nb_ * world), right_synth=dict(n=nb_ * world, unique=True), how=orc.LEFT)
        wl = (want["n"], want["checksum"], want["sum_left"], want["sum_right"], want["unmatched_left"])
        wi = (want["n"] - want["unmatched_left"], (want["checksum"] - want["checksum_unmatched"]) & mask, None, want["sum_right"], 0)
        ok = got["left"] == wl and got["inner"][0] == wi[0] and got["inner"][1] == wi[1] and got["inner"][3] == wi[3] and got["inner"][4] == 0
        rec["what"] = "count / 64-bit multiset checksum / index sums of the Inner and Left pairs == typed oracle on the same keys (all ranks' shards summed)"
    rec["ok"] = bool(ok)
    return rec


def e2e_join(env, pb, probe, build, np_, nb_, args):
    import numpy as np
    import torch
    ctx = env.ctx
    hp = np.empty(np_, np.int64); hb = np.empty(nb_, np.int64)
    ctx.memcpy(hp.ctypes.data, probe.ptr, np_ * 8, 1)
    ctx.memcpy(hb.ctypes.data, build.ptr, nb_ * 8, 1)
    cp, cb = pb.Column(pb.I64, hp), pb.Column(pb.I64, hb)
    d2h = [0]

    def step():
        j = ctx.join_pairs(cp, cb, pb.INNER)
        li, ri = j.indices()
        d2h[0] = li.nbytes + ri.nbytes
        j.close()
    step()
    steps = max(2, min(args.e2e_steps, 3))
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    env.barrier()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    return {"value": (np_ + nb_) / (ms * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": (np_ + nb_) * 8, "d2h_bytes_per_step": int(d2h[0]), "ms_per_step": ms, "steps": steps,
            "host_memory": "pageable (numpy arrays)", "note": "pdrs_join_pairs on host key columns + pdrs_join_indices copy-out of all pairs; PCIe-bound (wall clock)"}


# ---------------------------------------------------------------- configs[4]: Q1-style filter -> groupby, 5 value columns
def make_c5(ctx, pb, torch, n, seed):
    dev = torch.device("cuda", ctx.device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    CH = 1 << 26

    def uniform(lo, hi):
        out = torch.empty(n, dtype=torch.float64, device=dev)
        for a in range(0, n, CH):
            b = min(n, a + CH)
            out[a:b] = torch.rand(b - a, device=dev, generator=g, dtype=torch.float64) * (hi - lo) + lo
        return out
    rf = torch.empty(n, dtype=torch.int32, device=dev)
    ls = torch.empty(n, dtype=torch.int32, device=dev)
    ship = torch.empty(n, dtype=torch.int64, device=dev)
    for a in range(0, n, CH):
        b = min(n, a + CH)
        rf[a:b] = torch.randint(0, 3, (b - a,), device=dev, generator=g, dtype=torch.int32)
        ls[a:b] = torch.randint(0, 2, (b - a,), device=dev, generator=g, dtype=torch.int32)
        ship[a:b] = torch.randint(8000, 10600, (b - a,), device=dev, generator=g, dtype=torch.int64)
    qty, price, disc, tax = uniform(1, 50), uniform(900, 105000), uniform(0, 0.1), uniform(0, 0.08)
    disc_price = torch.empty_like(price)
    charge = torch.empty_like(price)
    for a in range(0, n, CH):
        b = min(n, a + CH)
        disc_price[a:b] = price[a:b] * (1 - disc[a:b])
        charge[a:b] = disc_price[a:b] * (1 + tax[a:b])
    del tax
    cutoff = 10_548                                    # ~98% of the rows have shipdate <= cutoff
    nbytes = (n + 63) // 64 * 8
    mask = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    wts = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.int32, device=dev)
    for a in range(0, n, CH):
        b = min(n, a + CH)
        f = torch.zeros((b - a + 7) // 8 * 8, dtype=torch.int32, device=dev)
        f[: b - a] = (ship[a:b] <= cutoff).to(torch.int32)
        mask[a // 8: a // 8 + f.numel() // 8] = (f.view(-1, 8) * wts).sum(1).to(torch.uint8)
    torch.cuda.current_stream(dev).synchronize()

    def col(dtype, t, length=None):
        return pb.Column(dtype, device_ptr=t.data_ptr(), length=int(length if length is not None else t.numel()), owner=t)
    keys = [col(pb.DICT_U32, rf), col(pb.DICT_U32, ls)]
    vals = [col(pb.F64, t) for t in (qty, price, disc_price, charge, disc)]
    aggs = [(0, pb.SUM), (1, pb.SUM), (2, pb.SUM), (3, pb.SUM), (0, pb.MEAN), (1, pb.MEAN), (4, pb.MEAN), (0, pb.COUNT)]
    tensors = dict(rf=rf, ls=ls, ship=ship, qty=qty, price=price, disc_price=disc_price, charge=charge, disc=disc, mask=mask)
    return keys, vals, aggs, col(pb.BOOL_BITS, mask, n), (col(pb.I64, ship), pb.CMP_LE, cutoff), tensors


def bench_c5(env, args, pb):
    import numpy as np
    import torch
    ctx, world, rank = env.ctx, env.world, env.rank
    total = args.rows or 750_000_000
    n = total // world if args.scaling == "strong" else total
    log("generating configs[4] columns")
    keys, vals, aggs, fmask, pred, T = make_c5(ctx, pb, torch, n, 4242 + rank)

    def run(filt=fmask, pr=None, kk=keys, vv=vals):
        if world > 1:
            return env.comm.groupby_agg(kk, vv, aggs, filter=filt, pred=pr, result_mode=pb.Comm.REPLICATED)
        return ctx.groupby_agg(kk, vv, aggs, filter=filt, pred=pr)

    def step():
        run().close()
    ms_per_step, k_ms, launches, clk = timed_region(env, step, args.warmup, args.steps)
    value = n * world / (ms_per_step * 1e-3)
    bpr = 8 + 5 * 8 + 0.125
    achieved = bpr * n / (k_ms * 1e-3) / 1e9
    algo = ctx.stats()["groupby_algo_used"]
    out = {"metric": "groupby_agg_rows_per_s", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": c5_config(n), "gpu_launches": launches, "clocks": clk,
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": env.peak, "unit": "GB/s", "frac": achieved / env.peak, "traffic": None,
                        "kernel": "gb_few_kernel" if algo == pb.GB_FEW else f"algo {algo}", "kernel_ms": k_ms, "alg_bytes_per_row": bpr,
                        "peak_source": "MEASURED_PEAKS.json (of measured)" if env.peaks else "fallback 6.65 TB/s (of fallback)"}}
    if world > 1:
        ex_ms, ex_bytes = env.comm.last_exchange()
        out["exchange"] = {"kind": "ncclAllGather of the 6 groups' states + merge", "ms": ex_ms, "bytes_to_peers_per_gpu": ex_bytes}
    # the date predicate evaluated in the scan instead of the precomputed mask (48.125 B/row)
    try:
        run(None, pred).close()
        ctx.timer_begin()
        run(None, pred).close()
        pms = ctx.timer_end()
        out["predicate_in_scan"] = {"ms": env.max_over_ranks(pms), "kernel_ms": ctx.stats()["main_kernel_ms"], "alg_bytes_per_row": 56.0,
                                    "roofline_frac": 56.0 * n / (ctx.stats()["main_kernel_ms"] * 1e-3) / 1e9 / env.peak,
                                    "note": "shipdate <= cutoff evaluated from the i64 column inside the scan (pdrs_groupby_agg_where): 8 B keys + 40 B values + 8 B date"}
    except Exception as e:  # noqa: BLE001
        out["predicate_in_scan"] = {"error": str(e)[:200]}
    if not args.no_parity:
        # merged result vs torch reductions over all ranks' shards (independent code), plus oracle parity on a 2M-row prefix of rank 0
        import oracle as orc
        r = run()
        try:
            k0, _ = r.key(0); k1, _ = r.key(1)
            res = {(int(a), int(b)): (int(c), [float(r.agg(i)[g]) for i in range(len(aggs))]) for g, (a, b, c) in enumerate(zip(k0, k1, r.group_rows()))}
        finally:
            r.close()
        keep = T["ship"] <= pred[2]
        ref = torch.zeros((6, 3), dtype=torch.float64, device="cuda")
        gid = (T["rf"].to(torch.int64) * 2 + T["ls"].to(torch.int64))
        CH = 1 << 26
        for a in range(0, n, CH):
            b = min(n, a + CH)
            kk = keep[a:b]
            ref[:, 0].index_add_(0, gid[a:b][kk], torch.ones(int(kk.sum().item()), dtype=torch.float64, device="cuda"))
            ref[:, 1].index_add_(0, gid[a:b][kk], T["qty"][a:b][kk])
            ref[:, 2].index_add_(0, gid[a:b][kk], T["charge"][a:b][kk])
        if env.dist is not None:
            env.dist.all_reduce(ref)
        ref = ref.cpu().numpy()
        ok, worst = len(res) == 6, 0.0
        for (a, b), (cnt, ag) in res.items():
            g = a * 2 + b
            ok = ok and cnt == int(ref[g, 0]) and ag[7] == float(cnt)
            for got, want in ((ag[0], ref[g, 1]), (ag[3], ref[g, 2]), (ag[4], ref[g, 1] / max(ref[g, 0], 1))):
                rel = abs(got - want) / max(abs(want), 1e-300)
                worst = max(worst, rel)
                ok = ok and rel < 1e-9
        rec = {"groups": len(res), "rows_kept_total": int(ref[:, 0].sum()), "worst_rel_err_vs_torch_reduction": worst}
        if rank == 0:
            orc.build()
            m = min(n, 2_000_000)
            hk = [orc.Col(orc.DICT_U32, T["rf"][:m].cpu().numpy().view(np.uint32)), orc.Col(orc.DICT_U32, T["ls"][:m].cpu().numpy().view(np.uint32))]
            hf = orc.Col(orc.BOOL_BITS, T["mask"][: (m + 7) // 8].cpu().numpy(), length=m)
            pk = [pb.Column(c.dtype, device_ptr=c.ptr, length=m, owner=c) for c in keys]
            pv = [pb.Column(c.dtype, device_ptr=c.ptr, length=m, owner=c) for c in vals]
            rr = ctx.groupby_agg(pk, pv, aggs, filter=pb.Column(pb.BOOL_BITS, device_ptr=fmask.ptr, length=m, owner=fmask))
            try:
                g0, _ = rr.key(0); g1, _ = rr.key(1)
                gres = {(int(a), int(b)): [float(rr.agg(i)[g]) for i in range(len(aggs))] for g, (a, b) in enumerate(zip(g0, g1))}
            finally:
                rr.close()
            pok, pworst = True, 0.0
            for vi, name in ((0, "qty"), (1, "price"), (2, "disc_price"), (3, "charge"), (4, "disc")):
                tg = orc.typed_groupby(hk, orc.Col(orc.F64, T[name][:m].cpu().numpy()), filter=hf)
                for g in range(tg["n_groups"]):
                    kt = (int(tg["keys"][0][0][g]), int(tg["keys"][1][0][g]))
                    for ai, (v, op) in enumerate(aggs):
                        if v != vi:
                            continue
                        want = {pb.SUM: tg["sum"][g], pb.MEAN: tg["mean"][g], pb.COUNT: float(tg["group_rows"][g])}[op]
                        rel = abs(gres[kt][ai] - want) / max(abs(want), 1e-300)
                        pworst = max(pworst, rel)
                        pok = pok and (rel <= 1e-12 if op != pb.COUNT else rel == 0)
            rec["oracle_prefix"] = {"rows": m, "ok": bool(pok), "worst_rel_err_f64": pworst}
            ok = ok and pok
        rec["ok"] = bool(ok)
        out["parity_check"] = rec
    return out


# ---------------------------------------------------------------- extras of the groupby headline (N = 1)
def dist_join(env, pb, n_probe):
    """configs[2] sharded over the ranks through pdrs_join_pairs_dist (weak scaling)."""
    ctx, world, rank = env.ctx, env.world, env.rank
    np_, nb_ = n_probe, n_probe // 10
    build = ctx.synth_join_keys(nb_, unique=True, row0=rank * nb_)
    probe = ctx.synth_join_keys(np_, domain=2 * nb_ * world, row0=rank * np_)
    best = None
    for rep in range(3):
        env.barrier()
        t0 = time.perf_counter()
        j = env.comm.join_pairs(probe, build, pb.INNER, rank * np_, rank * nb_, np_, nb_, nb_ * world)
        local_ms = ctx.stats()["total_ms"]
        pairs = j.n
        j.close()
        wall = env.max_over_ranks((time.perf_counter() - t0) * 1e3)
        ex_ms, ex_bytes = env.comm.last_exchange()
        rec = (env.max_over_ranks(ex_ms), env.max_over_ranks(local_ms), wall, env.sum_over_ranks(pairs), ex_bytes)
        if rep and (best is None or rec[0] + rec[1] < best[0] + best[1]):
            best = rec
    sh, lo, wl, total_pairs, sent = best
    nvl = float(env.peaks.get("nvlink_gbs", 770.0))
    return {"rows_per_s": (np_ + nb_) * world / ((sh + lo) * 1e-3), "shuffle_ms": sh, "local_ms": lo, "wall_ms": wl, "pairs": total_pairs,
            "rows_per_gpu": np_ + nb_, "nvlink_bytes_per_gpu": sent, "nvlink_gbs_achieved": sent / (sh * 1e-3) / 1e9, "nvlink_peak_gbs": nvl,
            "nvlink_frac": sent / (sh * 1e-3) / 1e9 / nvl}


def extras(ctx, pb, args, n, keys, vals, peak):
    """Sum-only groupby, the 10M-group groupby, the inner / left join (pairs, and with 2 payload columns) of BASELINE.json
    configs[1..2], then configs[3..4] (extras_c4_c5)."""
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
    log("extras: sum only, 10M groups, joins")
    ms, kms = timed(lambda: gb(keys, vals, [(0, pb.SUM)]))
    ex["groupby_sum_1k"] = {"rows_per_s": n / (ms * 1e-3), "ms": ms, "kernel_ms": kms, "roofline_frac": ALG_BYTES_PER_ROW * n / (kms * 1e-3) / 1e9 / peak}
    # rows next to the path: the row lists of the groups (par_groupby, grouping.rs:124-331) and Median over them
    try:
        def rows():
            r = ctx.groupby_rows([keys])
            r.close()
        ms, _ = timed(rows, reps=2)
        ex["par_groupby_rows_1k"] = {"rows_per_s": n / (ms * 1e-3), "ms": ms, "alg_bytes_per_row": 16.0, "roofline_frac": 16.0 * n / (ms * 1e-3) / 1e9 / peak,
                                     "what": "pdrs_groupby_rows: keys in (8 B/row), ascending row ids per group out (8 B/row)"}
        m = min(n, 200_000_000)
        km, vm = pb.Column(pb.I64, device_ptr=keys.ptr, length=m), pb.Column(pb.F64, device_ptr=vals.ptr, nulls_ptr=vals.nulls_ptr, null_len=(m + 7) // 8, length=m)
        r = ctx.groupby_rows([km])
        try:
            ms, _ = timed(lambda: r.agg(vm, pb.MEDIAN), reps=2)
        finally:
            r.close()
        ex["median_1k_2e8_rows"] = {"rows_per_s": m / (ms * 1e-3), "ms": ms, "rows": m}
    except Exception as e:  # noqa: BLE001
        ex["par_groupby_rows_1k"] = {"error": str(e)[:200]}
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
        # configs[2] with its 2 build-side payload columns (i64, f64) materialised with the pairs (join.rs:290-552): pdrs_join_gather
        p1, p2 = ctx.synth_keys(nb_, card=1 << 40, seed=9), ctx.synth_vals(nb_, seed=9)
        m = [0]

        def jn_payload():
            j = ctx.join_gather(probe, build, pb.INNER, [p1, p2])
            m[0] = j.n
            j.close()
        log("extra join_inner_2payload ...")
        ms, _ = timed(jn_payload, reps=2)
        alg = 8 * np_ + nb_ * (8 + 16) + m[0] * (16 + 16)      # SURVEY.md §8(d), P = 2 payload columns
        ex["join_inner_2payload"] = {"rows_per_s": (np_ + nb_) / (ms * 1e-3), "ms": ms, "pairs": m[0], "alg_bytes": alg, "roofline_frac": alg / (ms * 1e-3) / 1e9 / peak}
        del p1, p2, build, probe
    except Exception as e:  # noqa: BLE001
        ex["join"] = {"error": str(e)[:200]}
    try:
        ex.update(extras_c4_c5(ctx, pb, peak, timed, args.extras_scale))
    except Exception as e:  # noqa: BLE001
        ex["c4_c5"] = {"error": str(e)[:200]}
    return ex


def extras_c4_c5(ctx, pb, peak, timed, scale=1.0):
    """BASELINE.json configs[3] (multi-key + dictionary key, 500M rows, Zipf 1.1) and one GPU's shard of configs[4]
    (Q1-style filter -> groupby(returnflag, linestatus), 7.5e8 rows).  Columns are generated with torch on the device
    (same stream as the context) and handed over as borrowed device pointers."""
    import torch
    dev = torch.device("cuda", ctx.device)
    g = torch.Generator(device=dev)
    g.manual_seed(4242)
    CH = 1 << 26

    def zipf(n, domain, dtype, s=1.1):
        w = torch.arange(1, domain + 1, device=dev, dtype=torch.float64).pow(-s)
        cdf = (w.cumsum(0) / w.sum()).to(torch.float32)
        out = torch.empty(n, dtype=dtype, device=dev)
        for a in range(0, n, CH):
            b = min(n, a + CH)
            out[a:b] = torch.searchsorted(cdf, torch.rand(b - a, device=dev, generator=g)).clamp_(max=domain - 1).to(dtype)
        return out

    def uniform(n, lo, hi):
        out = torch.empty(n, dtype=torch.float64, device=dev)
        for a in range(0, n, CH):
            b = min(n, a + CH)
            out[a:b] = torch.rand(b - a, device=dev, generator=g, dtype=torch.float64) * (hi - lo) + lo
        return out

    def col(dtype, t, n=None):
        return pb.Column(dtype, device_ptr=t.data_ptr(), length=int(n if n is not None else t.numel()), owner=t)

    ALL6 = (pb.SUM, pb.MEAN, pb.MIN, pb.MAX, pb.COUNT, pb.STD)
    ex = {}

    def run(name, keys, vals, aggs, alg_bytes, n, filt=None, pred=None, reps=2):
        ng = [0]

        def f():
            r = ctx.groupby_agg(keys, vals, aggs, filter=filt, pred=pred)
            ng[0] = r.n_groups
            r.close()
        log(f"extra {name} ...")
        try:
            ms, kms = timed(f, reps=reps)
            log(f"extra {name}: {ms:.2f} ms")
            ex[name] = {"rows_per_s": n / (ms * 1e-3), "ms": ms, "kernel_ms": kms, "groups": ng[0], "alg_bytes_per_row": alg_bytes / n,
                        "roofline_frac": alg_bytes / (ms * 1e-3) / 1e9 / peak, "algo": ctx.stats()["groupby_algo_used"]}
        except Exception as e:  # noqa: BLE001
            ex[name] = {"error": str(e)[:200]}

    # ---- configs[3]
    n = int(500_000_000 * scale)
    log("generating configs[3] columns")
    k1 = zipf(n, 1000, torch.int32)
    k2 = zipf(n, 100_000, torch.int64)
    k3 = zipf(n, 10_000, torch.int32)           # dictionary ids over a 10,000-string pool (u32, same bits)
    v = uniform(n, 0.0, 1000.0)
    torch.cuda.current_stream(dev).synchronize()
    torch.cuda.empty_cache()                    # the generator's temporaries go back to the driver (the library checks what it can still allocate)
    aggs6 = [(0, op) for op in ALL6]
    run("c4_dict_key_zipf", [col(pb.DICT_U32, k3)], [col(pb.F64, v)], aggs6, n * 12.0, n)
    run("c4_i32_i64_keys_zipf", [col(pb.I32, k1), col(pb.I64, k2)], [col(pb.F64, v)], aggs6, n * 20.0, n)
    run("c4_i32_i64_dict_keys_zipf", [col(pb.I32, k1), col(pb.I64, k2), col(pb.DICT_U32, k3)], [col(pb.F64, v)], aggs6, n * 24.0, n)
    del k1, k2, k3, v
    torch.cuda.empty_cache()

    # ---- configs[4], one GPU's shard: 6e9 / 8 rows (python bench.py --workload c5 is the full benchmark of it)
    n = int(750_000_000 * scale)
    log("generating configs[4] columns")
    keys, vals, aggs, fmask, pred, T = make_c5(ctx, pb, torch, n, 4242)
    run("c5_q1_shard_filter_groupby", keys, vals, aggs, n * (8 + 5 * 8 + 0.125), n, filt=fmask)
    run("c5_q1_shard_predicate_in_scan", keys, vals, aggs, n * (8 + 5 * 8 + 8.0), n, pred=pred)
    return ex


# ---------------------------------------------------------------- the CUDA arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch

    import pandrs_b200 as pb

    env = Env()
    env.rank = int(os.environ.get("RANK", "0"))
    env.world = int(os.environ.get("WORLD_SIZE", "1"))
    env.local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(env.local)
    env.dist = None
    if env.world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", env.local))
        env.dist = dist
    env.stream = torch.cuda.current_stream()
    env.ctx = pb.Context(device=env.local, stream=env.stream.cuda_stream)
    for kv in filter(None, os.environ.get("PDRS_OPTS", "").split(",")):      # e.g. PDRS_OPTS=timing=2,trace=1 (debugging)
        k, v = kv.split("=")
        env.ctx.set_option(k, int(v))
    env.comm = pb.Comm(env.ctx, env.rank, env.world, pb.torch_broadcast_id(env.dist, torch.device("cuda", env.local))) if env.world > 1 else None

    def barrier():
        if env.dist is not None:
            env.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if env.dist is None:
            return float(x)
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        env.dist.all_reduce(t, op=env.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if env.dist is None:
            return int(x)
        t = torch.tensor([int(x)], device="cuda", dtype=torch.int64)
        env.dist.all_reduce(t)
        return int(t.item())
    env.barrier, env.max_over_ranks, env.sum_over_ranks = barrier, max_over_ranks, sum_over_ranks
    env.peaks = {}
    try:
        env.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    env.peak = float(env.peaks.get("hbm_gbs", 6650.0))

    if args.metric == "join":
        out = bench_join(env, args, pb)
    elif args.workload == "c5":
        out = bench_c5(env, args, pb)
    else:
        out = bench_groupby(env, args, pb, None)

    # ---- the reference's CPU algorithm on this box's host cores, bounded sample
    if env.rank == 0 and not args.no_cpu:
        import oracle as orc
        orc.build()
        threads = os.cpu_count() or 1
        if args.metric == "join":
            out["cpu_baseline"] = cpu_baseline_join(orc, threads, args.cpu_rows or 4_000_000)
        else:
            out["cpu_baseline"] = cpu_baseline_groupby(orc, args, threads, args.cpu_rows or 4_000_000)
    w = getattr(env, "step_walls", None)
    if w:      # rank 0's host wall clock of the timed steps of the headline: the mean (= ms_per_step) is sensitive to a few slow steps
        out["step_ms_host"] = {"min": w[0], "median": w[len(w) // 2], "max": w[-1]}
    if env.rank == 0:
        print(json.dumps(out))
    if env.comm is not None:
        env.comm.close()
    env.ctx.close()
    if env.dist is not None:
        env.dist.destroy_process_group()


if __name__ == "__main__":
    main()
