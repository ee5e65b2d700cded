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
    ap.add_argument("--extras-scale", type=float, default=1.0, help="scales the row counts of the configs[3..4] extras (debugging)")
    ap.add_argument("--dist-join", action="store_true", help="also time the sharded inner join (fused partition + NVLink shuffle), rows/GPU = --rows x --rows/10")
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


def cpu_idealised(orc, n, groups, threads, seed=42):
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
        "cpu_baseline": {"value": val, "unit": "rows/s", "cores": threads, "kind": "port", "sample": sample,
                         "idealised_typed_key_rows_per_s": cpu_idealised(orc, 8 * n, args.groups, threads)},
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

    log(f"headline: {ms_per_step:.3f} ms/step")
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

    log("e2e done")
    # ---- the other configs, a few steps each (explanatory; not the headline)
    if not args.no_extras and world == 1:
        out["extras"] = extras(ctx, pb, args, n, keys, vals, peak)

    if args.dist_join:
        try:
            out["dist_join"] = dist_join(ctx, pb, dist, rank, world, n, peaks)
        except Exception as e:  # noqa: BLE001
            out["dist_join"] = {"error": str(e)[:300]}

    # ---- the reference's CPU algorithm on this box's host cores, bounded sample
    if rank == 0 and not args.no_cpu:
        import oracle as orc
        orc.build()
        threads = os.cpu_count() or 1
        cpu_reference_step(orc, 200_000, args.groups, threads)
        dt = cpu_reference_step(orc, args.cpu_rows, args.groups, threads)
        out["cpu_baseline"] = {"value": args.cpu_rows / dt, "unit": "rows/s", "cores": threads, "kind": "port",
                               "sample": f"{args.cpu_rows} rows, same generator/cardinality/NULLs; oracle restatement of grouping.rs + aggregation.rs (string keys, serial grouping, {threads}-thread aggregation)",
                               # not the reference's algorithm: typed keys, per-thread flat tables (SURVEY.md §8d "idealised CPU"), 8x the sample
                               "idealised_typed_key_rows_per_s": cpu_idealised(orc, 8 * args.cpu_rows, args.groups, threads)}
    if rank == 0:
        print(json.dumps(out))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


class _OneRank:
    """torch.distributed stand-in for --dist-join on one GPU (every collective is a copy)."""
    class ReduceOp:
        MIN = MAX = None
    @staticmethod
    def get_world_size(): return 1
    @staticmethod
    def get_rank(): return 0
    @staticmethod
    def all_gather(outs, t): outs[0].copy_(t)
    @staticmethod
    def all_reduce(t, op=None): return None
    @staticmethod
    def barrier(): return None


def dist_join(ctx, pb, dist, rank, world, n_probe, peaks):
    """BASELINE.json configs[2] sharded over the ranks (weak scaling: n_probe probe rows and n_probe / 10 unique build
    rows per GPU, keys drawn over the GLOBAL domain so ~(world-1)/world of all rows change GPU): the one-pass radix
    partition stores its bucket runs straight into the destination GPU's receive area through NVLink
    (pdrs_xjoin_shuffle), one barrier, local build / probe (pdrs_xjoin_local).  Times are CUDA events of the two
    phases, max over ranks; the NVLink figure counts the 12-byte (key, row) records that leave each GPU."""
    import torch
    from pandrs_b200.dist import DistJoin
    d = dist if dist is not None else _OneRank
    if world > 2:      # staged exchange: left rows travel as (source rank << (32 - log2 world) | local row)
        n_probe = min(n_probe, (1 << (32 - (world - 1).bit_length())) // 100_000_000 * 100_000_000 or n_probe)
    nb_, np_ = n_probe // 10, n_probe
    build = ctx.synth_join_keys(nb_, unique=True, row0=rank * nb_)
    probe = ctx.synth_join_keys(np_, domain=2 * nb_ * world, row0=rank * np_)
    if os.environ.get("PDRS_XJOIN_MODE"):
        ctx.set_option("xjoin_mode", int(os.environ["PDRS_XJOIN_MODE"]))      # 1 = fused, 2 = staged (default: auto)
    dj = DistJoin(ctx, d)
    if not dj.setup_fused(np_, nb_, nb_ * world):
        return {"error": "fused setup failed: " + getattr(dj, "fused_error", "?")[:200]}
    best = None
    pairs = 0
    for rep in range(3):
        tm = {}
        t0 = time.perf_counter()
        j = dj.join_pairs_fused(probe, build, pb.INNER, rank * np_, rank * nb_, timings=tm)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        if j is None:
            return {"error": "sub-bucket overflow: " + getattr(dj, "fused_error", "?")[:200]}
        pairs = j.n
        j.close()
        t = torch.tensor([tm["shuffle_ms"], tm["local_ms"], wall], device="cuda", dtype=torch.float64)
        tot = torch.tensor([float(pairs)], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        sh, lo, wl = (float(v) for v in t.tolist())
        if rep and (best is None or sh + lo < best[0] + best[1]):
            best = (sh, lo, wl, float(tot.item()))
    dj.x.close()
    # the same join through the plain path: pdrs_hash_partition + gathers + NCCL all_to_all + pdrs_join_pairs (wall clock, max over ranks)
    base_ms = None
    try:
        if dist is not None:
            for rep in range(2):
                dist.barrier(); torch.cuda.synchronize()
                t0 = time.perf_counter()
                gl, gr = dj.join_pairs(probe, build, pb.INNER, rank * np_, rank * nb_)
                torch.cuda.synchronize()
                tb = torch.tensor([(time.perf_counter() - t0) * 1e3], device="cuda", dtype=torch.float64)
                del gl, gr
                dist.all_reduce(tb, op=dist.ReduceOp.MAX)
                base_ms = float(tb.item())
    except Exception as e:  # noqa: BLE001
        base_ms = "error: " + str(e)[:120]
    sh, lo, wl, total_pairs = best
    sent = (np_ + nb_) * 12.0 * (world - 1) / world          # bytes leaving each GPU
    nvl_peak = float(peaks.get("nvlink_gbs", 770.0))
    return {"rows_per_s": (np_ + nb_) * world / ((sh + lo) * 1e-3), "shuffle_ms": sh, "local_ms": lo, "wall_ms": wl, "pairs": total_pairs,
            "rows_per_gpu": np_ + nb_, "nvlink_bytes_per_gpu": sent, "nvlink_gbs_achieved": sent / (sh * 1e-3) / 1e9 if world > 1 else None,
            "nvlink_peak_gbs": nvl_peak, "nvlink_frac": sent / (sh * 1e-3) / 1e9 / nvl_peak if world > 1 else None,
            "nccl_all_to_all_path_wall_ms": base_ms,
            "xjoin_mode": os.environ.get("PDRS_XJOIN_MODE", "auto"),
            "note": "weak scaling; shuffle_ms = partition kernel storing into the peers' receive areas, both sides; local_ms = (staged mode: local radix partition +) table memset + build + probe/emit"}


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
        # configs[2] with its 2 build-side payload columns (i64, f64) gathered into the result (join.rs:290-552)
        p1, p2 = ctx.synth_keys(nb_, card=1 << 40, seed=9), ctx.synth_vals(nb_, seed=9)
        o1, o2 = ctx.dev_alloc(np_ * 8), ctx.dev_alloc(np_ * 8)
        m = [0]

        def jn_payload():
            j = ctx.join_pairs(probe, build, pb.INNER)
            m[0] = j.n
            ctx.gather(p1, j.right_dev(), n=j.n, idx_dev=True, out_dev=o1)
            ctx.gather(p2, j.right_dev(), n=j.n, idx_dev=True, out_dev=o2)
            j.close()
        log("extra join_inner_2payload ...")
        ms, _ = timed(jn_payload, reps=2)
        alg = 8 * np_ + nb_ * (8 + 16) + m[0] * (16 + 16)      # SURVEY.md §8(d), P = 2 payload columns
        ex["join_inner_2payload"] = {"rows_per_s": (np_ + nb_) / (ms * 1e-3), "ms": ms, "pairs": m[0], "alg_bytes": alg, "roofline_frac": alg / (ms * 1e-3) / 1e9 / peak}
        ctx.dev_free(o1); ctx.dev_free(o2)
        del p1, p2, build, probe
    except Exception as e:  # noqa: BLE001
        ex["join"] = {"error": str(e)[:200]}
    try:
        ex.update(extras_c4_c5(ctx, pb, peak, timed, args.extras_scale))
    except Exception as e:  # noqa: BLE001
        ex["c4_c5"] = {"error": str(e)[:200]}
    return ex


def log(msg):
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


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

    def bits(n, p_true):
        nbytes = (n + 63) // 64 * 8
        out = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        wts = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.int32, device=dev)
        for a in range(0, n, CH):
            b = min(n, a + CH)
            f = torch.zeros((b - a + 7) // 8 * 8, dtype=torch.int32, device=dev)
            f[: b - a] = (torch.rand(b - a, device=dev, generator=g) < p_true).to(torch.int32)
            out[a // 8: a // 8 + f.numel() // 8] = (f.view(-1, 8) * wts).sum(1).to(torch.uint8)
        return out

    def col(dtype, t, n=None):
        return pb.Column(dtype, device_ptr=t.data_ptr(), length=int(n if n is not None else t.numel()), owner=t)

    ALL6 = (pb.SUM, pb.MEAN, pb.MIN, pb.MAX, pb.COUNT, pb.STD)
    ex = {}

    def run(name, keys, vals, aggs, alg_bytes, n, filt=None, reps=2):
        ng = [0]

        def f():
            r = ctx.groupby_agg(keys, vals, aggs, filter=filt)
            ng[0] = r.n_groups
            r.close()
        log(f"extra {name} ...")
        try:
            ms, _ = timed(f, reps=reps)
            log(f"extra {name}: {ms:.2f} ms")
            ex[name] = {"rows_per_s": n / (ms * 1e-3), "ms": ms, "groups": ng[0], "alg_bytes_per_row": alg_bytes / n,
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
    aggs6 = [(0, op) for op in ALL6]
    run("c4_dict_key_zipf", [col(pb.DICT_U32, k3)], [col(pb.F64, v)], aggs6, n * 12.0, n)
    run("c4_i32_i64_keys_zipf", [col(pb.I32, k1), col(pb.I64, k2)], [col(pb.F64, v)], aggs6, n * 20.0, n)
    run("c4_i32_i64_dict_keys_zipf", [col(pb.I32, k1), col(pb.I64, k2), col(pb.DICT_U32, k3)], [col(pb.F64, v)], aggs6, n * 24.0, n)
    del k1, k2, k3, v
    torch.cuda.empty_cache()

    # ---- configs[4], one GPU's shard: 6e9 / 8 rows
    n = int(750_000_000 * scale)
    log("generating configs[4] columns")
    rf = torch.empty(n, dtype=torch.int32, device=dev)
    ls = torch.empty(n, dtype=torch.int32, device=dev)
    for a in range(0, n, CH):
        b = min(n, a + CH)
        rf[a:b] = torch.randint(0, 3, (b - a,), device=dev, generator=g, dtype=torch.int32)
        ls[a:b] = torch.randint(0, 2, (b - a,), device=dev, generator=g, dtype=torch.int32)
    qty, price, disc, tax = uniform(n, 1, 50), uniform(n, 900, 105000), uniform(n, 0, 0.1), uniform(n, 0, 0.08)
    disc_price = torch.empty_like(price)
    charge = torch.empty_like(price)
    for a in range(0, n, CH):
        b = min(n, a + CH)
        disc_price[a:b] = price[a:b] * (1 - disc[a:b])
        charge[a:b] = disc_price[a:b] * (1 + tax[a:b])
    del tax
    mask = bits(n, 0.98)                         # shipdate <= cutoff as a precomputed Boolean column (data_ops.rs:37-62)
    torch.cuda.current_stream(dev).synchronize()
    vals = [col(pb.F64, t) for t in (qty, price, disc_price, charge, disc)]
    aggs = [(0, pb.SUM), (1, pb.SUM), (2, pb.SUM), (3, pb.SUM), (0, pb.MEAN), (1, pb.MEAN), (4, pb.MEAN), (0, pb.COUNT)]
    run("c5_q1_shard_filter_groupby", [col(pb.DICT_U32, rf), col(pb.DICT_U32, ls)], vals, aggs, n * (8 + 5 * 8 + 0.125), n,
        filt=col(pb.BOOL_BITS, mask, n))
    return ex


if __name__ == "__main__":
    main()
