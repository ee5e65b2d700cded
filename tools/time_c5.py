"""configs[4] on one GPU: python tools/time_c5.py [rows] - times the few-groups kernel with the precomputed mask and with the
date predicate evaluated in the scan, and the round-1 path (one tile-sort pass per value column) for comparison."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import pandrs_b200 as pb

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 750_000_000
torch.cuda.set_device(0)
ctx = pb.Context(0, stream=torch.cuda.current_stream().cuda_stream)
keys, vals, aggs, fmask, pred, T = bench.make_c5(ctx, pb, torch, n, 4242)
peak = 6504.1


def run(name, bpr, **kw):
    best, kms, algo = 1e9, 0, 0
    for _ in range(4):
        ctx.timer_begin()
        r = ctx.groupby_agg(keys, vals, aggs, **kw)
        r.close()
        ms = ctx.timer_end()
        if ms < best:
            best, kms, algo = ms, ctx.stats()["main_kernel_ms"], ctx.stats()["groupby_algo_used"]
    print(f"{name}: {best:.3f} ms (kernels {kms:.3f} ms, algo {algo}) = {bpr * n / (kms * 1e-3) / 1e9:.0f} GB/s = {bpr * n / (kms * 1e-3) / 1e9 / peak * 100:.1f}% of {peak} GB/s", flush=True)


run("mask filter, few-groups kernel (48.125 B/row)", 48.125, filter=fmask)
run("date predicate in the scan (56 B/row)", 56.0, pred=pred)
run("no filter (48 B/row)", 48.0)
ctx.set_option("few", 0)
run("mask filter, round-1 path: one pass per value column", 48.125, filter=fmask)
