"""One device-resident groupby (or join) call for ncu: python tools/prof_groupby.py --rows N --groups G --aggs all|sum [--join]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pandrs_b200 as pb

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=200_000_000)
ap.add_argument("--groups", type=int, default=1000)
ap.add_argument("--aggs", default="all")
ap.add_argument("--join", action="store_true")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--opt", action="append", default=[], help="name=value passed to pdrs_set_option")
a = ap.parse_args()
ctx = pb.Context(0)
for o in a.opt:
    k, v = o.split("=")
    ctx.set_option(k, int(v))
if a.join:
    build = ctx.synth_join_keys(a.rows // 10, unique=True)
    probe = ctx.synth_join_keys(a.rows, domain=2 * (a.rows // 10))
    for _ in range(a.reps):
        j = ctx.join_pairs(probe, build, pb.INNER)
        print("pairs", j.n, ctx.stats())
        j.close()
else:
    keys = ctx.synth_keys(a.rows, card=a.groups)
    vals = ctx.synth_vals(a.rows, null_per_million=50_000)
    ops = [pb.SUM, pb.MEAN, pb.MIN, pb.MAX, pb.COUNT, pb.STD] if a.aggs == "all" else [pb.SUM]
    for _ in range(a.reps):
        r = ctx.groupby_agg([keys], [vals], [(0, op) for op in ops])
        print("groups", r.n_groups, ctx.stats())
        r.close()
ctx.close()
