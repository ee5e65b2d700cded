"""Phase timings of the single-GPU join (BASELINE.json configs[2]): python tools/time_join.py [rows_probe] [sweep]
Prints the [pdrs join] marks (timing = 2) of the bucket-at-a-time path, pairs only and with 1 / 2 payload columns, and of
the one-table path; `sweep` also varies the region size / slots per key."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pandrs_b200 as pb

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
sweep = len(sys.argv) > 2
ctx = pb.Context(0)
nb = n // 10
build = ctx.synth_join_keys(nb, unique=True)
probe = ctx.synth_join_keys(n, domain=2 * nb)
p1 = ctx.synth_keys(nb, card=1 << 40, seed=9)
p2 = ctx.synth_vals(nb, seed=9)


def run(name, fn, reps=3):
    fn()
    best = 1e9
    for r in range(reps):
        ctx.set_option("timing", 2 if r == reps - 1 else 1)
        ctx.timer_begin()
        fn()
        best = min(best, ctx.timer_end())
    ctx.set_option("timing", 1)
    print(f"== {name}: best {best:.3f} ms  ({(n + nb) / best / 1e6:.1f} G rows/s)", flush=True)


def pairs(how):
    def f():
        j = ctx.join_pairs(probe, build, how)
        j.close()
    return f


def gath(cols, how=pb.INNER):
    def f():
        j = ctx.join_gather(probe, build, how, cols)
        j.close()
    return f


run("inner pairs, bucket-at-a-time", pairs(pb.INNER))
run("left pairs, bucket-at-a-time", pairs(pb.LEFT))
run("inner + 1 payload column", gath([p2]))
run("inner + 2 payload columns", gath([p1, p2]))
run("left + 2 payload columns", gath([p1, p2], pb.LEFT))
ctx.set_option("join_bucketwise", 0)
run("inner pairs, one table (round-1 path)", pairs(pb.INNER), reps=2)
run("inner + 2 payload columns, one table + gathers (round-1 path)", gath([p1, p2]), reps=2)
ctx.set_option("join_bucketwise", 1)
if sweep:
    for mb in (12, 16, 24, 32, 48):
        for spk in (2, 3):
            ctx.set_option("join_region_mb", mb); ctx.set_option("join_slots_mult", spk)
            run(f"inner pairs region {mb} MB, {spk} slots/key", pairs(pb.INNER), reps=2)
            run(f"inner + 2 payload region {mb} MB, {spk} slots/key", gath([p1, p2]), reps=2)
