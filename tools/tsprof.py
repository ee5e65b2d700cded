"""Per-phase SM-cycle breakdown of the tile-sort kernel (library built with -DTS_PROFILE): python tools/tsprof.py [rows] [groups]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pandrs_b200 as pb
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000_000
groups = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ctx = pb.Context(0)
keys = ctx.synth_keys(rows, card=groups)
vals = ctx.synth_vals(rows, null_per_million=50_000)
ops = [pb.SUM, pb.MEAN, pb.MIN, pb.MAX, pb.COUNT, pb.STD]
L = ctx.L
buf = (C.c_ulonglong * 8)()
r = ctx.groupby_agg([keys], [vals], [(0, op) for op in ops]); r.close()
L.pdrs_debug_tsprof(None, 1)
r = ctx.groupby_agg([keys], [vals], [(0, op) for op in ops]); r.close()
print(ctx.stats())
L.pdrs_debug_tsprof(buf, 0)
names = ["phase1 (ids+hist)", "barrier 1", "phase2 (scan, 2 barriers)", "phase3 (scatter)", "barrier 3", "phase4 (reduce)"]
nwarps = 148 * 16
batches = rows / 32 / 148
tot = sum(buf[i] for i in range(6))
for i, nm in enumerate(names):
    print(f"{nm:28s} {buf[i] / nwarps / batches:7.2f} cycles per 32-row batch  ({100.0 * buf[i] / tot:4.1f}%)")
print(f"{'total':28s} {tot / nwarps / batches:7.2f}")
