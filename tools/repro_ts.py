"""Tile-sort kernel repro: python tools/repro_ts.py VARIANT N"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import pandrs_b200 as pb
from _util import Spec
variant, n = sys.argv[1], int(sys.argv[2])
rng = np.random.default_rng(10)
knull = rng.random(n) < 0.01 if "k" in variant else None
k = Spec(pb.I64, rng.integers(0, 700, n), nulls=knull)
v = Spec(pb.F64, 1e6 + rng.normal(0, 3.0, n), nulls=rng.random(n) < 0.05)
vi = Spec(pb.I64, rng.integers(-10**9, 10**9, n))
vals = []
if "f" in variant: vals.append(v.gpu(pb))
if "i" in variant: vals.append(vi.gpu(pb))
ctx = pb.Context(0)
if "p" in variant:
    r = ctx.groupby_partial([k.gpu(pb)], vals, all_stats=True)
else:
    r = ctx.groupby_agg([k.gpu(pb)], vals, [(0, pb.SUM), (0, pb.STD)])
print(variant, n, "groups", r.n_groups, ctx.stats())
