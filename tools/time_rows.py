"""Timings of the rows next to the path on one B200: row lists per group (par_groupby), Median / First / Last,
Arrow validity -> null mask, dictionary encoding.  python tools/time_rows.py [rows]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import pandrs_b200 as pb

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
ctx = pb.Context(0)


def best(fn, reps=3):
    fn()
    b = 1e30
    for _ in range(reps):
        ctx.sync()
        t0 = time.perf_counter()
        fn()
        ctx.sync()
        b = min(b, (time.perf_counter() - t0) * 1e3)
    return b


vals = ctx.synth_vals(n, null_per_million=50_000)
for card in (1000, 10_000_000):
    keys = ctx.synth_keys(n, card=card)
    holder = []

    def rows():
        r = ctx.groupby_rows([keys])
        holder.append(r.n_groups)
        r.close()
    ms = best(rows)
    print(f"pdrs_groupby_rows  {n:.0e} rows, {holder[-1]} groups: {ms:9.2f} ms  = {n / ms / 1e6:7.2f} G rows/s  ({16.0 * n / ms / 1e6:7.1f} GB/s of keys in + row ids out)", flush=True)
    r = ctx.groupby_rows([keys])
    for op, name in ((pb.FIRST, "first"), (pb.LAST, "last")):
        ms = best(lambda: r.agg(vals, op))
        print(f"  {name:6s} over the row lists: {ms:9.2f} ms", flush=True)
    r.close()
    ctx.free(keys)
m = min(n, 200_000_000)
keys = ctx.synth_keys(m, card=1000)
v2 = ctx.synth_vals(m, null_per_million=50_000)
r = ctx.groupby_rows([keys])
ms = best(lambda: r.agg(v2, pb.MEDIAN), reps=2)
print(f"  median over the row lists, {m:.0e} rows, 1000 groups: {ms:9.2f} ms = {m / ms / 1e6:6.2f} G rows/s", flush=True)
r.close()
ctx.free(keys); ctx.free(v2); ctx.free(vals)

# Arrow validity -> pandrs null mask, device to device
nb = (n + 7) // 8
src, dst = ctx.dev_alloc(nb + 64), ctx.dev_alloc(nb + 64)
import ctypes as C
cnt = C.c_int64()
def conv():
    ctx._chk(ctx.L.pdrs_arrow_validity_to_nulls(ctx._h, src, pb.MEM_DEVICE, 3, n - 8, dst, pb.MEM_DEVICE, C.byref(cnt)))
ms = best(conv)
print(f"pdrs_arrow_validity_to_nulls  {n:.0e} rows (bit offset 3): {ms:7.3f} ms = {2 * nb / ms / 1e6:7.1f} GB/s", flush=True)
ctx.dev_free(src); ctx.dev_free(dst)

# dictionary encoding: m strings of 4 - 19 bytes over `card` distinct values, buffers resident on the device
for m, card in ((100_000_000, 10_000), (100_000_000, 10_000_000)):
    rng = np.random.default_rng(1)
    ids = rng.integers(0, card, m)
    lens = (4 + ids % 16).astype(np.int64)         # 4 id bytes + 0 - 15 padding bytes: equal ids <=> equal strings
    off = np.zeros(m + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    total = int(off[-1])
    # bytes: the id in base 256 repeated - equal ids give equal strings, distinct ids distinct strings of the right length
    data = np.zeros(total, np.uint8)
    for b in range(4):
        pos = off[:-1] + b
        ok = lens > b
        data[pos[ok]] = ((ids[ok] >> (8 * b)) & 255).astype(np.uint8)
    d_off, d_data = ctx.dev_alloc(8 * (m + 1)), ctx.dev_alloc(total + 64)
    ctx.memcpy(d_off, off.ctypes.data, 8 * (m + 1), 0)
    ctx.memcpy(d_data, data.ctypes.data, total, 0)
    h = C.c_void_p()
    uniq = [0]
    def enc():
        ctx._chk(ctx.L.pdrs_dict_encode(ctx._h, d_off, 1, d_data, total, None, 0, m, pb.MEM_DEVICE, C.byref(h)))
        uniq[0] = ctx.L.pdrs_dict_n_unique(h)
        ctx.L.pdrs_dict_free(h)
    ms = best(enc, reps=2)
    print(f"pdrs_dict_encode  {m:.0e} strings ({total / m:.1f} bytes each), {uniq[0]} distinct: {ms:8.2f} ms = {m / ms / 1e6:6.2f} G strings/s, {(total + 12 * m) / ms / 1e6:6.1f} GB/s (bytes + offsets in, ids out)", flush=True)
    ctx.dev_free(d_off); ctx.dev_free(d_data)
