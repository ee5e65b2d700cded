"""Shared helpers for the parity tests: build the same column for the CUDA path (pandrs_b200.Column)
and for the CPU oracle (oracle.Col), run both, and compare after the canonical sort of SURVEY.md §9.1.4."""
from __future__ import annotations

import numpy as np

EMPTY_ID = 0xFFFFFFFF   # dictionary id the CUDA path reports for the empty string "" that a compat filter turns NULL keys into
RTOL = 1e-12   # north star: f64 sum/mean/std within 1e-12 relative; everything else bit-exact


class Spec:
    """dtype + numpy payload + optional bool null flags (+ string pool for dictionary columns)."""

    def __init__(self, dtype, values, nulls=None, pool=None, null_alias=-1):
        self.dtype, self.values, self.pool, self.null_alias = dtype, np.asarray(values), pool, null_alias
        self.nulls = None if nulls is None else np.asarray(nulls, dtype=bool)

    def __len__(self):
        return len(self.values)

    def gpu(self, pb):
        nb = None if self.nulls is None else pb.pack_bits(self.nulls)
        if self.dtype == pb.BOOL_BITS:
            return pb.Column(pb.BOOL_BITS, pb.pack_bits(self.values.astype(bool)), nb, length=len(self.values))
        return pb.Column(self.dtype, self.values, nb, null_alias=self.null_alias)

    def cpu(self, o):
        nb = None if self.nulls is None else o.pack_bits(self.nulls)
        if self.dtype == o.BOOL_BITS:
            return o.Col(o.BOOL_BITS, o.pack_bits(self.values.astype(bool)), nb, length=len(self.values))
        return o.Col(self.dtype, self.values, nb, pool=self.pool)


def f64_display(v: float) -> str:
    if np.isnan(v):
        return "NaN"
    if np.isinf(v):
        return "-inf" if v < 0 else "inf"
    return np.format_float_positional(v, trim="-")


def key_string(dtype, value, isnull, pool, pb) -> str:
    if isnull:
        return "NULL"
    if dtype in (pb.I64, pb.I32):
        return str(int(value))
    if dtype == pb.F64:
        return f64_display(float(value))
    if dtype == pb.DICT_U32:
        if int(value) == EMPTY_ID:
            return ""
        return pool[int(value)] if pool is not None else "#" + str(int(value))
    return "true" if value else "false"


def gpu_groupby_dict(pb, res, key_specs, naggs):
    """{key tuple of strings: (group_rows, [agg values])} from a GroupByResult."""
    cols = [res.key(k) for k in range(len(key_specs))]
    rows = res.group_rows()
    aggs = [res.agg(a) for a in range(naggs)]
    out = {}
    for g in range(res.n_groups):
        kt = tuple(key_string(s.dtype, cols[k][0][g], cols[k][1][g], s.pool, pb) for k, s in enumerate(key_specs))
        assert kt not in out, f"duplicate group {kt}"
        out[kt] = (int(rows[g]), [a[g] for a in aggs])
    return out


MEASURED = {}   # op -> [max error vs the reference (in units of the tolerance scale), max relative error vs the exact value]


def note_error(op, vs_ref, vs_exact):
    m = MEASURED.setdefault(int(op), [0.0, 0.0])
    m[0] = max(m[0], float(vs_ref))
    m[1] = max(m[1], float(vs_exact))


def oracle_groupby_dict(o, key_specs, val_specs, aggs, filter_spec=None, compat_nulls=False, mode=None):
    """Oracle groupby; a row filter is applied the way the reference does it: filter() first
    (data_ops.rs:37-121), then group_by on the filtered frame."""
    keys = [s for s in key_specs]
    vals = [s for s in val_specs]
    if filter_spec is not None:
        keep = filter_spec.values.astype(bool)
        if filter_spec.nulls is not None:
            keep &= ~filter_spec.nulls
        idx = np.nonzero(keep)[0]

        def take(s):
            nulls = None if s.nulls is None else s.nulls[idx]
            v = s.values[idx]
            if compat_nulls and nulls is not None:      # NULLs become defaults and the mask is dropped
                v = v.copy()
                v[nulls] = 0
                nulls = None
            return Spec(s.dtype, v, nulls, s.pool, s.null_alias)

        def take_key(s):
            nulls = None if s.nulls is None else s.nulls[idx]
            v = s.values[idx]
            pool = s.pool
            if compat_nulls and nulls is not None:
                v = v.copy()
                if s.dtype == o.DICT_U32:          # "" is interned on demand, like StringColumn::new does for the filtered values
                    pool = list(pool) if pool is not None else None
                    if pool is not None and "" not in pool:
                        pool.append("")
                    v[nulls] = pool.index("") if pool is not None else EMPTY_ID
                else:
                    v[nulls] = 0
                nulls = None
            return Spec(s.dtype, v, nulls, pool, s.null_alias)
        vals = [take(s) for s in vals]
        # the reference's filter() defaults the NULLs of EVERY column of the kept rows, key columns included
        # (data_ops.rs:64-108): a NULL Int64 key joins group "0", a NULL string key the group of the empty string
        keys = [take_key(s) for s in keys]
    r = o.groupby([s.cpu(o) for s in keys], [s.cpu(o) for s in vals], aggs, mode=o.MODE_AGGREGATE if mode is None else mode)
    assert r["error"] == 0
    out = {}
    for g, kt in enumerate(r["key_strings"]):
        out[kt] = (int(r["group_rows"][g]), [a[g] for a in r["aggs"]])
    return out


_CMP = {0: np.less, 1: np.less_equal, 2: np.greater, 3: np.greater_equal, 4: np.equal, 5: np.not_equal}


def compare_groupby(pb, o, ctx, key_specs, val_specs, aggs, filter_spec=None, device=False, compat_nulls=False, rtol=RTOL, pred=None):
    """Runs both paths and asserts parity.  Returns the GPU dict.  pred = (Spec, cmp op, constant): a typed predicate the CUDA
    path evaluates in its scan; the oracle gets the Boolean column a pandrs caller would have built (NULL -> not kept)."""
    kc = [s.gpu(pb) for s in key_specs]
    vc = [s.gpu(pb) for s in val_specs]
    fc = None if filter_spec is None else filter_spec.gpu(pb)
    pc = None if pred is None else pred[0].gpu(pb)
    ups = []
    if device:
        kc = [ctx.upload(c) for c in kc]
        vc = [ctx.upload(c) for c in vc]
        ups = kc + vc
        if fc is not None:
            fc = ctx.upload(fc)
            ups.append(fc)
        if pc is not None:
            pc = ctx.upload(pc)
            ups.append(pc)
    if pred is not None:
        with np.errstate(invalid="ignore"):
            keep = _CMP[pred[1]](pred[0].values, pred[2])
        if pred[0].nulls is not None:
            keep = keep & ~pred[0].nulls
        if filter_spec is not None:
            keep = keep & filter_spec.values.astype(bool)
            if filter_spec.nulls is not None:
                keep = keep & ~filter_spec.nulls
        filter_spec = Spec(pb.BOOL_BITS, keep)
    res = ctx.groupby_agg(kc, vc, aggs, filter=fc, pred=None if pred is None else (pc, pred[1], pred[2]))
    try:
        got = gpu_groupby_dict(pb, res, key_specs, len(aggs))
    finally:
        res.close()
        for c in ups:
            ctx.free(c)
    # Floating-point tolerance (north star: 1e-12 RELATIVE).  SUM: relative to sum|x| of the group (a sum of mixed signs has
    # no better conditioning than that); MEAN = SUM / n: relative to mean|x|; STD / VAR: relative to the value itself.  The
    # reference's own arithmetic (sequential f64 sums, mean rounded before the second pass) is not exact either: the oracle
    # is run a second time in ORC_MODE_EXACT (80-bit accumulation) and the reference's measured error |w - exact| is added
    # to the allowance - so the assertion reads "the CUDA result is within 1e-12 relative of the reference, or as close to it
    # as the reference is to the truth", never "within 1e-12 of max|x|".
    fp = [a for a, (v, op) in enumerate(aggs) if op in (pb.SUM, pb.MEAN, pb.STD, pb.VAR) and val_specs[v].dtype == pb.F64]
    fp += [a for a, (v, op) in enumerate(aggs) if op in (pb.STD, pb.VAR) and val_specs[v].dtype == pb.I64]
    scale_vals = list(val_specs)
    scale_aggs = []
    scale_of = {}
    for a, (v, op) in enumerate(aggs):
        if op in (pb.SUM, pb.MEAN) and val_specs[v].dtype == pb.F64:
            x = np.where(np.isfinite(val_specs[v].values), np.abs(val_specs[v].values), 0.0)
            scale_vals.append(Spec(pb.F64, x, val_specs[v].nulls))
            scale_of[a] = len(scale_aggs)
            scale_aggs.append((len(scale_vals) - 1, o.SUM if op == pb.SUM else o.MEAN))
    want = oracle_groupby_dict(o, key_specs, scale_vals, list(aggs) + scale_aggs, filter_spec, compat_nulls)
    exact = oracle_groupby_dict(o, key_specs, val_specs, [aggs[a] for a in fp], filter_spec, compat_nulls, mode=o.MODE_EXACT) if fp else {}
    assert set(got) == set(want), f"group keys differ: only gpu {sorted(set(got) - set(want))[:5]}, only oracle {sorted(set(want) - set(got))[:5]}"
    for kt, (rows, vals) in want.items():
        grows, gvals = got[kt]
        assert grows == rows, (kt, grows, rows)
        for a, (v, op) in enumerate(aggs):
            w, g = vals[a], gvals[a]
            is_exact = op in (pb.COUNT, pb.MIN, pb.MAX) or val_specs[v].dtype != pb.F64 and op in (pb.SUM, pb.MEAN)
            if is_exact:
                assert g == w or (np.isnan(g) and np.isnan(w)), (kt, a, op, g, w)
            elif np.isnan(w) or np.isinf(w):
                assert (np.isnan(g) and np.isnan(w)) or g == w, (kt, a, op, g, w)
            else:
                x = exact[kt][1][fp.index(a)]
                ref_err = abs(w - x) if np.isfinite(x) else 0.0
                scale = abs(w)
                if a in scale_of:
                    scale = max(scale, vals[len(aggs) + scale_of[a]])
                err = abs(g - w)
                note_error(op, err / max(scale, 1e-300), abs(g - x) / max(abs(x), 1e-300) if np.isfinite(x) else 0.0)
                assert err <= rtol * scale + 2.0 * ref_err + 1e-300, (kt, a, op, g, w, x, err / max(scale, 1e-300))
    return got


def canon_pairs(li, ri):
    order = np.lexsort((ri, li))
    return np.asarray(li)[order], np.asarray(ri)[order]


def compare_join(pb, o, ctx, left: Spec, right: Spec, how, device=False, check_order=True):
    lc, rc = left.gpu(pb), right.gpu(pb)
    ups = []
    if device:
        lc, rc = ctx.upload(lc), ctx.upload(rc)
        ups = [lc, rc]
    res = ctx.join_pairs(lc, rc, how)
    try:
        gl, gr = res.indices()
    finally:
        res.close()
        for c in ups:
            ctx.free(c)
    wl, wr = o.join(left.cpu(o), right.cpu(o), how)
    assert len(gl) == len(wl), (len(gl), len(wl))
    if check_order:   # the CUDA path keeps the reference's order: left-row-major, ascending right row
        assert np.array_equal(gl, wl) and np.array_equal(gr, wr)
    else:
        a, b = canon_pairs(gl, gr), canon_pairs(wl, wr)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    return gl, gr


# ---------------------------------------------------------------- large inputs: the typed-key oracle (oracle/typed_oracle.cpp)
def _sorted_typed(tg):
    cols = []
    for v, isn in reversed(tg["keys"]):
        cols += [v, isn.astype(np.uint8)]
    return np.lexsort(cols)


def compare_groupby_typed(pb, res, tg, key_dtypes, ops, rtol=RTOL):
    """GroupByResult `res` (aggregates `ops` of value column 0, in that order) against a typed-oracle result `tg`
    (oracle.typed_groupby / typed_groupby_synth), vectorised: group keys / rows / valid counts / min / max bit-exact,
    f64 sum / mean / std / var within rtol RELATIVE plus twice the reference's own measured rounding error (tg["exact"])."""
    G = res.n_groups
    assert G == tg["n_groups"], (G, tg["n_groups"])
    gk = []
    for k, dt in enumerate(key_dtypes):
        v, isn = res.key(k)
        if dt in (pb.I64, pb.I32):
            u = v.astype(np.int64).view(np.uint64)          # sign-extended, like the oracle's key words
        elif dt == pb.F64:
            u = np.where(np.isnan(v), np.float64("nan"), v).view(np.uint64)
        else:
            u = v.astype(np.uint64)
        gk.append((np.where(isn, np.uint64(0), u), isn))
    gcols = []
    for v, isn in reversed(gk):
        gcols += [v, isn.astype(np.uint8)]
    go, to = np.lexsort(gcols), _sorted_typed(tg)
    for (gv, gn), (tv, tn) in zip(gk, tg["keys"]):
        assert np.array_equal(gv[go], tv[to]) and np.array_equal(gn[go], tn[to]), "group keys differ"
    assert np.array_equal(res.group_rows()[go], tg["group_rows"][to]), "group sizes differ"
    assert np.array_equal(res.valid_n(0)[go], tg["valid_n"][to]), "valid counts differ"
    names = {pb.SUM: "sum", pb.MEAN: "mean", pb.MIN: "min", pb.MAX: "max", pb.STD: "std", pb.VAR: "var"}
    worst = {}
    for a, op in enumerate(ops):
        g = res.agg(a)[go]
        if op == pb.COUNT:
            assert np.array_equal(g, tg["group_rows"][to].astype(np.float64))
            continue
        w = tg[names[op]][to]
        if op in (pb.MIN, pb.MAX):
            assert np.array_equal(g, w), names[op]
            continue
        x = tg["exact"][names[op]][to]
        allow = rtol * np.abs(w) + 2.0 * np.abs(w - x) + 1e-300
        err = np.abs(g - w)
        rel = float((err / np.maximum(np.abs(w), 1e-300)).max()) if G else 0.0
        relx = float((np.abs(g - x) / np.maximum(np.abs(x), 1e-300)).max()) if G else 0.0
        note_error(op, rel, relx)
        worst[names[op]] = (rel, relx)
        bad = np.nonzero(err > allow)[0]
        assert len(bad) == 0, (names[op], len(bad), g[bad[:3]], w[bad[:3]], x[bad[:3]])
    return worst


def device_pair_stats(ctx, j, chunk=1 << 27):
    """(count, checksum, sum_left, sum_right over r >= 0, number of r == -1) of a JoinResult, computed on the device with
    torch integer ops (wrapping 64-bit arithmetic) - the same quantities oracle.typed_join reports."""
    import torch
    A = 0x9E3779B97F4A7C15 - (1 << 64)
    B = 0xC2B2AE3D27D4EB4F - (1 << 64)
    dev = torch.device("cuda", ctx.device)
    cs = sl = sr = un = 0
    for off in range(0, j.n, chunk):
        m = min(chunk, j.n - off)
        l = torch.empty(m, dtype=torch.int64, device=dev)
        r = torch.empty(m, dtype=torch.int64, device=dev)
        torch.cuda.synchronize(dev)
        ctx.memcpy(l.data_ptr(), j.left_dev() + 8 * off, 8 * m, 2)
        ctx.memcpy(r.data_ptr(), j.right_dev() + 8 * off, 8 * m, 2)
        cs += int(((l * A) ^ (r * B)).sum().item())
        sl += int(l.sum().item())
        sr += int(torch.where(r >= 0, r, torch.zeros_like(r)).sum().item())
        un += int((r < 0).sum().item())
        del l, r
    return j.n, cs & ((1 << 64) - 1), sl, sr, un


def mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0
