"""Shared helpers for the parity tests: build the same column for the CUDA path (pandrs_b200.Column)
and for the CPU oracle (oracle.Col), run both, and compare after the canonical sort of SURVEY.md §9.1.4."""
from __future__ import annotations

import numpy as np

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


def oracle_groupby_dict(o, key_specs, val_specs, aggs, filter_spec=None, compat_nulls=False):
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
        vals = [take(s) for s in vals]
        keys = [Spec(s.dtype, s.values[idx], None if s.nulls is None else s.nulls[idx], s.pool, s.null_alias) for s in keys]
    r = o.groupby([s.cpu(o) for s in keys], [s.cpu(o) for s in vals], aggs)
    assert r["error"] == 0
    out = {}
    for g, kt in enumerate(r["key_strings"]):
        out[kt] = (int(r["group_rows"][g]), [a[g] for a in r["aggs"]])
    return out


def compare_groupby(pb, o, ctx, key_specs, val_specs, aggs, filter_spec=None, device=False, compat_nulls=False, rtol=RTOL):
    """Runs both paths and asserts parity.  Returns the GPU dict."""
    kc = [s.gpu(pb) for s in key_specs]
    vc = [s.gpu(pb) for s in val_specs]
    fc = None if filter_spec is None else filter_spec.gpu(pb)
    ups = []
    if device:
        kc = [ctx.upload(c) for c in kc]
        vc = [ctx.upload(c) for c in vc]
        ups = kc + vc
        if fc is not None:
            fc = ctx.upload(fc)
            ups.append(fc)
    res = ctx.groupby_agg(kc, vc, aggs, filter=fc)
    try:
        got = gpu_groupby_dict(pb, res, key_specs, len(aggs))
    finally:
        res.close()
        for c in ups:
            ctx.free(c)
    # scale for the floating-point tolerance: sum of |x| and rms per group, from the oracle on |x| and x^2
    scale_vals = list(val_specs)
    scale_aggs = []
    scale_of = {}
    for a, (v, op) in enumerate(aggs):
        if op in (pb.SUM, pb.MEAN, pb.STD, pb.VAR) and val_specs[v].dtype == pb.F64:
            x = np.where(np.isfinite(val_specs[v].values), np.abs(val_specs[v].values), 0.0)
            scale_vals.append(Spec(pb.F64, x, val_specs[v].nulls))
            scale_of[a] = len(scale_aggs)
            scale_aggs.append((len(scale_vals) - 1, o.MAX if op in (pb.STD, pb.VAR) else (o.SUM if op == pb.SUM else o.MAX)))
    want = oracle_groupby_dict(o, key_specs, scale_vals, list(aggs) + scale_aggs, filter_spec, compat_nulls)
    assert set(got) == set(want), f"group keys differ: only gpu {sorted(set(got) - set(want))[:5]}, only oracle {sorted(set(want) - set(got))[:5]}"
    for kt, (rows, vals) in want.items():
        grows, gvals = got[kt]
        assert grows == rows, (kt, grows, rows)
        for a, (v, op) in enumerate(aggs):
            w, g = vals[a], gvals[a]
            exact = op in (pb.COUNT, pb.MIN, pb.MAX) or val_specs[v].dtype != pb.F64 and op == pb.SUM
            if exact:
                assert g == w or (np.isnan(g) and np.isnan(w)), (kt, a, op, g, w)
            elif np.isnan(w) or np.isinf(w):
                assert (np.isnan(g) and np.isnan(w)) or g == w, (kt, a, op, g, w)
            else:
                scale = abs(w)
                if a in scale_of:
                    s = vals[len(aggs) + scale_of[a]]
                    scale = max(scale, s * s if op == pb.VAR else s)
                assert abs(g - w) <= rtol * scale + 1e-300, (kt, a, op, g, w, abs(g - w) / max(scale, 1e-300))
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
