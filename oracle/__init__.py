"""ctypes wrapper around the CPU oracle (oracle/pandrs_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  pandrs_b200/ never imports this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpandrs_oracle.so")

I64, F64, DICT_U32, BOOL_BITS, I32 = 0, 1, 2, 3, 4
SUM, MEAN, MIN, MAX, COUNT, STD, VAR, MEDIAN, FIRST, LAST = range(10)
MODE_AGGREGATE, MODE_PAR_AGGREGATE, MODE_LAZY, MODE_EXACT = 0, 1, 2, 3
INNER, LEFT, RIGHT, OUTER = 0, 1, 2, 3

_NP = {I64: np.int64, F64: np.float64, DICT_U32: np.uint32, BOOL_BITS: np.uint8, I32: np.int32}


class _Col(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("_pad", C.c_int32), ("data", C.c_void_p),
                ("null_bits", C.c_void_p), ("null_len", C.c_int64), ("len", C.c_int64),
                ("pool", C.POINTER(C.c_char_p)), ("pool_len", C.c_int64)]


class _JSynth(C.Structure):
    _fields_ = [("n", C.c_int64), ("seed", C.c_uint64), ("domain", C.c_uint64), ("unique", C.c_int)]


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("pandrs_oracle.cpp", "typed_oracle.cpp")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.run(["make", "-C", _HERE, "-B", "libpandrs_oracle.so"], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.orc_groupby.restype = C.c_void_p
        L.orc_groupby.argtypes = [C.POINTER(_Col), C.c_int, C.POINTER(_Col), C.POINTER(C.c_int32),
                                  C.POINTER(C.c_int32), C.c_int, C.c_int64, C.c_int, C.c_int]
        L.orc_gb_error.argtypes = [C.c_void_p]
        L.orc_gb_ngroups.restype = C.c_int64
        L.orc_gb_ngroups.argtypes = [C.c_void_p]
        for f in (L.orc_gb_first_rows, L.orc_gb_group_rows):
            f.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_gb_agg.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_gb_key.restype = C.c_char_p
        L.orc_gb_key.argtypes = [C.c_void_p, C.c_int64, C.c_int]
        L.orc_gb_free.argtypes = [C.c_void_p]
        L.orc_par_groupby.restype = C.c_void_p
        L.orc_par_groupby.argtypes = [C.POINTER(_Col), C.c_int, C.c_int64]
        L.orc_pg_ngroups.restype = C.c_int64
        L.orc_pg_ngroups.argtypes = [C.c_void_p]
        L.orc_pg_label.restype = C.c_char_p
        L.orc_pg_label.argtypes = [C.c_void_p, C.c_int64]
        L.orc_pg_size.restype = C.c_int64
        L.orc_pg_size.argtypes = [C.c_void_p, C.c_int64]
        L.orc_pg_rows.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_pg_free.argtypes = [C.c_void_p]
        L.orc_join.restype = C.c_void_p
        L.orc_join.argtypes = [C.POINTER(_Col), C.POINTER(_Col), C.c_int]
        L.orc_join_len.restype = C.c_int64
        L.orc_join_len.argtypes = [C.c_void_p]
        L.orc_join_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_join_free.argtypes = [C.c_void_p]
        L.orc_gather.argtypes = [C.POINTER(_Col), C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_filter_indices.restype = C.c_int64
        L.orc_filter_indices.argtypes = [C.POINTER(_Col), C.c_void_p]
        L.orc_ideal_groupby_checksum.restype = C.c_double
        L.orc_ideal_groupby_checksum.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_int64)]
        L.orc_splitmix64.restype = C.c_uint64
        L.orc_splitmix64.argtypes = [C.c_uint64]
        L.orc_synth_keys.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_uint64, C.c_int]
        L.orc_synth_vals.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64]
        L.orc_synth_nulls.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_uint32]
        L.orc_synth_join_keys.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_uint64, C.c_int]
        # typed_oracle.cpp
        L.orc_typed_groupby.restype = C.c_void_p
        L.orc_typed_groupby.argtypes = [C.POINTER(_Col), C.POINTER(C.c_int64), C.c_int, C.POINTER(_Col), C.POINTER(_Col), C.c_int, C.c_int64, C.c_int]
        L.orc_typed_groupby_synth.restype = C.c_void_p
        L.orc_typed_groupby_synth.argtypes = [C.c_int64, C.c_int64, C.c_uint64, C.c_uint64, C.c_int, C.c_uint32, C.c_int]
        L.orc_tg_ngroups.restype = C.c_int64
        L.orc_tg_ngroups.argtypes = [C.c_void_p]
        L.orc_tg_key.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_tg_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_tg_aggs.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.orc_tg_free.argtypes = [C.c_void_p]
        L.orc_typed_join.restype = C.c_void_p
        L.orc_typed_join.argtypes = [C.POINTER(_Col), C.POINTER(_JSynth), C.POINTER(_Col), C.POINTER(_JSynth), C.c_int, C.c_int, C.c_int]
        L.orc_tj_len.restype = C.c_int64
        L.orc_tj_len.argtypes = [C.c_void_p]
        L.orc_tj_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_tg_exact.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.orc_tj_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_tj_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class Col:
    """A pandrs-layout column: typed data + optional null bitmap (bit set = NULL, LSB first)."""

    def __init__(self, dtype, data, nulls=None, pool=None, length=None):
        self.dtype = dtype
        self.data = np.ascontiguousarray(data, dtype=_NP[dtype])
        self.len = int(length if length is not None else (len(self.data) if dtype != BOOL_BITS else len(self.data) * 8))
        self.nulls = None if nulls is None else np.ascontiguousarray(nulls, dtype=np.uint8)
        self.pool = pool
        self._pool_c = None
        if pool is not None:
            self._pool_c = (C.c_char_p * len(pool))(*[s.encode() for s in pool])

    def c(self) -> _Col:
        return _Col(self.dtype, 0, self.data.ctypes.data if self.data.size else None,
                    self.nulls.ctypes.data if self.nulls is not None and self.nulls.size else None,
                    0 if self.nulls is None else self.nulls.size, self.len,
                    C.cast(self._pool_c, C.POINTER(C.c_char_p)) if self._pool_c is not None else None,
                    0 if self.pool is None else len(self.pool))


def pack_bits(flags) -> np.ndarray:
    """create_bitmask (core/column.rs:163-177): LSB-first packing of a bool vector."""
    return np.packbits(np.asarray(flags, dtype=bool), bitorder="little")


def groupby(keys, vals, aggs, mode=MODE_AGGREGATE, nthreads=1, want_key_strings=True):
    """keys/vals: lists of Col; aggs: list of (value_col_index, op).

    Returns dict(n_groups, first_row, group_rows, aggs=[np.float64 arrays], key_strings=[tuple,...], error).
    Group order = first appearance."""
    L = lib()
    nrows = keys[0].len if keys else (vals[0].len if vals else 0)
    kc = (_Col * max(1, len(keys)))(*[k.c() for k in keys])
    vc = (_Col * max(1, len(vals)))(*[v.c() for v in vals])
    ac = (C.c_int32 * max(1, len(aggs)))(*[a[0] for a in aggs])
    ao = (C.c_int32 * max(1, len(aggs)))(*[a[1] for a in aggs])
    h = L.orc_groupby(kc, len(keys), vc, ac, ao, len(aggs), nrows, mode, nthreads)
    try:
        G = L.orc_gb_ngroups(h)
        first = np.empty(G, np.int64)
        rows = np.empty(G, np.int64)
        L.orc_gb_first_rows(h, first.ctypes.data)
        L.orc_gb_group_rows(h, rows.ctypes.data)
        out = []
        for a in range(len(aggs)):
            v = np.empty(G, np.float64)
            L.orc_gb_agg(h, a, v.ctypes.data)
            out.append(v)
        ks = None
        if want_key_strings:
            ks = [tuple(L.orc_gb_key(h, g, k).decode() for k in range(len(keys))) for g in range(G)]
        return dict(n_groups=G, first_row=first, group_rows=rows, aggs=out, key_strings=ks, error=L.orc_gb_error(h))
    finally:
        L.orc_gb_free(h)


def par_groupby(keys) -> dict:
    """grouping.rs:124-331: {label: ascending row ids}; label = key parts joined with "_", NULL -> "NA"."""
    L = lib()
    kc = (_Col * max(1, len(keys)))(*[k.c() for k in keys])
    h = L.orc_par_groupby(kc, len(keys), keys[0].len)
    try:
        out = {}
        for g in range(L.orc_pg_ngroups(h)):
            rows = np.empty(L.orc_pg_size(h, g), np.int64)
            L.orc_pg_rows(h, g, rows.ctypes.data)
            out[L.orc_pg_label(h, g).decode()] = rows
        return out
    finally:
        L.orc_pg_free(h)


def join(left: Col, right: Col, how=INNER):
    L = lib()
    lc, rc = left.c(), right.c()
    h = L.orc_join(C.byref(lc), C.byref(rc), how)
    try:
        n = L.orc_join_len(h)
        li = np.empty(n, np.int64)
        ri = np.empty(n, np.int64)
        L.orc_join_pairs(h, li.ctypes.data, ri.ctypes.data)
        return li, ri
    finally:
        L.orc_join_free(h)


def gather(col: Col, idx) -> np.ndarray:
    idx = np.ascontiguousarray(idx, np.int64)
    out = np.empty(len(idx), _NP[col.dtype])
    cc = col.c()
    lib().orc_gather(C.byref(cc), idx.ctypes.data, len(idx), out.ctypes.data)
    return out


def filter_indices(mask: Col) -> np.ndarray:
    out = np.empty(mask.len, np.int64)
    cc = mask.c()
    n = lib().orc_filter_indices(C.byref(cc), out.ctypes.data)
    return out[:n].copy()


def synth_keys(n, seed=42, card=1000, scramble=False, row0=0):
    out = np.empty(n, np.int64)
    lib().orc_synth_keys(out.ctypes.data, n, row0, seed, card, int(scramble))
    return out


def synth_vals(n, seed=42, row0=0):
    out = np.empty(n, np.float64)
    lib().orc_synth_vals(out.ctypes.data, n, row0, seed)
    return out


def synth_nulls(n, seed=42, per_million=50000, row0=0):
    out = np.empty((n + 7) // 8, np.uint8)
    lib().orc_synth_nulls(out.ctypes.data, n, row0, seed, per_million)
    return out


def synth_join_keys(n, seed=42, domain=1, unique=False, row0=0):
    out = np.empty(n, np.int64)
    lib().orc_synth_join_keys(out.ctypes.data, n, row0, seed, domain, int(unique))
    return out


def ideal_groupby(keys, vals, vnull=None, nthreads=1):
    ng = C.c_int64(0)
    cs = lib().orc_ideal_groupby_checksum(keys.ctypes.data, vals.ctypes.data,
                                          None if vnull is None else vnull.ctypes.data, len(keys), nthreads, C.byref(ng))
    return cs, ng.value


# ---------------------------------------------------------------- typed-key oracle (typed_oracle.cpp)
# Same results as groupby() / join() above, bit for bit (tests/test_oracle_golden.py), at 100+ M rows/s: the checker
# for the CUDA path at BASELINE.json's sizes.
def _tg_result(h, nkeys):
    L = lib()
    try:
        G = L.orc_tg_ngroups(h)
        keys = []
        for k in range(nkeys):
            v = np.empty(G, np.uint64)
            isn = np.empty(G, np.uint8)
            L.orc_tg_key(h, k, v.ctypes.data, isn.ctypes.data)
            keys.append((v, isn.astype(bool)))
        rows, nv, first = np.empty(G, np.int64), np.empty(G, np.int64), np.empty(G, np.int64)
        L.orc_tg_rows(h, rows.ctypes.data, nv.ctypes.data, first.ctypes.data)
        a = [np.empty(G, np.float64) for _ in range(6)]
        L.orc_tg_aggs(h, *[x.ctypes.data for x in a])
        x = [np.empty(G, np.float64) for _ in range(4)]
        L.orc_tg_exact(h, *[y.ctypes.data for y in x])
        return dict(n_groups=G, keys=keys, group_rows=rows, valid_n=nv, first_row=first,
                    sum=a[0], mean=a[1], min=a[2], max=a[3], std=a[4], var=a[5],
                    exact=dict(sum=x[0], mean=x[1], std=x[2], var=x[3]))
    finally:
        L.orc_tg_free(h)


def typed_groupby(keys, val=None, filter=None, compat_nulls=False, null_alias=None, empty_id=0xFFFFFFFF, nthreads=None):
    """keys: list of Col, val: one Col or None.  Returns dict(n_groups, keys=[(u64 values, isnull)], group_rows, valid_n,
    first_row, sum, mean, min, max, std, var) - group order unspecified.  Key values are the physical bits as u64
    (i64 / i32 sign-extended, f64 bit pattern with all NaNs canonical, dictionary ids, 0 / 1 for bools)."""
    L = lib()
    nthreads = nthreads or os.cpu_count() or 1
    kc = (_Col * len(keys))(*[k.c() for k in keys])
    na = None
    if null_alias is not None:
        na = (C.c_int64 * len(keys))(*[int(x) for x in null_alias])
    vc = val.c() if val is not None else None
    fc = filter.c() if filter is not None else None
    h = L.orc_typed_groupby(kc, na, len(keys), C.byref(vc) if vc is not None else None, C.byref(fc) if fc is not None else None,
                            int(compat_nulls), int(empty_id), nthreads)
    return _tg_result(h, len(keys))


def typed_groupby_synth(n, seed=42, card=1000, scramble=False, null_per_million=0, row0=0, nthreads=None):
    """BASELINE.json configs[1] without materialising the columns: key = synth_keys, value = synth_vals, NULLs = synth_nulls."""
    h = lib().orc_typed_groupby_synth(n, row0, seed, card, int(scramble), null_per_million, nthreads or os.cpu_count() or 1)
    return _tg_result(h, 1)


def typed_join(left=None, right=None, how=INNER, left_synth=None, right_synth=None, want_pairs=False, nthreads=None):
    """left / right: Col, or *_synth = dict(n, seed, domain, unique) for synth_join_keys sides.  Returns dict(n, checksum,
    sum_left, sum_right, unmatched_left[, left, right]); checksum = sum over pairs of (l * A) ^ (r * B) mod 2^64."""
    L = lib()

    def synth(d):
        return _JSynth(int(d["n"]), int(d.get("seed", 42)), int(d.get("domain", 1)), int(bool(d.get("unique", False))))
    lc = left.c() if left is not None else None
    rc = right.c() if right is not None else None
    ls = synth(left_synth) if left_synth is not None else None
    rs = synth(right_synth) if right_synth is not None else None
    h = L.orc_typed_join(C.byref(lc) if lc is not None else None, C.byref(ls) if ls is not None else None,
                         C.byref(rc) if rc is not None else None, C.byref(rs) if rs is not None else None, how, int(want_pairs),
                         nthreads or os.cpu_count() or 1)
    try:
        n = L.orc_tj_len(h)
        cs, cu, sl, sr, ul = C.c_uint64(), C.c_uint64(), C.c_int64(), C.c_int64(), C.c_int64()
        L.orc_tj_stats(h, C.byref(cs), C.byref(cu), C.byref(sl), C.byref(sr), C.byref(ul))
        # a Left join of unique build keys holds the Inner join inside it: pairs with r >= 0
        out = dict(n=n, checksum=cs.value, checksum_unmatched=cu.value, sum_left=sl.value, sum_right=sr.value, unmatched_left=ul.value)
        if want_pairs:
            out["left"], out["right"] = np.empty(n, np.int64), np.empty(n, np.int64)
            L.orc_tj_pairs(h, out["left"].ctypes.data, out["right"].ctypes.data)
        return out
    finally:
        L.orc_tj_free(h)


PAIR_MIX_A, PAIR_MIX_B = 0x9E3779B97F4A7C15, 0xC2B2AE3D27D4EB4F


def pair_checksum(left, right) -> int:
    """The checksum typed_join reports, from numpy index arrays."""
    with np.errstate(over="ignore"):
        return int(((np.asarray(left).astype(np.uint64) * np.uint64(PAIR_MIX_A)) ^ (np.asarray(right).astype(np.uint64) * np.uint64(PAIR_MIX_B))).sum(dtype=np.uint64))


# ---------------------------------------------------------------- the LEGACY DataFrame::groupby (string-materialised Series)
_LEGACY_F64 = None


def _rust_parse_f64(s: str):
    """`str::parse::<f64>()`: sign, digits with optional fraction / exponent, or inf / infinity / nan in any case; nothing else."""
    global _LEGACY_F64
    if _LEGACY_F64 is None:
        import re
        _LEGACY_F64 = re.compile(r"^[+-]?((\d+\.?\d*|\.\d+)([eE][+-]?\d+)?|inf|infinity|nan)$", re.IGNORECASE)
    return float(s) if _LEGACY_F64.match(s) else None


def rust_f64_to_string(v: float) -> str:
    if np.isnan(v):
        return "NaN"
    if np.isinf(v):
        return "-inf" if v < 0 else "inf"
    return np.format_float_positional(v, trim="-")


def legacy_groupby(columns: dict, by, aggs) -> dict:
    """Pure-Python restatement (small inputs only) of src/dataframe/groupby.rs: DataFrameGroupBy::new (:196-228, groups keyed by
    the key STRINGS), agg (:258-300) and calculate_aggregation (:443-532: the value cells of a group parsed as f64, unparseable ones
    skipped; no parseable cell -> 0.0; Count = parseable cells; Std / Var two-pass with n - 1, <= 1 value -> 0.0; Median of the
    sorted values).  columns: {name: [str, ...]}; by: [names]; aggs: [(column, func name, alias)].
    Returns {key tuple: {alias: result string}} with the results formatted like `agg_result.to_string()` (:291)."""
    n = len(next(iter(columns.values()))) if columns else 0
    groups = {}
    for row in range(n):
        groups.setdefault(tuple(columns[c][row] for c in by), []).append(row)
    out = {}
    for key, rows in groups.items():
        rec = {}
        for col, func, alias in aggs:
            vals = [p for p in (_rust_parse_f64(columns[col][r]) for r in rows) if p is not None]
            if not vals:
                r = 0.0
            elif func == "sum":
                r = 0.0
                for x in vals:
                    r += x
            elif func == "mean":
                s = 0.0
                for x in vals:
                    s += x
                r = s / len(vals)
            elif func == "min":
                r = float("inf")
                for x in vals:
                    r = r if x != x else (x if r != r else min(r, x))          # f64::min ignores a NaN operand
            elif func == "max":
                r = float("-inf")
                for x in vals:
                    r = r if x != x else (x if r != r else max(r, x))
            elif func == "count":
                r = float(len(vals))
            elif func in ("std", "var"):
                if len(vals) <= 1:
                    r = 0.0
                else:
                    s = 0.0
                    for x in vals:
                        s += x
                    m = s / len(vals)
                    q = 0.0
                    for x in vals:
                        q += (x - m) * (x - m)
                    r = q / (len(vals) - 1)
                    if func == "std":
                        r = float(np.sqrt(r))
            elif func == "median":
                sv = sorted(vals)
                mid = len(sv) // 2
                r = (sv[mid - 1] + sv[mid]) / 2.0 if len(sv) % 2 == 0 else sv[mid]
            else:
                raise ValueError(func)
            rec[alias] = rust_f64_to_string(r)
        out[key] = rec
    return out
