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
SUM, MEAN, MIN, MAX, COUNT, STD, VAR = range(7)
MODE_AGGREGATE, MODE_PAR_AGGREGATE, MODE_LAZY = 0, 1, 2
INNER, LEFT, RIGHT, OUTER = 0, 1, 2, 3

_NP = {I64: np.int64, F64: np.float64, DICT_U32: np.uint32, BOOL_BITS: np.uint8, I32: np.int32}


class _Col(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("_pad", C.c_int32), ("data", C.c_void_p),
                ("null_bits", C.c_void_p), ("null_len", C.c_int64), ("len", C.c_int64),
                ("pool", C.POINTER(C.c_char_p)), ("pool_len", C.c_int64)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pandrs_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
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
