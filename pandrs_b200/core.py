"""Thin object layer over the C ABI: Context, Column (pandrs column layout), results.

Stands in for the Rust host side (src/optimized/split_dataframe/{group,join}.rs calling the shim); every
compute call goes through libpandrs_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _native as N
from ._native import (BOOL_BITS, COUNT, DICT_U32, F64, I32, I64, INNER, LEFT, MAX, MEAN, MEM_DEVICE, MEM_HOST, MIN, STD, SUM,
                      VAR)

NP_DTYPE = {I64: np.int64, F64: np.float64, DICT_U32: np.uint32, BOOL_BITS: np.uint8, I32: np.int32}
KEY_OUT_DTYPE = {I64: np.int64, F64: np.float64, DICT_U32: np.uint32, BOOL_BITS: np.uint8, I32: np.int32}
OP_NAMES = {SUM: "sum", MEAN: "mean", MIN: "min", MAX: "max", COUNT: "count", STD: "std", VAR: "var"}


class PandrsError(RuntimeError):
    """Mirrors pandrs's Error enum for this path (core/error.rs): .kind names the variant."""

    KINDS = {N.ERR_BAD_ARG: "InvalidInput", N.ERR_TYPE_MISMATCH: "ColumnTypeMismatch", N.ERR_OOM: "Computation(OOM)",
             N.ERR_CUDA: "Computation", N.ERR_NCCL: "Computation(NCCL)", N.ERR_UNSUPPORTED: "OperationFailed"}

    def __init__(self, code: int, msg: str):
        self.code = code
        self.kind = self.KINDS.get(code, "Unknown")
        super().__init__(f"{self.kind} ({code}): {msg}")


def pack_bits(flags) -> np.ndarray:
    """create_bitmask (core/column.rs:163-177): LSB-first packing of a bool vector."""
    return np.packbits(np.asarray(flags, dtype=bool), bitorder="little")


class Column:
    """One pandrs-layout column: typed data + optional null bitmap (bit set = NULL, LSB first).

    Host columns hold numpy arrays; device columns hold raw device pointers (from Context.upload,
    Context.dev_alloc or e.g. a torch tensor's data_ptr())."""

    def __init__(self, dtype: int, data=None, nulls=None, length: Optional[int] = None, null_alias: int = -1,
                 device_ptr: Optional[int] = None, nulls_ptr: Optional[int] = None, null_len: Optional[int] = None, owner=None):
        self.dtype = dtype
        self.null_alias = null_alias
        self._owner = owner            # keeps device memory / tensors alive
        self._uploaded = None          # PdrsCol returned by pdrs_col_upload (freed by Context.free)
        if device_ptr is not None:
            self.mem = MEM_DEVICE
            self.ptr = device_ptr
            self.nulls_ptr = nulls_ptr or None
            self.len = int(length)
            self.null_len = int(null_len if null_len is not None else ((self.len + 7) // 8 if nulls_ptr else 0))
            self.data = None
            self.nulls = None
        else:
            self.mem = MEM_HOST
            self.data = np.ascontiguousarray(data, dtype=NP_DTYPE[dtype])
            self.nulls = None if nulls is None else np.ascontiguousarray(nulls, dtype=np.uint8)
            self.len = int(length if length is not None else (len(self.data) * 8 if dtype == BOOL_BITS else len(self.data)))
            self.ptr = self.data.ctypes.data if self.data.size else None
            self.nulls_ptr = self.nulls.ctypes.data if self.nulls is not None and self.nulls.size else None
            self.null_len = 0 if self.nulls is None else int(self.nulls.size)

    def c(self) -> N.PdrsCol:
        return N.PdrsCol(self.dtype, self.mem, self.ptr, self.nulls_ptr, self.null_len, self.len, self.null_alias)

    # -- constructors mirroring Int64Column::new / with_nulls etc. (src/column/*.rs)
    @staticmethod
    def int64(values, nulls=None):
        return Column(I64, values, None if nulls is None else pack_bits(nulls))

    @staticmethod
    def float64(values, nulls=None):
        return Column(F64, values, None if nulls is None else pack_bits(nulls))

    @staticmethod
    def int32(values, nulls=None):
        return Column(I32, values, None if nulls is None else pack_bits(nulls))

    @staticmethod
    def dict_ids(ids, nulls=None, null_alias: int = -1):
        return Column(DICT_U32, ids, None if nulls is None else pack_bits(nulls), null_alias=null_alias)

    @staticmethod
    def boolean(values, nulls=None):
        v = np.asarray(values, dtype=bool)
        return Column(BOOL_BITS, pack_bits(v), None if nulls is None else pack_bits(nulls), length=len(v))


class GroupByResult:
    """Owns a pdrs_groupby_result; arrays are copied to the host lazily."""

    def __init__(self, ctx: "Context", handle, key_dtypes, nvals, naggs):
        self.ctx, self._h, self.key_dtypes, self.nvals, self.naggs = ctx, handle, key_dtypes, nvals, naggs
        self.n_groups = int(ctx.L.pdrs_groupby_n_groups(handle))

    def key(self, k: int):
        out = np.empty(self.n_groups, KEY_OUT_DTYPE[self.key_dtypes[k]])
        isnull = np.empty(self.n_groups, np.uint8)
        if self.n_groups:
            self.ctx._chk(self.ctx.L.pdrs_groupby_key(self._h, k, out.ctypes.data, isnull.ctypes.data))
        return out, isnull.astype(bool)

    def agg(self, a: int) -> np.ndarray:
        out = np.empty(self.n_groups, np.float64)
        if self.n_groups:
            self.ctx._chk(self.ctx.L.pdrs_groupby_agg_values(self._h, a, out.ctypes.data))
        return out

    def group_rows(self) -> np.ndarray:
        out = np.empty(self.n_groups, np.int64)
        if self.n_groups:
            self.ctx._chk(self.ctx.L.pdrs_groupby_group_rows(self._h, out.ctypes.data))
        return out

    def valid_n(self, v: int) -> np.ndarray:
        out = np.empty(self.n_groups, np.int64)
        if self.n_groups:
            self.ctx._chk(self.ctx.L.pdrs_groupby_valid_n(self._h, v, out.ctypes.data))
        return out

    # device pointers (valid until close())
    def key_dev(self, k): return self.ctx.L.pdrs_groupby_key_dev(self._h, k)
    def key_null_dev(self, k): return self.ctx.L.pdrs_groupby_key_null_dev(self._h, k)
    def agg_dev(self, a): return self.ctx.L.pdrs_groupby_agg_dev(self._h, a)
    def rows_dev(self): return self.ctx.L.pdrs_groupby_group_rows_dev(self._h)
    def states_dev(self, v): return self.ctx.L.pdrs_groupby_states_dev(self._h, v)

    def close(self):
        if self._h:
            self.ctx.L.pdrs_groupby_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GroupRows:
    """Owns a pdrs_group_rows: the row lists of every group (par_groupby, grouping.rs:124-331) and the order-dependent
    aggregates over them (Median / First / Last, aggregation.rs:585-624, 703-742)."""

    def __init__(self, ctx: "Context", handle, key_dtypes):
        self.ctx, self._h, self.key_dtypes = ctx, handle, key_dtypes
        self.n_groups = int(ctx.L.pdrs_group_rows_n_groups(handle))
        self.n_rows = int(ctx.L.pdrs_group_rows_n_rows(handle))

    def key(self, k: int):
        out = np.empty(self.n_groups, KEY_OUT_DTYPE[self.key_dtypes[k]])
        isnull = np.empty(self.n_groups, np.uint8)
        if self.n_groups:
            self.ctx._chk(self.ctx.L.pdrs_group_rows_key(self._h, k, out.ctypes.data, isnull.ctypes.data))
        return out, isnull.astype(bool)

    def offsets(self) -> np.ndarray:
        out = np.zeros(self.n_groups + 1, np.int64)
        if self.n_groups:
            self.ctx._chk(self.ctx.L.pdrs_group_rows_offsets(self._h, out.ctypes.data))
        return out

    def ids(self) -> np.ndarray:
        out = np.empty(self.n_rows, np.int64)
        if self.n_rows:
            self.ctx._chk(self.ctx.L.pdrs_group_rows_ids(self._h, out.ctypes.data))
        return out

    def ids_dev(self): return self.ctx.L.pdrs_group_rows_ids_dev(self._h)
    def offsets_dev(self): return self.ctx.L.pdrs_group_rows_offsets_dev(self._h)

    def agg(self, val: "Column", op: int) -> np.ndarray:
        out = np.zeros(self.n_groups, np.float64)
        vc = val.c()
        self.ctx._chk(self.ctx.L.pdrs_group_rows_agg(self._h, C.byref(vc), int(op), out.ctypes.data if self.n_groups else None))
        return out

    def close(self):
        if self._h:
            self.ctx.L.pdrs_group_rows_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DictEncoded:
    """Owns a pdrs_dict: dictionary ids of an Arrow string array built on the device (string_pool.rs:28-52)."""

    def __init__(self, ctx: "Context", handle, length: int):
        self.ctx, self._h, self.len = ctx, handle, int(length)
        self.n_unique = int(ctx.L.pdrs_dict_n_unique(handle))

    def ids(self) -> np.ndarray:
        out = np.empty(self.len, np.uint32)
        if self.len:
            self.ctx._chk(self.ctx.L.pdrs_dict_ids(self._h, out.ctypes.data))
        return out

    def first_rows(self) -> np.ndarray:
        out = np.empty(self.n_unique, np.int64)
        if self.n_unique:
            self.ctx._chk(self.ctx.L.pdrs_dict_first_rows(self._h, out.ctypes.data))
        return out

    def ids_dev(self): return self.ctx.L.pdrs_dict_ids_dev(self._h)
    def nulls_dev(self): return self.ctx.L.pdrs_dict_nulls_dev(self._h)

    def remap(self, new_ids):
        m = np.ascontiguousarray(new_ids, np.uint32)
        assert len(m) == self.n_unique
        self.ctx._chk(self.ctx.L.pdrs_dict_remap(self._h, m.ctypes.data if m.size else None))

    def close(self):
        if self._h:
            self.ctx.L.pdrs_dict_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class JoinResult:
    def __init__(self, ctx: "Context", handle, right_dtypes=()):
        self.ctx, self._h = ctx, handle
        self.n = int(ctx.L.pdrs_join_len(handle))
        self.right_dtypes = list(right_dtypes)

    def right_col(self, k: int) -> np.ndarray:
        """Column k of the right frame materialised along the pairs (pdrs_join_gather)."""
        out = np.empty(self.n, NP_DTYPE[self.right_dtypes[k]])
        if self.n:
            self.ctx._chk(self.ctx.L.pdrs_join_right_col(self._h, k, out.ctypes.data))
        return out

    def right_col_dev(self, k: int): return self.ctx.L.pdrs_join_right_col_dev(self._h, k)

    def indices(self):
        li = np.empty(self.n, np.int64)
        ri = np.empty(self.n, np.int64)
        if self.n:
            self.ctx._chk(self.ctx.L.pdrs_join_indices(self._h, li.ctypes.data, ri.ctypes.data))
        return li, ri

    def left_dev(self): return self.ctx.L.pdrs_join_left_dev(self._h)
    def right_dev(self): return self.ctx.L.pdrs_join_right_dev(self._h)

    def close(self):
        if self._h:
            self.ctx.L.pdrs_join_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class XJoin:
    """pdrs_xjoin: this rank's end of the fused partition + shuffle join (include/pandrs_b200.h)."""

    def __init__(self, ctx: "Context", rank: int, world: int, max_left_rows: int, max_right_rows: int, total_right_rows: int):
        self.ctx, self.rank, self.world = ctx, rank, world
        h = C.c_void_p()
        ctx._chk(ctx.L.pdrs_xjoin_create(ctx._h, rank, world, int(max_left_rows), int(max_right_rows), int(total_right_rows), C.byref(h)))
        self._h = h
        self.bytes = int(ctx.L.pdrs_xjoin_bytes(h))
        self.base = ctx.L.pdrs_xjoin_base(h)

    def ipc_handle(self) -> bytes:
        buf = (C.c_uint8 * 64)()
        self.ctx._chk(self.ctx.L.pdrs_xjoin_ipc_handle(self._h, buf))
        return bytes(buf)

    def attach_ipc(self, handles: Sequence[bytes]):
        blob = b"".join(handles)
        assert len(blob) == 64 * self.world
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        self.ctx._chk(self.ctx.L.pdrs_xjoin_attach_ipc(self._h, buf))

    def attach_ptrs(self, bases: Sequence[int]):
        arr = (C.c_void_p * self.world)(*bases)
        self.ctx._chk(self.ctx.L.pdrs_xjoin_attach_ptrs(self._h, arr))

    def shuffle(self, left: Column, right: Column, right_row0: int):
        lc, rc = left.c(), right.c()
        self.ctx._chk(self.ctx.L.pdrs_xjoin_shuffle(self._h, C.byref(lc), C.byref(rc), int(right_row0)))

    def local(self, how: int, left_row0: Sequence[int]) -> JoinResult:
        arr = (C.c_int64 * self.world)(*[int(v) for v in left_row0])
        h = C.c_void_p()
        self.ctx._chk(self.ctx.L.pdrs_xjoin_local(self._h, how, arr, C.byref(h)))
        return JoinResult(self.ctx, h)

    def close(self):
        if self._h:
            self.ctx.L.pdrs_xjoin_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            if self.ctx._h:
                self.close()
        except Exception:
            pass


class Context:
    """pdrs_ctx: one device + one stream, not re-entrant (include/pandrs_b200.h)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None, groupby_algo: int = N.GB_AUTO, groups_hint: int = 0,
                 compat_filter_nulls: bool = False):
        self.L = N.lib()
        opts = N.PdrsOptions(device, groupby_algo, groups_hint, stream, int(compat_filter_nulls), 0)
        h = C.c_void_p()
        rc = self.L.pdrs_ctx_create(C.byref(opts), C.byref(h))
        if rc != 0:
            raise PandrsError(rc, (self.L.pdrs_last_error(None) or b"").decode())
        self._h = h
        self.device = device

    def _chk(self, rc: int):
        if rc != 0:
            raise PandrsError(rc, (self.L.pdrs_last_error(self._h) or b"").decode())

    def close(self):
        if self._h:
            self.L.pdrs_ctx_destroy(self._h)
            self._h = None

    def set_option(self, name: str, value: int):
        self._chk(self.L.pdrs_set_option(self._h, name.encode(), int(value)))

    def stats(self) -> dict:
        s = N.PdrsStats()
        self._chk(self.L.pdrs_get_stats(self._h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in N.PdrsStats._fields_}

    def sync(self):
        self._chk(self.L.pdrs_sync(self._h))

    # ---- memory
    def upload(self, col: Column) -> Column:
        src, dst = col.c(), N.PdrsCol()
        self._chk(self.L.pdrs_col_upload(self._h, C.byref(src), C.byref(dst)))
        out = Column(col.dtype, device_ptr=dst.data, nulls_ptr=dst.null_bits, null_len=dst.null_len, length=col.len, null_alias=col.null_alias)
        out._uploaded = dst
        return out

    def free(self, col: Column):
        if col._uploaded is not None:
            self._chk(self.L.pdrs_col_free(self._h, C.byref(col._uploaded)))
            col._uploaded = None

    def dev_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._chk(self.L.pdrs_dev_alloc(self._h, int(nbytes), C.byref(p)))
        return p.value

    def dev_free(self, ptr: int):
        self._chk(self.L.pdrs_dev_free(self._h, ptr))

    def host_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._chk(self.L.pdrs_host_alloc(self._h, int(nbytes), C.byref(p)))
        return p.value

    def host_free(self, ptr: int):
        self._chk(self.L.pdrs_host_free(self._h, ptr))

    def memcpy(self, dst: int, src: int, nbytes: int, kind: int):
        self._chk(self.L.pdrs_memcpy(self._h, dst, src, int(nbytes), kind))

    def to_host(self, ptr: int, count: int, dtype) -> np.ndarray:
        out = np.empty(count, dtype)
        if count:
            self.memcpy(out.ctypes.data, ptr, out.nbytes, 1)
        return out

    def flush_l2(self):
        self._chk(self.L.pdrs_flush_l2(self._h))

    def timer_begin(self):
        self._chk(self.L.pdrs_timer_begin(self._h))

    def timer_end(self) -> float:
        ms = C.c_float()
        self._chk(self.L.pdrs_timer_end(self._h, C.byref(ms)))
        return ms.value

    # ---- synthetic device columns (same arithmetic as the oracle's generators)
    def synth_keys(self, n, seed=42, card=1000, scramble=False, row0=0) -> Column:
        p = self.dev_alloc(max(n, 1) * 8 + 64)
        self._chk(self.L.pdrs_synth_keys(self._h, p, n, row0, seed, card, int(scramble)))
        return Column(I64, device_ptr=p, length=n, owner=_DevOwner(self, p))

    def synth_vals(self, n, seed=42, row0=0, null_per_million: int = 0) -> Column:
        p = self.dev_alloc(max(n, 1) * 8 + 64)
        self._chk(self.L.pdrs_synth_vals(self._h, p, n, row0, seed))
        owner = [_DevOwner(self, p)]
        q, nl = None, 0
        if null_per_million:
            nl = ((n + 7) // 8 + 7) // 8 * 8
            q = self.dev_alloc(nl + 64)
            self._chk(self.L.pdrs_synth_nulls(self._h, q, n, row0, seed, null_per_million))
            owner.append(_DevOwner(self, q))
        return Column(F64, device_ptr=p, nulls_ptr=q, null_len=nl, length=n, owner=owner)

    def synth_join_keys(self, n, seed=42, domain=1, unique=False, row0=0) -> Column:
        p = self.dev_alloc(max(n, 1) * 8 + 64)
        self._chk(self.L.pdrs_synth_join_keys(self._h, p, n, row0, seed, domain, int(unique)))
        return Column(I64, device_ptr=p, length=n, owner=_DevOwner(self, p))

    # ---- the hot path
    @staticmethod
    def _cols(cols: Sequence[Column]):
        arr = (N.PdrsCol * max(1, len(cols)))(*[c.c() for c in cols])
        return arr

    def groupby_agg(self, keys: Sequence[Column], vals: Sequence[Column], aggs: Sequence[tuple], filter: Optional[Column] = None,
                    pred: Optional[tuple] = None) -> GroupByResult:
        """aggs: [(value_col_index, op)].  Replaces group_by(...).aggregate(...) (grouping.rs:38, aggregation.rs:763).
        pred: (column, cmp op, constant) - a typed row predicate evaluated inside the scan (pdrs_groupby_agg_where)."""
        ka, va = self._cols(keys), self._cols(vals)
        aa = (N.PdrsAgg * max(1, len(aggs)))(*[N.PdrsAgg(int(v), int(op)) for v, op in aggs])
        f = filter.c() if filter is not None else None
        h = C.c_void_p()
        if pred is not None:
            col, op, const = pred
            pp = N.PdrsPred(col.c(), int(op), 0, int(const) if col.dtype == I64 else 0, float(const))
            self._chk(self.L.pdrs_groupby_agg_where(self._h, ka, len(keys), va, len(vals), aa, len(aggs), C.byref(f) if f is not None else None, C.byref(pp), C.byref(h)))
        else:
            self._chk(self.L.pdrs_groupby_agg(self._h, ka, len(keys), va, len(vals), aa, len(aggs), C.byref(f) if f is not None else None, C.byref(h)))
        return GroupByResult(self, h, [k.dtype for k in keys], len(vals), len(aggs))

    def groupby_rows(self, keys: Sequence[Column]) -> GroupRows:
        """Row lists per group: replaces par_groupby's grouping (grouping.rs:124-331) and GroupBy.groups for Median / First / Last."""
        ka = self._cols(keys)
        h = C.c_void_p()
        self._chk(self.L.pdrs_groupby_rows(self._h, ka, len(keys), C.byref(h)))
        return GroupRows(self, h, [k.dtype for k in keys])

    def arrow_validity_to_nulls(self, validity, length: int, bit_offset: int = 0):
        """Arrow validity bitmap (numpy uint8 array, or None = all valid) -> (pandrs null mask as numpy uint8, number of NULLs)."""
        out = np.zeros((length + 7) // 8, np.uint8)
        n = C.c_int64()
        v = None if validity is None else np.ascontiguousarray(validity, np.uint8)
        self._chk(self.L.pdrs_arrow_validity_to_nulls(self._h, v.ctypes.data if v is not None and v.size else None, MEM_HOST, int(bit_offset), int(length),
                                                      out.ctypes.data if out.size else None, MEM_HOST, C.byref(n)))
        return out, int(n.value)

    def dict_encode(self, offsets, data, validity=None, bit_offset: int = 0, length: Optional[int] = None) -> DictEncoded:
        """Arrow Utf8 / LargeUtf8 buffers (numpy int32 / int64 offsets, uint8 bytes, optional validity bitmap) -> dictionary ids."""
        off = np.ascontiguousarray(offsets)
        assert off.dtype in (np.int32, np.int64)
        n = int(length if length is not None else len(off) - 1)
        by = np.ascontiguousarray(data, np.uint8)
        v = None if validity is None else np.ascontiguousarray(validity, np.uint8)
        h = C.c_void_p()
        self._chk(self.L.pdrs_dict_encode(self._h, off.ctypes.data if off.size else None, int(off.dtype == np.int64), by.ctypes.data if by.size else None, by.size,
                                          v.ctypes.data if v is not None and v.size else None, int(bit_offset), n, MEM_HOST, C.byref(h)))
        return DictEncoded(self, h, n)

    def groupby_partial(self, keys, vals, filter=None, all_stats=True) -> GroupByResult:
        ka, va = self._cols(keys), self._cols(vals)
        f = filter.c() if filter is not None else None
        h = C.c_void_p()
        self._chk(self.L.pdrs_groupby_partial(self._h, ka, len(keys), va, len(vals), C.byref(f) if f is not None else None, int(all_stats), C.byref(h)))
        return GroupByResult(self, h, [k.dtype for k in keys], len(vals), 0)

    def groupby_merge(self, keys: Sequence[Column], states_ptrs: Sequence[int], val_is_int: Sequence[bool], n_state_rows: int, aggs) -> GroupByResult:
        ka = self._cols(keys)
        sp = (C.c_void_p * max(1, len(states_ptrs)))(*states_ptrs)
        vi = (C.c_int32 * max(1, len(states_ptrs)))(*[int(b) for b in val_is_int])
        aa = (N.PdrsAgg * max(1, len(aggs)))(*[N.PdrsAgg(int(v), int(op)) for v, op in aggs])
        h = C.c_void_p()
        self._chk(self.L.pdrs_groupby_merge(self._h, ka, len(keys), sp, vi, len(states_ptrs), n_state_rows, aa, len(aggs), C.byref(h)))
        return GroupByResult(self, h, [k.dtype for k in keys], len(states_ptrs), len(aggs))

    def hash_partition(self, keys: Sequence[Column], nparts: int, perm_dev: int) -> np.ndarray:
        ka = self._cols(keys)
        counts = (C.c_int64 * nparts)()
        self._chk(self.L.pdrs_hash_partition(self._h, ka, len(keys), nparts, perm_dev, counts))
        return np.array(list(counts), np.int64)

    def join_pairs(self, left: Column, right: Column, how: int = INNER) -> JoinResult:
        """Replaces the build/probe of join_impl (join.rs:107-208)."""
        lc, rc = left.c(), right.c()
        h = C.c_void_p()
        self._chk(self.L.pdrs_join_pairs(self._h, C.byref(lc), C.byref(rc), how, C.byref(h)))
        return JoinResult(self, h)

    def join_gather(self, left: Column, right: Column, how: int, right_cols: Sequence[Column]) -> JoinResult:
        """join_impl's build / probe plus the materialisation of the right frame's columns (join.rs:107-208, 290-552)."""
        lc, rc = left.c(), right.c()
        cols = self._cols(right_cols)
        h = C.c_void_p()
        self._chk(self.L.pdrs_join_gather(self._h, C.byref(lc), C.byref(rc), how, cols, len(right_cols), C.byref(h)))
        return JoinResult(self, h, [c.dtype for c in right_cols])

    def gather(self, col: Column, idx, n: Optional[int] = None, idx_dev: bool = False, out_dev: Optional[int] = None):
        """join.rs:290-552 / data_ops.rs:124-211: default-filled gather without a null mask."""
        cc = col.c()
        if idx_dev:
            ip, cnt = idx, int(n)
        else:
            idx = np.ascontiguousarray(idx, np.int64)
            ip, cnt = idx.ctypes.data if idx.size else None, len(idx)
        if out_dev is not None:
            self._chk(self.L.pdrs_gather(self._h, C.byref(cc), ip, MEM_DEVICE if idx_dev else MEM_HOST, cnt, out_dev, MEM_DEVICE))
            return out_dev
        out = np.empty(cnt, NP_DTYPE[col.dtype])
        self._chk(self.L.pdrs_gather(self._h, C.byref(cc), ip, MEM_DEVICE if idx_dev else MEM_HOST, cnt, out.ctypes.data if cnt else None, MEM_HOST))
        return out

    def filter_indices(self, mask: Column) -> np.ndarray:
        """data_ops.rs:37-62: ascending row ids where the Boolean column is Some(true)."""
        buf = self.dev_alloc(max(mask.len, 1) * 8)
        try:
            n = C.c_int64()
            mc = mask.c()
            self._chk(self.L.pdrs_filter_indices(self._h, C.byref(mc), buf, C.byref(n)))
            return self.to_host(buf, n.value, np.int64)
        finally:
            self.dev_free(buf)


class Comm:
    """pdrs_comm: this rank's end of the multi-GPU operators (one process per GPU, NCCL bound inside the library).

    `exchange_id(id_bytes_or_None) -> bytes` is the host's broadcast of rank 0's 128-byte NCCL id (torch.distributed, MPI ...);
    it is only called when world > 1."""
    REPLICATED, SHARDED, AUTO = 1, 2, 0

    def __init__(self, ctx: "Context", rank: int, world: int, broadcast_id=None):
        self.ctx, self.rank, self.world = ctx, rank, world
        ident = None
        if world > 1:
            mine = None
            if rank == 0:
                buf = (C.c_uint8 * 128)()
                rc = ctx.L.pdrs_comm_unique_id(buf)
                if rc != 0:
                    raise PandrsError(rc, (ctx.L.pdrs_last_error(None) or b"").decode())
                mine = bytes(buf)
            ident = broadcast_id(mine)
            assert len(ident) == 128
        h = C.c_void_p()
        idbuf = (C.c_uint8 * 128).from_buffer_copy(ident) if ident is not None else None
        ctx._chk(ctx.L.pdrs_comm_init(ctx._h, world, rank, idbuf, C.byref(h)))
        self._h = h

    def set_option(self, name: str, value: int):
        self.ctx._chk(self.ctx.L.pdrs_comm_set_option(self._h, name.encode(), int(value)))

    def barrier(self):
        self.ctx._chk(self.ctx.L.pdrs_comm_barrier(self._h))

    def last_exchange(self):
        ms, b = C.c_float(), C.c_int64()
        self.ctx._chk(self.ctx.L.pdrs_comm_last_exchange(self._h, C.byref(ms), C.byref(b)))
        return ms.value, b.value

    def groupby_agg(self, keys, vals, aggs, filter=None, pred=None, result_mode: int = 0) -> GroupByResult:
        """groupby over the union of the ranks' rows (pdrs_groupby_agg_dist); collective."""
        ctx = self.ctx
        ka, va = ctx._cols(keys), ctx._cols(vals)
        aa = (N.PdrsAgg * max(1, len(aggs)))(*[N.PdrsAgg(int(v), int(op)) for v, op in aggs])
        f = filter.c() if filter is not None else None
        pp = None
        if pred is not None:
            col, op, const = pred
            pp = N.PdrsPred(col.c(), int(op), 0, int(const) if col.dtype == I64 else 0, float(const))
        h = C.c_void_p()
        ctx._chk(ctx.L.pdrs_groupby_agg_dist(self._h, ka, len(keys), va, len(vals), aa, len(aggs), C.byref(f) if f is not None else None,
                                             C.byref(pp) if pp is not None else None, int(result_mode), C.byref(h)))
        nvs = sum(1 for v in vals if v.dtype in (I64, F64))
        return GroupByResult(ctx, h, [k.dtype for k in keys], nvs, len(aggs))

    def join_pairs(self, left: Column, right: Column, how: int, left_row0: int, right_row0: int, max_left_rows: int, max_right_rows: int,
                   total_right_rows: int) -> JoinResult:
        """Inner / Left join over the union of the ranks' rows (pdrs_join_pairs_dist); collective; GLOBAL row numbers."""
        lc, rc = left.c(), right.c()
        h = C.c_void_p()
        self.ctx._chk(self.ctx.L.pdrs_join_pairs_dist(self._h, C.byref(lc), C.byref(rc), how, int(left_row0), int(right_row0), int(max_left_rows),
                                                      int(max_right_rows), int(total_right_rows), C.byref(h)))
        return JoinResult(self.ctx, h)

    def close(self):
        if self._h:
            self.ctx.L.pdrs_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            if self.ctx._h:
                self.close()
        except Exception:
            pass


def torch_broadcast_id(dist, device):
    """broadcast_id callback for Comm on top of torch.distributed (any backend)."""
    import torch

    def f(mine):
        t = torch.zeros(128, dtype=torch.uint8, device=device)
        if mine is not None:
            t.copy_(torch.tensor(list(mine), dtype=torch.uint8))
        dist.broadcast(t, src=0)
        return bytes(t.cpu().tolist())
    return f


class _DevOwner:
    def __init__(self, ctx: Context, ptr: int):
        self.ctx, self.ptr = ctx, ptr

    def __del__(self):
        try:
            if self.ctx._h:
                self.ctx.dev_free(self.ptr)
        except Exception:
            pass
