"""Host-side mirror of pandrs's OptimizedDataFrame / GroupBy / LazyFrame API for the groupby-aggregate and
inner / left join hot paths (same names, argument meaning and error behaviour), standing in for the Rust
methods that would call the C ABI:

  OptimizedDataFrame::group_by / group_by_with_options   src/optimized/split_dataframe/group/grouping.rs:38-115
  GroupBy::aggregate / par_aggregate / agg / sum / ...    group/aggregation.rs:22-182, 763-871; group/operations.rs:438-547
  LazyFrame::aggregate(...).execute()                     src/optimized/lazy.rs:186-404
  OptimizedDataFrame::inner_join / left_join              split_dataframe/join.rs:32-47, 76-555
  OptimizedDataFrame::filter                              split_dataframe/data_ops.rs:37-121
  OptimizedDataFrame::par_groupby                         split_dataframe/group/grouping.rs:124-331
  ArrowConverter::record_batch_to_dataframe               src/arrow_integration.rs:75-100, 160-225 (typed columns of this path)

Every compute step (grouping, aggregation, build / probe, gathers, filter indices) runs in libpandrs_b200.so;
this module only does what the Rust side does around the shim: schema checks, key re-stringification, result
frame assembly.  There is no CPU fallback: unsupported operations raise OperationFailed.
"""
from __future__ import annotations

import copy
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _native as N
from .core import Column as RawColumn
from .core import Context, PandrsError

__all__ = ["AggregateOp", "JoinType", "ColumnType", "Int64Column", "Float64Column", "StringColumn", "BooleanColumn",
           "OptimizedDataFrame", "GroupBy", "LazyFrame", "ColumnNotFound", "ColumnTypeMismatch", "OperationFailed",
           "DuplicateColumnName", "InconsistentRowCount", "get_context", "set_context", "GLOBAL_STRING_POOL"]


# ---------------------------------------------------------------- errors (src/core/error.rs)
class ColumnNotFound(KeyError):
    pass


class ColumnTypeMismatch(TypeError):
    pass


class OperationFailed(RuntimeError):
    pass


class DuplicateColumnName(ValueError):
    pass


class InconsistentRowCount(ValueError):
    pass


# ---------------------------------------------------------------- enums
class AggregateOp:
    """group/types.rs:11-34, same discriminants.  Median / First / Last are computed from the row lists of the groups
    (pdrs_groupby_rows + pdrs_group_rows_agg); Custom needs a host closure and is outside the accelerated path (no CPU fallback)."""
    Sum, Mean, Min, Max, Count, Std, Var = N.SUM, N.MEAN, N.MIN, N.MAX, N.COUNT, N.STD, N.VAR
    Median, First, Last, Custom = N.MEDIAN, N.FIRST, N.LAST, 10
    NAMES = {N.SUM: "sum", N.MEAN: "mean", N.MIN: "min", N.MAX: "max", N.COUNT: "count", N.STD: "std", N.VAR: "var",
             N.MEDIAN: "median", N.FIRST: "first", N.LAST: "last", 10: "custom"}


class JoinType:
    """split_dataframe/join.rs:11-20"""
    Inner, Left, Right, Outer = 0, 1, 2, 3


class ColumnType:
    """core/column.rs:9-14"""
    Int64, Float64, String, Boolean = "Int64", "Float64", "String", "Boolean"


# ---------------------------------------------------------------- the process-global string pool (column/string_pool.rs:6-52)
class _StringPool:
    def __init__(self):
        self.strings: List[str] = []
        self.index: Dict[str, int] = {}

    def get_or_insert(self, s: str) -> int:
        i = self.index.get(s)
        if i is None:
            i = len(self.strings)
            self.strings.append(s)
            self.index[s] = i
        return i

    def get(self, i: int) -> str:
        return self.strings[i] if 0 <= i < len(self.strings) else ""

    def __len__(self):
        return len(self.strings)


GLOBAL_STRING_POOL = _StringPool()


# ---------------------------------------------------------------- columns (src/column/*.rs)
class _TypedColumn:
    column_type = None
    dtype = None

    def __init__(self, raw: RawColumn, nulls: Optional[np.ndarray]):
        self.raw = raw
        self.nulls = nulls          # bool flags or None
        self._packed = None         # columns built from packed buffers (Arrow ingest): the pandrs null mask itself

    def __len__(self):
        return self.raw.len

    def is_null(self, i: int) -> bool:
        if self._packed is not None:
            return bool(i // 8 < len(self._packed) and (self._packed[i // 8] >> (i % 8)) & 1)
        return bool(self.nulls is not None and i < len(self.nulls) and self.nulls[i])

    @classmethod
    def _from_buffers(cls, dtype, data, packed_nulls, length, **kw):
        """A column over existing buffers (values / bit-packed Booleans / pool ids + pandrs null mask): nothing is re-packed."""
        self = cls.__new__(cls)
        _TypedColumn.__init__(self, RawColumn(dtype, data, packed_nulls, length=length, **kw), None)
        self._packed = packed_nulls
        return self


class Int64Column(_TypedColumn):
    column_type, dtype = ColumnType.Int64, N.I64

    def __init__(self, values, nulls=None):
        v = np.asarray(values, dtype=np.int64)
        n = None if nulls is None else np.asarray(nulls, dtype=bool)
        super().__init__(RawColumn.int64(v, n), n)
        self.values = v

    @staticmethod
    def with_nulls(values, nulls):
        return Int64Column(values, nulls)

    def get(self, i):
        return None if self.is_null(i) else int(self.values[i])


class Float64Column(_TypedColumn):
    column_type, dtype = ColumnType.Float64, N.F64

    def __init__(self, values, nulls=None):
        v = np.asarray(values, dtype=np.float64)
        n = None if nulls is None else np.asarray(nulls, dtype=bool)
        super().__init__(RawColumn.float64(v, n), n)
        self.values = v

    @staticmethod
    def with_nulls(values, nulls):
        return Float64Column(values, nulls)

    def get(self, i):
        return None if self.is_null(i) else float(self.values[i])


class StringColumn(_TypedColumn):
    """Dictionary-encoded through the global pool (string_column.rs:61-72): equal id <=> equal string."""
    column_type, dtype = ColumnType.String, N.DICT_U32

    def __init__(self, strings: Sequence[str], nulls=None, _ids=None):
        ids = np.asarray(_ids, dtype=np.uint32) if _ids is not None else np.fromiter((GLOBAL_STRING_POOL.get_or_insert(s) for s in strings), dtype=np.uint32, count=len(strings))
        n = None if nulls is None else np.asarray(nulls, dtype=bool)
        alias = GLOBAL_STRING_POOL.index.get("NULL", -1)     # a literal "NULL" merges with the NULL group (grouping.rs:69-98)
        super().__init__(RawColumn.dict_ids(ids, n, null_alias=alias), n)
        self.ids = ids

    def get(self, i):
        return None if self.is_null(i) else GLOBAL_STRING_POOL.get(int(self.ids[i]))

    def to_list(self):
        return [self.get(i) for i in range(len(self))]


class BooleanColumn(_TypedColumn):
    column_type, dtype = ColumnType.Boolean, N.BOOL_BITS

    def __init__(self, values, nulls=None):
        v = np.asarray(values, dtype=bool)
        n = None if nulls is None else np.asarray(nulls, dtype=bool)
        super().__init__(RawColumn.boolean(v, n), n)
        self.values = v

    def get(self, i):
        return None if self.is_null(i) else bool(self.values[i])


# ---------------------------------------------------------------- context
_CTX: Optional[Context] = None


def get_context() -> Context:
    """The process-wide pdrs_ctx (reference: global GpuManager behind a Mutex, src/gpu/mod.rs:249-251)."""
    global _CTX
    if _CTX is None:
        _CTX = Context(device=0)
    return _CTX


def set_context(ctx: Optional[Context]):
    global _CTX
    _CTX = ctx


def _f64_display(v: float) -> str:
    if np.isnan(v):
        return "NaN"
    if np.isinf(v):
        return "-inf" if v < 0 else "inf"
    return np.format_float_positional(v, trim="-")


def _key_strings(col: _TypedColumn, values: np.ndarray, isnull: np.ndarray) -> List[str]:
    """`Some(v) -> v.to_string()`, `None -> "NULL"` (grouping.rs:69-98)."""
    out = []
    for v, nl in zip(values, isnull):
        if nl:
            out.append("NULL")
        elif col.dtype in (N.I64, N.I32):
            out.append(str(int(v)))
        elif col.dtype == N.F64:
            out.append(_f64_display(float(v)))
        elif col.dtype == N.DICT_U32:
            out.append(GLOBAL_STRING_POOL.get(int(v)))
        else:
            out.append("true" if v else "false")
    return out


# ---------------------------------------------------------------- the frame
class OptimizedDataFrame:
    def __init__(self):
        self._cols: Dict[str, _TypedColumn] = {}
        self._order: List[str] = []
        self._rows = 0
        self.index = None            # list of key tuples when a multi-key groupby returns a StringMultiIndex

    @staticmethod
    def new():
        return OptimizedDataFrame()

    # -- construction (split_dataframe/core.rs add_column & friends)
    def add_column(self, name: str, col: _TypedColumn):
        if name in self._cols:
            raise DuplicateColumnName(name)
        if self._order and len(col) != self._rows:
            raise InconsistentRowCount(f"expected {self._rows} rows, column '{name}' has {len(col)}")
        self._cols[name] = col
        self._order.append(name)
        self._rows = len(col)
        return self

    def add_int_column(self, name, values):
        return self.add_column(name, Int64Column(values))

    def add_float_column(self, name, values):
        return self.add_column(name, Float64Column(values))

    def add_string_column(self, name, values):
        return self.add_column(name, StringColumn(values))

    def add_boolean_column(self, name, values):
        return self.add_column(name, BooleanColumn(values))

    def row_count(self) -> int:
        return self._rows

    def column_count(self) -> int:
        return len(self._order)

    def column_names(self) -> List[str]:
        return list(self._order)

    def contains_column(self, name: str) -> bool:
        return name in self._cols

    def column(self, name: str) -> _TypedColumn:
        if name not in self._cols:
            raise ColumnNotFound(name)
        return self._cols[name]

    def column_type(self, name: str) -> str:
        return self.column(name).column_type

    # -- grouping (grouping.rs:38-115)
    def group_by(self, columns: Sequence[str]) -> "GroupBy":
        return self.group_by_with_options(columns, True)

    def group_by_with_options(self, columns: Sequence[str], as_multi_index: bool) -> "GroupBy":
        if isinstance(columns, str):
            columns = [columns]
        for c in columns:
            if c not in self._cols:
                raise ColumnNotFound(c)
        return GroupBy(self, list(columns), as_multi_index)

    # -- filter (data_ops.rs:37-121): rows where the Boolean column is Some(true); NULLs in the kept rows become
    #    type defaults and the masks are dropped (data_ops.rs:64-71)
    def filter(self, condition_column: str) -> "OptimizedDataFrame":
        cond = self.column(condition_column)
        if cond.column_type != ColumnType.Boolean:
            raise ColumnTypeMismatch(f"column '{condition_column}': expected Boolean, found {cond.column_type}")
        ctx = get_context()
        idx = ctx.filter_indices(cond.raw)
        return self._take(idx)

    par_filter = filter

    # -- par_groupby (grouping.rs:124-331): one sub-frame per group.  Group label: parts joined with "_", NULL -> "NA" (:158-186);
    #    labels that collide ("a_b" + "c" vs "a" + "b_c", or a literal "NA") share one entry, like the reference's HashMap<String, _>
    def par_groupby(self, group_by_columns: Sequence[str]) -> Dict[str, "OptimizedDataFrame"]:
        if isinstance(group_by_columns, str):
            group_by_columns = [group_by_columns]
        for c in group_by_columns:
            if c not in self._cols:
                raise ColumnNotFound(c)
        if self._rows == 0:
            return {}
        ctx = get_context()
        kcols = [self._cols[c] for c in group_by_columns]
        raws = []
        for k in kcols:                                            # a literal "NULL" string is its own group here (only group_by merges it)
            r = copy.copy(k.raw)
            r.null_alias = -1
            raws.append(r)
        try:
            gr = ctx.groupby_rows(raws)
        except PandrsError as e:
            raise OperationFailed(str(e)) from e
        try:
            parts = []
            for i, k in enumerate(kcols):
                v, isnull = gr.key(i)
                parts.append([("NA" if s == "NULL" and nl else s) for s, nl in zip(_key_strings(k, v, isnull), isnull)])
            labels = ["_".join(p) for p in zip(*parts)]
            off = gr.offsets()
            # one gather per column over the whole permutation (filter_by_indices, data_ops.rs:124-211); a group is a slice of it
            gathered = {name: ctx.gather(self._cols[name].raw, gr.ids_dev(), n=gr.n_rows, idx_dev=True) for name in self._order}
            ids = None
            out: Dict[str, OptimizedDataFrame] = {}
            merged: Dict[str, List[int]] = {}
            for g, lab in enumerate(labels):
                merged.setdefault(lab, []).append(g)
            for lab, gs in merged.items():
                if len(gs) == 1:
                    sel = slice(int(off[gs[0]]), int(off[gs[0] + 1]))
                else:                                              # colliding labels: their rows interleave in ascending row order
                    if ids is None:
                        ids = gr.ids()
                    pos = np.concatenate([np.arange(off[g], off[g + 1]) for g in gs])
                    sel = pos[np.argsort(ids[pos], kind="stable")]
                sub = OptimizedDataFrame()
                for name in self._order:
                    sub.add_column(name, _wrap_gathered(self._cols[name], gathered[name][sel]))
                out[lab] = sub
            return out
        finally:
            gr.close()

    # -- Arrow ingest (src/arrow_integration.rs:75-100, 160-225): validity bitmaps are converted to pandrs null masks and string
    #    arrays are dictionary-encoded on the device; only the distinct strings are interned in the pool on the host
    @staticmethod
    def from_record_batch(batch) -> "OptimizedDataFrame":
        import pyarrow as pa
        ctx = get_context()
        out = OptimizedDataFrame()
        for name, arr in zip(batch.schema.names, batch.columns):
            if isinstance(arr, pa.ChunkedArray):
                arr = arr.combine_chunks()
            n, t, bufs = len(arr), arr.type, arr.buffers()
            validity = None if bufs[0] is None or arr.null_count == 0 else np.frombuffer(bufs[0], np.uint8)
            packed = None
            if pa.types.is_int64(t) or pa.types.is_float64(t):
                if validity is not None:
                    packed, _ = ctx.arrow_validity_to_nulls(validity, n, arr.offset)
                dt = np.int64 if pa.types.is_int64(t) else np.float64
                vals = np.frombuffer(bufs[1], dt)[arr.offset:arr.offset + n] if n else np.empty(0, dt)
                cls = Int64Column if pa.types.is_int64(t) else Float64Column
                col = cls._from_buffers(cls.dtype, vals, packed, n)
                col.values = vals
            elif pa.types.is_boolean(t):
                if arr.offset % 8:
                    arr = pa.concat_arrays([arr])                  # re-aligns the bit-packed values to offset 0
                    bufs = arr.buffers()
                    validity = None if bufs[0] is None or arr.null_count == 0 else np.frombuffer(bufs[0], np.uint8)
                if validity is not None:
                    packed, _ = ctx.arrow_validity_to_nulls(validity, n, arr.offset)
                bits = np.frombuffer(bufs[1], np.uint8)[arr.offset // 8:arr.offset // 8 + (n + 7) // 8].copy() if n else np.empty(0, np.uint8)
                if n % 8:
                    bits[-1] &= (1 << (n % 8)) - 1
                col = BooleanColumn._from_buffers(N.BOOL_BITS, bits, packed, n)
                col.values = np.unpackbits(bits, bitorder="little")[:n].astype(bool)
            elif pa.types.is_string(t) or pa.types.is_large_string(t):
                odt = np.int64 if pa.types.is_large_string(t) else np.int32
                offsets = np.frombuffer(bufs[1], odt)[arr.offset:arr.offset + n + 1] if n else np.zeros(1, odt)
                data = np.frombuffer(bufs[2], np.uint8) if bufs[2] is not None else np.empty(0, np.uint8)
                enc = ctx.dict_encode(offsets, data, validity, arr.offset, n)
                try:
                    first = enc.first_rows()
                    raw = data.tobytes()
                    nullrow = (lambda r: False) if validity is None else (lambda r: not (validity[(arr.offset + r) >> 3] >> ((arr.offset + r) & 7)) & 1)
                    texts = ["" if nullrow(int(r)) else raw[int(offsets[r]):int(offsets[r + 1])].decode("utf-8") for r in first]
                    enc.remap([GLOBAL_STRING_POOL.get_or_insert(s) for s in texts])      # local first-occurrence ids -> pool ids, same order
                    ids = enc.ids()
                    if validity is not None:
                        packed = ctx.to_host(enc.nulls_dev(), (n + 7) // 8, np.uint8)
                finally:
                    enc.close()
                col = StringColumn._from_buffers(N.DICT_U32, ids, packed, n, null_alias=GLOBAL_STRING_POOL.index.get("NULL", -1))
                col.ids = ids
            else:
                raise OperationFailed(f"column '{name}': Arrow type {t} has no typed pandrs column on this path")
            out.add_column(name, col)
        return out

    def _take(self, idx: np.ndarray) -> "OptimizedDataFrame":
        ctx = get_context()
        out = OptimizedDataFrame()
        for name in self._order:
            out.add_column(name, _gathered(ctx, self._cols[name], idx))
        return out

    # -- joins (join.rs:32-47 -> join_impl :76-555)
    def inner_join(self, other: "OptimizedDataFrame", left_on: str, right_on: str) -> "OptimizedDataFrame":
        return self._join(other, left_on, right_on, JoinType.Inner)

    def left_join(self, other: "OptimizedDataFrame", left_on: str, right_on: str) -> "OptimizedDataFrame":
        return self._join(other, left_on, right_on, JoinType.Left)

    def right_join(self, other: "OptimizedDataFrame", left_on: str, right_on: str) -> "OptimizedDataFrame":
        return self._join(other, left_on, right_on, JoinType.Right)

    def outer_join(self, other: "OptimizedDataFrame", left_on: str, right_on: str) -> "OptimizedDataFrame":
        return self._join(other, left_on, right_on, JoinType.Outer)

    def _join(self, other, left_on, right_on, how) -> "OptimizedDataFrame":
        if left_on not in self._cols:
            raise ColumnNotFound(left_on)
        if right_on not in other._cols:
            raise ColumnNotFound(right_on)
        lk, rk = self._cols[left_on], other._cols[right_on]
        if lk.column_type != rk.column_type:                     # join.rs:98-104
            raise ColumnTypeMismatch(f"column '{left_on}': expected {lk.column_type}, found {rk.column_type}")
        ctx = get_context()
        try:
            res = ctx.join_pairs(lk.raw, rk.raw, {JoinType.Inner: N.INNER, JoinType.Left: N.LEFT, JoinType.Right: N.RIGHT, JoinType.Outer: N.OUTER}[how])
        except PandrsError as e:
            raise (ColumnTypeMismatch if e.code == N.ERR_TYPE_MISMATCH else OperationFailed)(str(e)) from e
        li, ri = res.indices()
        res.close()
        out = OptimizedDataFrame()
        if len(li) == 0:                                          # join.rs:230-251: typed empty columns, the key column is omitted
            for name in self._order:
                if name != left_on:
                    out.add_column(name, _empty_like(self._cols[name]))
            for name in other._order:
                if name != right_on:
                    out.add_column(name + "_right" if name in out._cols else name, _empty_like(other._cols[name]))
            return out
        for name in self._order:                                  # join.rs:290-552: [left non-key..., key, right non-key...]
            if name != left_on:
                out.add_column(name, _gathered(ctx, self._cols[name], li))
        out.add_column(left_on, _gathered_key(ctx, lk, rk, li, ri))          # the left key, or the right key on right-only rows (join.rs:395-472)
        for name in other._order:
            if name != right_on:
                out.add_column(name + "_right" if name in out._cols else name, _gathered(ctx, other._cols[name], ri))
        return out


def _gathered_key(ctx: Context, lk: _TypedColumn, rk: _TypedColumn, li: np.ndarray, ri: np.ndarray) -> _TypedColumn:
    a = _gathered(ctx, lk, li)
    if not (li < 0).any():
        return a
    b = _gathered(ctx, rk, ri)
    pick = li < 0
    if lk.dtype == N.DICT_U32:
        return StringColumn(None, _ids=np.where(pick, b.ids, a.ids))
    return type(a)(np.where(pick, b.values, a.values))


def _empty_like(col: _TypedColumn) -> _TypedColumn:
    if col.dtype == N.I64:
        return Int64Column([])
    if col.dtype == N.F64:
        return Float64Column([])
    if col.dtype == N.DICT_U32:
        return StringColumn([])
    return BooleanColumn([])


def _gathered(ctx: Context, col: _TypedColumn, idx: np.ndarray) -> _TypedColumn:
    """pdrs_gather: idx < 0 or a NULL source value -> the type default, no null mask (join.rs:290-552)."""
    return _wrap_gathered(col, ctx.gather(col.raw, idx))


def _wrap_gathered(col: _TypedColumn, vals: np.ndarray) -> _TypedColumn:
    if col.dtype == N.I64:
        return Int64Column(vals)
    if col.dtype == N.F64:
        return Float64Column(vals)
    if col.dtype == N.DICT_U32:
        empty = GLOBAL_STRING_POOL.get_or_insert("")
        ids = np.where(vals == 0xFFFFFFFF, np.uint32(empty), vals)
        return StringColumn(None, _ids=ids)
    return BooleanColumn(vals.astype(bool))


# ---------------------------------------------------------------- GroupBy (group/types.rs:36-67)
AggSpec = Tuple[str, int, str]


class GroupBy:
    def __init__(self, df: OptimizedDataFrame, columns: List[str], as_multi_index: bool = True):
        self.df, self.group_by_columns, self.create_multi_index = df, columns, as_multi_index

    # aggregation.rs:763-871 (serial `aggregate` is the parity target; `par_aggregate` returns the same, aligned rows)
    def aggregate(self, aggregations: Sequence[AggSpec]) -> OptimizedDataFrame:
        return _aggregate(self.df, self.group_by_columns, list(aggregations), self.create_multi_index, None, strict=True)

    def par_aggregate(self, aggregations: Sequence[AggSpec]) -> OptimizedDataFrame:
        return _aggregate(self.df, self.group_by_columns, list(aggregations), self.create_multi_index, None, strict=False)

    # operations.rs:498-521: alias "<col>_<op>"
    def agg(self, aggs: Sequence[Tuple[str, int]]) -> OptimizedDataFrame:
        return self.aggregate([(c, op, f"{c}_{AggregateOp.NAMES[op]}") for c, op in aggs])

    par_agg = agg

    def _one(self, column: str, op: int) -> OptimizedDataFrame:
        return self.aggregate([(column, op, f"{column}_{AggregateOp.NAMES[op]}")])

    def sum(self, column): return self._one(column, AggregateOp.Sum)
    def mean(self, column): return self._one(column, AggregateOp.Mean)
    def min(self, column): return self._one(column, AggregateOp.Min)
    def max(self, column): return self._one(column, AggregateOp.Max)
    def count(self, column): return self._one(column, AggregateOp.Count)
    def std(self, column): return self._one(column, AggregateOp.Std)
    def var(self, column): return self._one(column, AggregateOp.Var)
    def median(self, column): return self._one(column, AggregateOp.Median)
    def first(self, column): return self._one(column, AggregateOp.First)
    def last(self, column): return self._one(column, AggregateOp.Last)
    par_sum, par_mean, par_min, par_max, par_count, par_std, par_var = sum, mean, min, max, count, std, var


def _aggregate(df: OptimizedDataFrame, keys: List[str], aggs: List[AggSpec], multi_index: bool, filter_col: Optional[str],
               strict: bool, allowed_ops=None) -> OptimizedDataFrame:
    for c in keys:
        if c not in df._cols:
            raise ColumnNotFound(c)
    numeric: List[str] = []
    call_pairs: List[Tuple[int, int]] = []
    zero_aggs: List[int] = []
    rows_aggs: List[int] = []                                      # Median / First / Last: from the row lists (aggregation.rs:585-624, 703-742)
    for i, (col, op, alias) in enumerate(aggs):
        if col not in df._cols:
            raise ColumnNotFound(col)
        if op not in (N.SUM, N.MEAN, N.MIN, N.MAX, N.COUNT, N.STD, N.VAR, N.MEDIAN, N.FIRST, N.LAST) or (allowed_ops is not None and op not in allowed_ops):
            raise OperationFailed(f"aggregate op '{AggregateOp.NAMES.get(op, op)}' is outside the accelerated path (no CPU fallback)")
        c = df._cols[col]
        if op in (N.MEDIAN, N.FIRST, N.LAST) and c.dtype in (N.I64, N.F64):
            rows_aggs.append(i)
            call_pairs.append((-1, N.COUNT))
            continue
        if op == N.COUNT:                                          # group size, NULLs included, any column type (aggregation.rs:743)
            call_pairs.append((-1, N.COUNT))
        elif c.dtype not in (N.I64, N.F64):
            if strict:                                             # aggregation.rs:748-752
                raise OperationFailed(f"aggregate op '{AggregateOp.NAMES[op]}' is not supported on a {c.column_type} column")
            zero_aggs.append(i)                                    # par_aggregate: silently 0.0 (aggregation.rs:114-117)
            call_pairs.append((-1, N.COUNT))
        else:
            if col not in numeric:
                numeric.append(col)
            call_pairs.append((numeric.index(col), op))
    ctx = get_context()
    kcols = [df._cols[k] for k in keys]
    fcol = None
    if filter_col is not None:
        f = df.column(filter_col)
        if f.column_type != ColumnType.Boolean:
            raise ColumnTypeMismatch(f"column '{filter_col}': expected Boolean, found {f.column_type}")
        fcol = f.raw
        ctx.set_option("compat_filter_nulls", 1)                   # the reference filters first: NULLs of kept rows become defaults (data_ops.rs:64-71)
        ctx.set_option("compat_empty_string_id", GLOBAL_STRING_POOL.get_or_insert(""))   # ... a NULL string key becomes "" (parallel.rs:224-231)
    try:
        res = ctx.groupby_agg([k.raw for k in kcols], [df._cols[v].raw for v in numeric], call_pairs, filter=fcol)
    except PandrsError as e:
        raise OperationFailed(str(e)) from e
    finally:
        if filter_col is not None:                                 # the context may be shared: leave no option behind
            ctx.set_option("compat_filter_nulls", 0)
            ctx.set_option("compat_empty_string_id", 0xFFFFFFFF)
    try:
        key_strs = []
        for i, k in enumerate(kcols):
            v, isnull = res.key(i)
            key_strs.append(_key_strings(k, v, isnull))
        cols = [res.agg(a) for a in range(len(call_pairs))]
    finally:
        res.close()
    for i in zero_aggs:
        cols[i] = np.zeros_like(cols[i])
    if rows_aggs:
        try:
            gr = ctx.groupby_rows([k.raw for k in kcols])
        except PandrsError as e:
            raise OperationFailed(str(e)) from e
        try:
            gk = []
            for i, k in enumerate(kcols):
                v, isnull = gr.key(i)
                gk.append(_key_strings(k, v, isnull))             # the key text identifies a group in both results ("NULL" = the NULL part)
            where = {kt: g for g, kt in enumerate(zip(*gk))}
            order = np.array([where[kt] for kt in zip(*key_strs)], dtype=np.int64)
            for i in rows_aggs:
                col, op, _ = aggs[i]
                cols[i] = gr.agg(df._cols[col].raw, op)[order]
        finally:
            gr.close()
    out = OptimizedDataFrame()
    if multi_index and len(keys) > 1:                              # aggregation.rs:812-853: keys only in the StringMultiIndex
        out.index = list(zip(*key_strs)) if key_strs else []
        out.index_names = list(keys)
    else:
        for name, ks in zip(keys, key_strs):
            out.add_column(name, StringColumn(ks))
    for (_, _, alias), c in zip(aggs, cols):
        out.add_column(alias, Float64Column(c))
    if multi_index and len(keys) > 1:
        out._rows = len(out.index)
    return out


# ---------------------------------------------------------------- LazyFrame (src/optimized/lazy.rs:60-557) - the operations on this path
class LazyFrame:
    def __init__(self, df: OptimizedDataFrame):
        self.df = df
        self._ops: List[tuple] = []

    @staticmethod
    def new(df):
        return LazyFrame(df)

    def filter(self, column: str) -> "LazyFrame":
        self._ops.append(("filter", column))
        return self

    def aggregate(self, group_by: Sequence[str], aggregations: Sequence[AggSpec]) -> "LazyFrame":
        self._ops.append(("aggregate", list(group_by), list(aggregations)))
        return self

    def join(self, right: OptimizedDataFrame, left_on: str, right_on: str, join_type: int = JoinType.Inner) -> "LazyFrame":
        self._ops.append(("join", right, left_on, right_on, join_type))
        return self

    def execute(self) -> OptimizedDataFrame:
        df = self.df
        ops = list(self._ops)
        i = 0
        while i < len(ops):
            op = ops[i]
            if op[0] == "filter" and i + 1 < len(ops) and ops[i + 1][0] == "aggregate":
                # filter -> aggregate is fused into one kernel pass: the filter is the row mask of pdrs_groupby_agg
                _, keys, aggs = ops[i + 1]
                df = _aggregate(df, keys, aggs, False, op[1], strict=True, allowed_ops=(N.SUM, N.MEAN, N.MIN, N.MAX, N.COUNT))
                i += 2
                continue
            if op[0] == "filter":
                df = df.filter(op[1])
            elif op[0] == "aggregate":                             # lazy.rs:267-383: Sum / Mean / Min / Max / Count only
                df = _aggregate(df, op[1], op[2], False, None, strict=True, allowed_ops=(N.SUM, N.MEAN, N.MIN, N.MAX, N.COUNT))
            elif op[0] == "join":
                df = df._join(op[1], op[2], op[3], op[4])
            i += 1
        return df
