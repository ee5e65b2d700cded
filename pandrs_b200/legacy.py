"""Host-side mirror of the LEGACY pandrs API the north star names - `DataFrame::groupby(...)` over string-materialised Series:

  DataFrame / Series (string values)                     src/dataframe/base.rs, src/series/base.rs
  GroupByExt::groupby / groupby_single                   src/dataframe/groupby.rs:606-623
  DataFrameGroupBy::new (row -> group by the key STRINGS) src/dataframe/groupby.rs:196-228
  DataFrameGroupBy::agg / sum / mean / ... / size        src/dataframe/groupby.rs:236-382
  calculate_aggregation (values parsed as f64,           src/dataframe/groupby.rs:443-532
      unparseable cells skipped, Count = parseable cells)
  optimize_dataframe (type inference by string parsing)  src/optimized/convert.rs:13-110, 253-255

The reference walks `HashMap<Vec<String>, Vec<usize>>` and re-parses the value column once per group and aggregate (O(groups x rows)).
Here the frame is lowered ONCE to typed columns - group keys to dictionary ids (equal string <=> equal id), value cells to f64 with
a null mask over the unparseable ones - and a single pdrs_groupby_agg call computes every aggregate; the legacy `Count` (parseable
cells, not group size) is the `valid_n` array the shim returns beside `group_rows`.  Results go back to strings
exactly like `agg_result.to_string()` (groupby.rs:291).  Median runs on the row lists (pdrs_group_rows_agg).  First / Last /
Nunique / Custom are outside the accelerated path (no CPU fallback): OperationFailed.
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _native as N
from .core import Column as RawColumn
from .core import PandrsError
from .frame import (GLOBAL_STRING_POOL, BooleanColumn, ColumnNotFound, Float64Column, Int64Column, OperationFailed, OptimizedDataFrame, StringColumn,
                    _f64_display, get_context)

__all__ = ["AggFunc", "NamedAgg", "Series", "DataFrame", "DataFrameGroupBy", "optimize_dataframe"]


class AggFunc:
    """src/dataframe/groupby.rs:17-45"""
    Sum, Mean, Min, Max, Count, Std, Var, Median, First, Last, Nunique, Custom = range(12)
    NAMES = ["sum", "mean", "min", "max", "count", "std", "var", "median", "first", "last", "nunique", "custom"]


class NamedAgg:
    """groupby.rs:52-92"""

    def __init__(self, column: str, func: int, alias: str):
        self.column, self.func, self.alias = column, func, alias


class Series:
    """A named vector of cells held as strings (what DataFrame::get_column_string_values returns)."""

    def __init__(self, values: Sequence, name: Optional[str] = None):
        self.values = [v if isinstance(v, str) else _cell_to_string(v) for v in values]
        self.name = name

    def __len__(self):
        return len(self.values)


def _cell_to_string(v) -> str:
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    if isinstance(v, (float, np.floating)):
        return _f64_display(float(v))
    return str(v)


# Rust's `str::parse::<f64>`: optional sign, decimal digits with optional fraction / exponent, or inf / infinity / nan (any case);
# no surrounding whitespace, no underscores (Python's float() accepts both)
_F64 = re.compile(r"^[+-]?((\d+\.?\d*|\.\d+)([eE][+-]?\d+)?|inf|infinity|nan)$", re.IGNORECASE)
_I64 = re.compile(r"^[+-]?\d+$")


def _parse_f64(s: str) -> Optional[float]:
    return float(s) if _F64.match(s) else None


def _parse_i64(s: str) -> Optional[int]:
    if not _I64.match(s):
        return None
    v = int(s)
    return v if -(1 << 63) <= v < (1 << 63) else None


class DataFrame:
    def __init__(self):
        self._cols: Dict[str, Series] = {}
        self._order: List[str] = []

    @staticmethod
    def new():
        return DataFrame()

    def add_column(self, name: str, series: Series):
        if name in self._cols:
            raise ValueError(f"duplicate column name '{name}'")
        if self._order and len(series) != self.row_count():
            raise ValueError(f"inconsistent row count: expected {self.row_count()}, column '{name}' has {len(series)}")
        self._cols[name] = series if isinstance(series, Series) else Series(series, name)
        self._order.append(name)
        return self

    def row_count(self) -> int:
        return len(self._cols[self._order[0]]) if self._order else 0

    def column_names(self) -> List[str]:
        return list(self._order)

    def contains_column(self, name: str) -> bool:
        return name in self._cols

    def get_column_string_values(self, name: str) -> List[str]:
        if name not in self._cols:
            raise ColumnNotFound(name)
        return list(self._cols[name].values)

    def get_column_numeric_values(self, name: str) -> List[float]:
        return [float(v) for v in self.get_column_string_values(name)]

    # groupby.rs:606-623
    def groupby(self, columns: Sequence[str]) -> "DataFrameGroupBy":
        return DataFrameGroupBy(self, [columns] if isinstance(columns, str) else list(columns))

    def groupby_single(self, column: str) -> "DataFrameGroupBy":
        return DataFrameGroupBy(self, [column])


def optimize_dataframe(df: DataFrame) -> OptimizedDataFrame:
    """optimized/convert.rs:13-110: per column, the first type every cell parses as - Int64 ("" -> 0), Float64 ("" -> 0.0), Boolean
    (true / false / 1 / 0, any case; "" -> false) - else String.  No null masks (empty cells become defaults)."""
    out = OptimizedDataFrame()
    for name in df.column_names():
        vals = df._cols[name].values
        if all(s == "" or _parse_i64(s) is not None for s in vals):
            out.add_column(name, Int64Column([(_parse_i64(s) or 0) if s else 0 for s in vals]))
        elif all(s == "" or _parse_f64(s) is not None for s in vals):
            out.add_column(name, Float64Column([_parse_f64(s) if s else 0.0 for s in vals]))
        elif all(s.lower() in ("", "true", "false", "1", "0") for s in vals):
            out.add_column(name, BooleanColumn([s.lower() in ("true", "1") for s in vals]))
        else:
            out.add_column(name, StringColumn(vals))
    return out


class DataFrameGroupBy:
    """groupby.rs:188-228: groups are keyed by the STRINGS of the grouping columns (a literal "NULL" is just a string here)."""

    _OPS = {AggFunc.Sum: N.SUM, AggFunc.Mean: N.MEAN, AggFunc.Min: N.MIN, AggFunc.Max: N.MAX, AggFunc.Std: N.STD, AggFunc.Var: N.VAR}

    def __init__(self, df: DataFrame, group_by_columns: List[str]):
        for c in group_by_columns:
            if not df.contains_column(c):
                raise ColumnNotFound(c)
        self.df, self.group_by_columns = df, group_by_columns
        self._keys = []
        for c in group_by_columns:             # dictionary ids through the pool: equal string <=> equal id; no NULLs, no "NULL" alias
            ids = np.fromiter((GLOBAL_STRING_POOL.get_or_insert(s) for s in df._cols[c].values), dtype=np.uint32, count=df.row_count())
            self._keys.append(RawColumn.dict_ids(ids, None, null_alias=-1))
        self._values: Dict[str, RawColumn] = {}

    def _value_column(self, name: str) -> RawColumn:
        """The cells of a column as f64 + a null mask over the cells `parse::<f64>()` rejects (groupby.rs:452-462)."""
        if name not in self._values:
            cells = self.df.get_column_string_values(name)
            parsed = [_parse_f64(s) for s in cells]
            nulls = np.fromiter((p is None for p in parsed), dtype=bool, count=len(cells))
            vals = np.fromiter((0.0 if p is None else p for p in parsed), dtype=np.float64, count=len(cells))
            self._values[name] = RawColumn.float64(vals, nulls if nulls.any() else None)
        return self._values[name]

    def _run(self, value_names: List[str], pairs):
        ctx = get_context()
        try:
            return ctx.groupby_agg(self._keys, [self._value_column(v) for v in value_names], pairs)
        except PandrsError as e:
            raise OperationFailed(str(e)) from e

    def _group_strings(self, res) -> List[List[str]]:
        return [[GLOBAL_STRING_POOL.get(int(i)) for i in res.key(k)[0]] for k in range(len(self.group_by_columns))]

    def ngroups(self) -> int:
        res = self._run([], [])
        try:
            return res.n_groups
        finally:
            res.close()

    def size(self) -> DataFrame:
        """groupby.rs:236-255: columns "group" (key parts joined with "_") and "size"."""
        res = self._run([], [])
        try:
            parts, rows = self._group_strings(res), res.group_rows()
        finally:
            res.close()
        out = DataFrame()
        out.add_column("group", Series(["_".join(p) for p in zip(*parts)], "group"))
        out.add_column("size", Series([str(int(r)) for r in rows], "size"))
        return out

    def agg(self, named_aggs: Sequence[NamedAgg]) -> DataFrame:
        """groupby.rs:258-300 + calculate_aggregation :443-532.  Every aggregate of every column in ONE pdrs_groupby_agg call."""
        if not named_aggs:
            raise ValueError("At least one aggregation must be specified")
        for a in named_aggs:
            if not self.df.contains_column(a.column):
                raise ColumnNotFound(a.column)
            if a.func not in self._OPS and a.func not in (AggFunc.Count, AggFunc.Median):
                raise OperationFailed(f"aggregation '{AggFunc.NAMES[a.func]}' of the legacy groupby is outside the accelerated path (no CPU fallback)")
        names: List[str] = []
        for a in named_aggs:
            if a.column not in names:
                names.append(a.column)
        # a Sum per value column makes the shim keep its valid count (= the legacy Count) even when no other aggregate reads it
        pairs = [(names.index(a.column), self._OPS.get(a.func, N.SUM)) for a in named_aggs] + [(v, N.SUM) for v in range(len(names))]
        res = self._run(names, pairs)
        try:
            parts = self._group_strings(res)
            valid = {v: res.valid_n(i) for i, v in enumerate(names)}
            cols = [res.agg(i) for i in range(len(named_aggs))]
            kv = [res.key(k)[0] for k in range(len(self.group_by_columns))]
        finally:
            res.close()
        med = None
        if any(a.func == AggFunc.Median for a in named_aggs):
            ctx = get_context()
            gr = ctx.groupby_rows(self._keys)
            try:
                gk = [gr.key(k)[0] for k in range(len(self.group_by_columns))]
                where = {t: g for g, t in enumerate(zip(*[x.tolist() for x in gk]))}
                order = np.array([where[t] for t in zip(*[x.tolist() for x in kv])], dtype=np.int64)
                med = {a.column: gr.agg(self._value_column(a.column), N.MEDIAN)[order] for a in named_aggs if a.func == AggFunc.Median}
            finally:
                gr.close()
        out = DataFrame()
        for name, p in zip(self.group_by_columns, parts):
            out.add_column(name, Series(p, name))
        for a, c in zip(named_aggs, cols):
            n = valid[a.column]
            if a.func == AggFunc.Count:
                c = n.astype(np.float64)
            elif a.func == AggFunc.Median:
                c = med[a.column]
            elif a.func in (AggFunc.Std, AggFunc.Var):
                c = np.where(n <= 1, 0.0, c)                       # groupby.rs:478-499
            c = np.where(n == 0, 0.0, c)                           # groupby.rs:464-466: no parseable cell -> 0.0
            out.add_column(a.alias, Series([_f64_display(float(x)) for x in c], a.alias))
        return out

    def _one(self, column: str, func: int) -> DataFrame:
        return self.agg([NamedAgg(column, func, f"{column}_{AggFunc.NAMES[func]}")])

    def sum(self, column): return self._one(column, AggFunc.Sum)
    def mean(self, column): return self._one(column, AggFunc.Mean)
    def min(self, column): return self._one(column, AggFunc.Min)
    def max(self, column): return self._one(column, AggFunc.Max)
    def count(self, column): return self._one(column, AggFunc.Count)
    def std(self, column): return self._one(column, AggFunc.Std)
    def var(self, column): return self._one(column, AggFunc.Var)
    def median(self, column): return self._one(column, AggFunc.Median)
