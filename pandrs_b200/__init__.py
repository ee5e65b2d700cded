"""pandrs_b200 — the B200 (sm_100a) execution path for pandrs's groupby-aggregate and inner/left
hash-join hot paths, behind a C ABI (include/pandrs_b200.h).

This package is only the host-side stand-in for the Rust caller: ctypes bindings (`_native`), a thin
object layer (`core`) and a mirror of the OptimizedDataFrame / LazyFrame API for this path (`frame`).
All compute happens in pandrs_b200/lib/libpandrs_b200.so; there is no CPU fallback.
"""
from ._native import (BOOL_BITS, CMP_EQ, CMP_GE, CMP_GT, CMP_LE, CMP_LT, CMP_NE, COUNT, DICT_U32, F64, GB_AUTO, GB_DENSE, GB_FEW, GB_GLOBAL, GB_PARTITIONED, GB_SHARED, GB_TILESORT, I32, I64, INNER, LEFT, MAX, MEAN, OUTER, RIGHT,
                      MEM_DEVICE, MEM_HOST, MIN, STD, SUM, VAR, MEDIAN, FIRST, LAST, build, lib)
from .core import Column, Comm, Context, DictEncoded, GroupByResult, GroupRows, JoinResult, PandrsError, XJoin, pack_bits, torch_broadcast_id

__all__ = ["Column", "Comm", "torch_broadcast_id", "Context", "GroupByResult", "GroupRows", "DictEncoded", "JoinResult", "PandrsError", "XJoin", "pack_bits", "build", "lib",
           "I64", "F64", "DICT_U32", "BOOL_BITS", "I32", "SUM", "MEAN", "MIN", "MAX", "COUNT", "STD", "VAR", "MEDIAN", "FIRST", "LAST",
           "INNER", "LEFT", "RIGHT", "OUTER", "MEM_HOST", "MEM_DEVICE", "GB_AUTO", "GB_SHARED", "GB_GLOBAL", "GB_DENSE", "GB_TILESORT", "GB_PARTITIONED", "GB_FEW",
           "CMP_LT", "CMP_LE", "CMP_GT", "CMP_GE", "CMP_EQ", "CMP_NE"]
