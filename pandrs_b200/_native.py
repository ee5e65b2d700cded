"""ctypes binding of libpandrs_b200.so (include/pandrs_b200.h).

The library is the product: there is no Python, numpy or oracle fallback behind these calls.  If the
shared object is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpandrs_b200.so")
CSRC = os.path.join(_HERE, "csrc")

# pdrs_dtype / pdrs_agg_op / pdrs_join_type / pdrs_mem / pdrs_groupby_algo (include/pandrs_b200.h)
I64, F64, DICT_U32, BOOL_BITS, I32 = 0, 1, 2, 3, 4
SUM, MEAN, MIN, MAX, COUNT, STD, VAR, MEDIAN, FIRST, LAST = range(10)
INNER, LEFT, RIGHT, OUTER = 0, 1, 2, 3
MEM_HOST, MEM_DEVICE = 0, 1
GB_AUTO, GB_SHARED, GB_GLOBAL, GB_DENSE, GB_TILESORT, GB_PARTITIONED, GB_FEW = 0, 1, 2, 3, 4, 5, 6
CMP_LT, CMP_LE, CMP_GT, CMP_GE, CMP_EQ, CMP_NE = 0, 1, 2, 3, 4, 5

OK, ERR_BAD_ARG, ERR_TYPE_MISMATCH, ERR_OOM, ERR_CUDA, ERR_NCCL, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6


class PdrsCol(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("mem", C.c_int32), ("data", C.c_void_p), ("null_bits", C.c_void_p),
                ("null_len", C.c_int64), ("len", C.c_int64), ("null_alias", C.c_int64)]


class PdrsAgg(C.Structure):
    _fields_ = [("value_col", C.c_int32), ("op", C.c_int32)]


class PdrsPred(C.Structure):
    _fields_ = [("col", PdrsCol), ("op", C.c_int32), ("reserved", C.c_int32), ("ival", C.c_int64), ("fval", C.c_double)]


class PdrsOptions(C.Structure):
    _fields_ = [("device", C.c_int32), ("groupby_algo", C.c_int32), ("groups_hint", C.c_int64), ("stream", C.c_void_p),
                ("compat_filter_nulls", C.c_int32), ("reserved0", C.c_int32), ("reserved", C.c_int64 * 4)]


class PdrsStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("groupby_algo_used", C.c_int32), ("retries", C.c_int32),
                ("est_groups", C.c_int64), ("table_slots", C.c_int64), ("spilled_rows", C.c_int64),
                ("main_kernel_ms", C.c_float), ("total_ms", C.c_float)]


_P = C.POINTER
_vp, _i32, _i64, _u64, _u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_uint32

# name -> (restype, argtypes): exactly the entry points include/pandrs_b200.h declares
SIGNATURES = {
    "pdrs_abi_version": (_i32, []),
    "pdrs_ctx_create": (_i32, [_P(PdrsOptions), _P(_vp)]),
    "pdrs_ctx_destroy": (None, [_vp]),
    "pdrs_last_error": (C.c_char_p, [_vp]),
    "pdrs_sync": (_i32, [_vp]),
    "pdrs_get_stats": (_i32, [_vp, _P(PdrsStats)]),
    "pdrs_set_option": (_i32, [_vp, C.c_char_p, _i64]),
    "pdrs_col_upload": (_i32, [_vp, _P(PdrsCol), _P(PdrsCol)]),
    "pdrs_col_free": (_i32, [_vp, _P(PdrsCol)]),
    "pdrs_host_alloc": (_i32, [_vp, _i64, _P(_vp)]),
    "pdrs_host_free": (_i32, [_vp, _vp]),
    "pdrs_groupby_agg": (_i32, [_vp, _P(PdrsCol), _i32, _P(PdrsCol), _i32, _P(PdrsAgg), _i32, _P(PdrsCol), _P(_vp)]),
    "pdrs_groupby_agg_where": (_i32, [_vp, _P(PdrsCol), _i32, _P(PdrsCol), _i32, _P(PdrsAgg), _i32, _P(PdrsCol), _P(PdrsPred), _P(_vp)]),
    "pdrs_groupby_n_groups": (_i64, [_vp]),
    "pdrs_groupby_key": (_i32, [_vp, _i32, _vp, _vp]),
    "pdrs_groupby_agg_values": (_i32, [_vp, _i32, _vp]),
    "pdrs_groupby_group_rows": (_i32, [_vp, _vp]),
    "pdrs_groupby_valid_n": (_i32, [_vp, _i32, _vp]),
    "pdrs_groupby_key_dev": (_vp, [_vp, _i32]),
    "pdrs_groupby_key_null_dev": (_vp, [_vp, _i32]),
    "pdrs_groupby_agg_dev": (_vp, [_vp, _i32]),
    "pdrs_groupby_group_rows_dev": (_vp, [_vp]),
    "pdrs_groupby_result_free": (None, [_vp]),
    "pdrs_groupby_partial": (_i32, [_vp, _P(PdrsCol), _i32, _P(PdrsCol), _i32, _P(PdrsCol), _i32, _P(_vp)]),
    "pdrs_groupby_states_dev": (_vp, [_vp, _i32]),
    "pdrs_groupby_merge": (_i32, [_vp, _P(PdrsCol), _i32, _P(_vp), _P(_i32), _i32, _i64, _P(PdrsAgg), _i32, _P(_vp)]),
    "pdrs_hash_partition": (_i32, [_vp, _P(PdrsCol), _i32, _i32, _vp, _P(_i64)]),
    "pdrs_groupby_rows": (_i32, [_vp, _P(PdrsCol), _i32, _P(_vp)]),
    "pdrs_group_rows_n_groups": (_i64, [_vp]),
    "pdrs_group_rows_n_rows": (_i64, [_vp]),
    "pdrs_group_rows_key": (_i32, [_vp, _i32, _vp, _vp]),
    "pdrs_group_rows_offsets": (_i32, [_vp, _vp]),
    "pdrs_group_rows_ids": (_i32, [_vp, _vp]),
    "pdrs_group_rows_offsets_dev": (_vp, [_vp]),
    "pdrs_group_rows_ids_dev": (_vp, [_vp]),
    "pdrs_group_rows_agg": (_i32, [_vp, _P(PdrsCol), _i32, _vp]),
    "pdrs_group_rows_free": (None, [_vp]),
    "pdrs_arrow_validity_to_nulls": (_i32, [_vp, _vp, _i32, _i64, _i64, _vp, _i32, _P(_i64)]),
    "pdrs_dict_encode": (_i32, [_vp, _vp, _i32, _vp, _i64, _vp, _i64, _i64, _i32, _P(_vp)]),
    "pdrs_dict_n_unique": (_i64, [_vp]),
    "pdrs_dict_ids": (_i32, [_vp, _vp]),
    "pdrs_dict_ids_dev": (_vp, [_vp]),
    "pdrs_dict_first_rows": (_i32, [_vp, _vp]),
    "pdrs_dict_nulls_dev": (_vp, [_vp]),
    "pdrs_dict_remap": (_i32, [_vp, _vp]),
    "pdrs_dict_free": (None, [_vp]),
    "pdrs_join_pairs": (_i32, [_vp, _P(PdrsCol), _P(PdrsCol), _i32, _P(_vp)]),
    "pdrs_join_gather": (_i32, [_vp, _P(PdrsCol), _P(PdrsCol), _i32, _P(PdrsCol), _i32, _P(_vp)]),
    "pdrs_join_right_col": (_i32, [_vp, _i32, _vp]),
    "pdrs_join_right_col_dev": (_vp, [_vp, _i32]),
    "pdrs_join_len": (_i64, [_vp]),
    "pdrs_join_indices": (_i32, [_vp, _vp, _vp]),
    "pdrs_join_left_dev": (_vp, [_vp]),
    "pdrs_join_right_dev": (_vp, [_vp]),
    "pdrs_join_result_free": (None, [_vp]),
    "pdrs_xjoin_create": (_i32, [_vp, _i32, _i32, _i64, _i64, _i64, _P(_vp)]),
    "pdrs_xjoin_bytes": (_i64, [_vp]),
    "pdrs_xjoin_base": (_vp, [_vp]),
    "pdrs_xjoin_ipc_handle": (_i32, [_vp, _vp]),
    "pdrs_xjoin_attach_ipc": (_i32, [_vp, _vp]),
    "pdrs_xjoin_attach_ptrs": (_i32, [_vp, _P(_vp)]),
    "pdrs_xjoin_shuffle": (_i32, [_vp, _P(PdrsCol), _P(PdrsCol), _i64]),
    "pdrs_xjoin_local": (_i32, [_vp, _i32, _P(_i64), _P(_vp)]),
    "pdrs_xjoin_destroy": (None, [_vp]),
    "pdrs_comm_unique_id": (_i32, [_vp]),
    "pdrs_comm_init": (_i32, [_vp, _i32, _i32, _vp, _P(_vp)]),
    "pdrs_comm_rank": (_i32, [_vp]),
    "pdrs_comm_size": (_i32, [_vp]),
    "pdrs_comm_set_option": (_i32, [_vp, C.c_char_p, _i64]),
    "pdrs_comm_barrier": (_i32, [_vp]),
    "pdrs_comm_last_exchange": (_i32, [_vp, _P(C.c_float), _P(_i64)]),
    "pdrs_comm_destroy": (None, [_vp]),
    "pdrs_groupby_agg_dist": (_i32, [_vp, _P(PdrsCol), _i32, _P(PdrsCol), _i32, _P(PdrsAgg), _i32, _P(PdrsCol), _P(PdrsPred), _i32, _P(_vp)]),
    "pdrs_join_pairs_dist": (_i32, [_vp, _P(PdrsCol), _P(PdrsCol), _i32, _i64, _i64, _i64, _i64, _i64, _P(_vp)]),
    "pdrs_gather": (_i32, [_vp, _P(PdrsCol), _vp, _i32, _i64, _vp, _i32]),
    "pdrs_filter_indices": (_i32, [_vp, _P(PdrsCol), _vp, _P(_i64)]),
    "pdrs_synth_keys": (_i32, [_vp, _vp, _i64, _i64, _u64, _u64, _i32]),
    "pdrs_synth_vals": (_i32, [_vp, _vp, _i64, _i64, _u64]),
    "pdrs_synth_nulls": (_i32, [_vp, _vp, _i64, _i64, _u64, _u32]),
    "pdrs_synth_join_keys": (_i32, [_vp, _vp, _i64, _i64, _u64, _u64, _i32]),
    "pdrs_dev_alloc": (_i32, [_vp, _i64, _P(_vp)]),
    "pdrs_dev_free": (_i32, [_vp, _vp]),
    "pdrs_memcpy": (_i32, [_vp, _vp, _vp, _i64, _i32]),
    "pdrs_flush_l2": (_i32, [_vp]),
    "pdrs_timer_begin": (_i32, [_vp]),
    "pdrs_timer_end": (_i32, [_vp, _P(C.c_float)]),
}


def build(force: bool = False, jobs: int = 8) -> str:
    """Compile libpandrs_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, f"-j{jobs}"] + (["-B"] if force else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libpandrs_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(pandrs_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)   # AttributeError here = the .so does not match the header
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib
