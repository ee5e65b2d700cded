"""Chunked (out-of-core) groupby over HOST columns: pdrs_groupby_agg / pdrs_groupby_agg_where cut large host inputs into chunks
that travel through the staging engine while the previous chunk is aggregated, and merge the chunks' states (SURVEY.md 8(f) row 4,
reference src/large/mod.rs).  Same parity bar as the one-shot path: the oracle on the whole input."""
import numpy as np
import pytest

import pandrs_b200 as pb
from _util import Spec, compare_groupby

pytestmark = pytest.mark.gpu

ALL6 = [pb.SUM, pb.MEAN, pb.MIN, pb.MAX, pb.COUNT, pb.STD]


@pytest.fixture
def sctx(ctx):
    # every host input streams: 65 536-row chunks
    ctx.set_option("stream_rows", 1)
    ctx.set_option("stream_chunk_rows", 65536)
    ctx.set_option("compat_empty_string_id", 0xFFFFFFFF)       # (the session context is shared: start from the default)
    yield ctx
    ctx.set_option("stream_rows", 1 << 25)
    ctx.set_option("stream_chunk_rows", 0)
    ctx.set_option("stream_compact_rows", 1 << 22)
    ctx.set_option("stage_threads", 0)
    ctx.set_option("compat_filter_nulls", 0)


@pytest.mark.parametrize("threads", [0, 3, -1])
def test_chunked_matches_oracle(sctx, oracle, threads):
    # 4 full chunks + a ragged one; NULL keys and values; all six aggregates of an f64 and an i64 column
    sctx.set_option("stage_threads", threads)
    rng = np.random.default_rng(21)
    n = 4 * 65536 + 12_345
    k = Spec(pb.I64, rng.integers(-500, 500, n), nulls=rng.random(n) < 0.01)
    v = Spec(pb.F64, rng.normal(3.0, 2.0, n), nulls=rng.random(n) < 0.05)
    w = Spec(pb.I64, rng.integers(-10**12, 10**12, n), nulls=rng.random(n) < 0.02)
    got = compare_groupby(pb, oracle, sctx, [k], [v, w], [(0, op) for op in ALL6] + [(1, op) for op in ALL6])
    assert len(got) == 1001
    # the value column no aggregate reads does not travel; Count on a string column is fine
    got = compare_groupby(pb, oracle, sctx, [k], [Spec(pb.DICT_U32, rng.integers(0, 5, n)), v], [(0, pb.COUNT), (1, pb.SUM)])
    assert len(got) == 1001


def test_chunked_multi_key_filter_and_predicate(sctx, oracle):
    rng = np.random.default_rng(22)
    n = 3 * 65536 + 77
    pool = ["s%d" % i for i in range(40)] + ["NULL"]
    k0 = Spec(pb.DICT_U32, rng.integers(0, 41, n), nulls=rng.random(n) < 0.03, pool=pool, null_alias=40)
    k1 = Spec(pb.I32, rng.integers(-7, 7, n).astype(np.int32), nulls=rng.random(n) < 0.03)
    k2 = Spec(pb.BOOL_BITS, rng.random(n) < 0.5)
    v = Spec(pb.F64, rng.normal(0.0, 1.0, n) * 1e6, nulls=rng.random(n) < 0.1)
    f = Spec(pb.BOOL_BITS, rng.random(n) < 0.8, nulls=rng.random(n) < 0.05)
    aggs = [(0, op) for op in ALL6 + [pb.VAR]]
    compare_groupby(pb, oracle, sctx, [k0, k1, k2], [v], aggs, filter_spec=f)
    d = Spec(pb.I64, rng.integers(0, 1000, n), nulls=rng.random(n) < 0.02)
    compare_groupby(pb, oracle, sctx, [k0, k1], [v], aggs, filter_spec=f, pred=(d, pb.CMP_LE, 600))
    sctx.set_option("compat_filter_nulls", 1)
    compare_groupby(pb, oracle, sctx, [k0, k1], [v], aggs, filter_spec=f, compat_nulls=True)


def test_chunked_high_cardinality_compaction_and_short_masks(sctx, oracle):
    # many groups per chunk: the appended state rows are compacted between chunks; the null mask of the value column is SHORT
    # (covers only the first 100 000 rows: missing bytes mean "not NULL", core/column.rs:163-177)
    sctx.set_option("stream_compact_rows", 1000)
    rng = np.random.default_rng(23)
    n = 5 * 65536 + 1
    kv = rng.integers(0, 30_000, n) * 7919
    vv = rng.normal(100.0, 5.0, n)
    nulls = np.zeros(n, dtype=bool)
    nulls[:100_000] = rng.random(100_000) < 0.1
    short = pb.Column(pb.F64, vv, pb.pack_bits(nulls[:100_000]))
    assert short.c().null_len == 12_500
    res = sctx.groupby_agg([pb.Column.int64(kv)], [short], [(0, op) for op in ALL6])
    ref = sctx.groupby_agg([sctx.upload(pb.Column.int64(kv))], [sctx.upload(pb.Column(pb.F64, vv, pb.pack_bits(nulls)))], [(0, op) for op in ALL6])
    try:
        assert res.n_groups == ref.n_groups == len(np.unique(kv))
        a, b = np.argsort(res.key(0)[0]), np.argsort(ref.key(0)[0])
        assert np.array_equal(res.group_rows()[a], ref.group_rows()[b]) and np.array_equal(res.valid_n(0)[a], ref.valid_n(0)[b])
        for i, op in enumerate(ALL6):
            x, y = res.agg(i)[a], ref.agg(i)[b]
            if op in (pb.MIN, pb.MAX, pb.COUNT):
                assert np.array_equal(x, y)
            else:
                assert np.allclose(x, y, rtol=1e-12, atol=0)
    finally:
        res.close(); ref.close()
    got = compare_groupby(pb, oracle, sctx, [Spec(pb.I64, kv)], [Spec(pb.F64, vv, nulls=nulls)], [(0, op) for op in ALL6])
    assert len(got) == len(np.unique(kv))


def test_chunked_pinned_source_and_empty_chunks(sctx, oracle):
    # a cudaHostAlloc'ed source takes the direct DMA path of the staging engine; a filter that rejects whole chunks
    n = 3 * 65536
    rng = np.random.default_rng(24)
    kv = rng.integers(0, 50, n)
    vv = rng.random(n)
    pk, pv = sctx.host_alloc(8 * n), sctx.host_alloc(8 * n)
    try:
        import ctypes
        ka = np.ctypeslib.as_array((ctypes.c_int64 * n).from_address(pk))
        va = np.ctypeslib.as_array((ctypes.c_double * n).from_address(pv))
        ka[:] = kv
        va[:] = vv
        keep = np.zeros(n, dtype=bool)
        keep[65536:2 * 65536] = True
        res = sctx.groupby_agg([pb.Column(pb.I64, ka)], [pb.Column(pb.F64, va)], [(0, pb.SUM), (0, pb.COUNT)], filter=pb.Column.boolean(keep))
        try:
            o = np.argsort(res.key(0)[0])
            sl = slice(65536, 2 * 65536)
            want = np.array([vv[sl][kv[sl] == g].sum() for g in range(50)])
            assert res.n_groups == 50 and np.allclose(res.agg(0)[o], want, rtol=1e-12)
            assert np.array_equal(res.agg(1)[o], np.bincount(kv[sl], minlength=50).astype(np.float64))
        finally:
            res.close()
        del ka, va
    finally:
        sctx.host_free(pk)
        sctx.host_free(pv)


def test_large_results_and_columns_through_the_staging_engine(ctx, oracle):
    # host key columns >= 64 MB are uploaded by the staging threads and index pairs >= 64 MB come back through them (pageable numpy
    # memory on both sides): same pairs as the oracle's join of the same keys
    n = 9_000_000
    rng = np.random.default_rng(31)
    right = rng.permutation(n).astype(np.int64) * 7 + 3
    left = (rng.integers(0, 2 * n, n) * 7 + 3).astype(np.int64)
    j = ctx.join_pairs(pb.Column.int64(left), pb.Column.int64(right), pb.LEFT)
    try:
        li, ri = j.indices()
    finally:
        j.close()
    assert len(li) == n and np.array_equal(np.sort(li), np.arange(n))
    where = np.full(2 * n, -1, np.int64)
    where[(right - 3) // 7] = np.arange(n)
    assert np.array_equal(ri, where[(left[li] - 3) // 7])
