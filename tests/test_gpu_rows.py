"""Row lists per group (par_groupby), the order-dependent aggregates (Median / First / Last) and the columnar ingest in front of
the path (Arrow validity -> null masks, device-side dictionary encoding) - all through the C ABI, against the CPU oracle / numpy.
Reference: grouping.rs:124-331, aggregation.rs:585-624 + 703-742, arrow_integration.rs:160-225, string_pool.rs:28-52."""
import numpy as np
import pytest

import pandrs_b200 as pb
from _util import Spec, key_string

pytestmark = pytest.mark.gpu


def _labels(res, specs):
    cols = [res.key(k) for k in range(len(specs))]
    out = []
    for g in range(res.n_groups):
        parts = []
        for k, s in enumerate(specs):
            t = key_string(s.dtype, cols[k][0][g], cols[k][1][g], s.pool, pb)
            parts.append("NA" if cols[k][1][g] else t)
        out.append("_".join(parts))
    return out


def _check_rows(ctx, oracle, specs):
    want = oracle.par_groupby([s.cpu(oracle) for s in specs])
    res = ctx.groupby_rows([s.gpu(pb) for s in specs])
    try:
        labels, off, ids = _labels(res, specs), res.offsets(), res.ids()
        assert res.n_rows == len(specs[0]) and off[0] == 0 and off[-1] == res.n_rows
        got = {}
        for g, lab in enumerate(labels):
            rows = ids[off[g]:off[g + 1]]
            assert np.all(np.diff(rows) > 0), "rows of a group must ascend"
            got[lab] = np.sort(np.concatenate([got[lab], rows])) if lab in got else rows     # colliding labels share an entry
        assert set(got) == set(want)
        for lab, rows in want.items():
            assert np.array_equal(got[lab], rows), lab
    finally:
        res.close()
    return len(want)


@pytest.mark.parametrize("n,card", [(1, 1), (33, 1), (1000, 5), (8193, 300), (100_003, 300), (70_001, 40_000), (300_000, 70_000)])
def test_row_lists_match_par_groupby(ctx, oracle, n, card):
    rng = np.random.default_rng(n + card)
    k = Spec(pb.I64, rng.integers(0, card, n) * 1_000_003 - 17, nulls=rng.random(n) < 0.02)
    assert _check_rows(ctx, oracle, [k]) >= 1


def test_row_lists_multi_key_and_label_collisions(ctx, oracle):
    rng = np.random.default_rng(5)
    n = 50_000
    pool = ["a", "a_b", "b_c", "c", "NA", "NULL", ""]
    # no null_alias here: par_groupby labels a NULL "NA" and keeps a literal "NULL" string as its own group (grouping.rs:158-186),
    # unlike group_by, where the literal merges with the NULL group (grouping.rs:69-98)
    k0 = Spec(pb.DICT_U32, rng.integers(0, len(pool), n), nulls=rng.random(n) < 0.05, pool=pool)
    k1 = Spec(pb.DICT_U32, rng.integers(0, len(pool), n), pool=pool)
    k2 = Spec(pb.BOOL_BITS, rng.random(n) < 0.5, nulls=rng.random(n) < 0.05)
    k3 = Spec(pb.F64, rng.integers(0, 3, n) * 0.5, nulls=rng.random(n) < 0.05)
    _check_rows(ctx, oracle, [k0, k1])          # ("a_b", "c") and ("a", "b_c") collide; a literal "NA" collides with NULL
    _check_rows(ctx, oracle, [k0, k1, k2, k3])
    _check_rows(ctx, oracle, [Spec(pb.I32, rng.integers(-3, 3, n).astype(np.int32), nulls=rng.random(n) < 0.1), k2])


def test_frame_par_groupby_fixtures(fctx, oracle):
    # tests/optimized_groupby_test.rs:6-31 and :138-171 (the reference only asserts "not empty"; the oracle restates the rest)
    from pandrs_b200 import frame as F
    ctx = fctx
    df = F.OptimizedDataFrame()
    df.add_int_column("values", [10, 20, 30, 40, 50])
    df.add_string_column("keys", ["A", "B", "A", "B", "C"])
    g = df.par_groupby(["keys"])
    assert sorted(g) == ["A", "B", "C"]
    assert list(g["A"].column("values").values) == [10, 30] and g["A"].column("keys").to_list() == ["A", "A"]
    assert list(g["B"].column("values").values) == [20, 40] and list(g["C"].column("values").values) == [50]
    df = F.OptimizedDataFrame()
    df.add_int_column("values", [10, 20, 30, 40, 50, 60])
    df.add_string_column("category", ["X", "X", "Y", "Y", "X", "Y"])
    df.add_string_column("group", ["A", "B", "A", "B", "A", "B"])
    g = df.par_groupby(["category", "group"])
    assert {k: list(v.column("values").values) for k, v in g.items()} == {"X_A": [10, 50], "X_B": [20], "Y_A": [30], "Y_B": [40, 60]}
    # NULL keys -> "NA"; NULL values of the kept rows become defaults (filter_by_indices, data_ops.rs:124-211)
    df = F.OptimizedDataFrame()
    df.add_column("k", F.Int64Column([1, 2, 1, 2, 7], nulls=[False, True, False, True, False]))
    df.add_column("v", F.Float64Column([1.5, 2.5, 3.5, 4.5, 5.5], nulls=[False, False, True, False, False]))
    g = df.par_groupby(["k"])
    assert sorted(g) == ["1", "7", "NA"] and list(g["1"].column("v").values) == [1.5, 0.0] and list(g["NA"].column("k").values) == [0, 0]
    with pytest.raises(F.ColumnNotFound):
        df.par_groupby(["nope"])
    assert F.OptimizedDataFrame().par_groupby([]) == {}


@pytest.mark.parametrize("n,card", [(10, 3), (5000, 7), (60_000, 900), (200_000, 50_000)])
def test_median_first_last(ctx, oracle, n, card):
    rng = np.random.default_rng(n)
    k = Spec(pb.I64, rng.integers(0, card, n), nulls=rng.random(n) < 0.01)
    f = Spec(pb.F64, rng.normal(0, 100, n).round(1), nulls=rng.random(n) < 0.2)
    big = np.iinfo(np.int64).max
    iv = rng.integers(-50, 50, n)
    iv[rng.random(n) < 0.1] = big                      # ties with the NULL sentinel of the sort; mid sums that wrap
    iv[rng.random(n) < 0.05] = -big - 1
    i = Spec(pb.I64, iv, nulls=rng.random(n) < 0.2)
    if n == 5000:
        f.nulls[k.values == 3] = True                 # an all-NULL group -> 0.0
    ops = [pb.MEDIAN, pb.FIRST, pb.LAST]
    want = oracle.groupby([k.cpu(oracle)], [f.cpu(oracle), i.cpu(oracle)], [(0, op) for op in ops] + [(1, op) for op in ops])
    assert want["error"] == 0
    res = ctx.groupby_rows([k.gpu(pb)])
    try:
        kv, kn = res.key(0)
        where = {("NULL" if nl else str(int(v))): g for g, (v, nl) in enumerate(zip(kv, kn))}
        order = np.array([where[kt[0]] for kt in want["key_strings"]])
        assert len(order) == res.n_groups
        for c, col in enumerate((f, i)):
            for j, op in enumerate(ops):
                got = res.agg(col.gpu(pb), op)[order]
                assert np.array_equal(got, want["aggs"][3 * c + j]), (c, op)
        with pytest.raises(pb.PandrsError) as e:       # aggregation.rs:748-752
            res.agg(pb.Column.dict_ids(np.zeros(n, np.uint32)), pb.MEDIAN)
        assert e.value.kind == "OperationFailed"
    finally:
        res.close()


def test_frame_median_first_last(fctx, oracle):
    from pandrs_b200 import frame as F
    ctx = fctx
    df = F.OptimizedDataFrame()
    df.add_string_column("k", ["A", "B", "A", "B", "A", "C"])
    df.add_column("v", F.Int64Column([5, 20, 1, 40, 3, 9], nulls=[False, False, False, False, False, True]))
    out = df.group_by(["k"]).aggregate([("v", F.AggregateOp.Median, "med"), ("v", F.AggregateOp.First, "first"), ("v", F.AggregateOp.Last, "last"),
                                        ("v", F.AggregateOp.Sum, "sum")])
    rows = {k: (m, a, b, s) for k, m, a, b, s in zip(out.column("k").to_list(), out.column("med").values, out.column("first").values,
                                                        out.column("last").values, out.column("sum").values)}
    assert rows == {"A": (3.0, 5.0, 3.0, 9.0), "B": (30.0, 20.0, 40.0, 60.0), "C": (0.0, 0.0, 0.0, 0.0)}
    assert list(df.group_by(["k"]).median("v").column_names()) == ["k", "v_median"]


# ---------------------------------------------------------------- ingest
@pytest.mark.parametrize("n,off", [(1, 0), (7, 3), (64, 0), (1000, 5), (100_003, 13), (65_536, 8)])
def test_arrow_validity_to_null_mask(ctx, n, off):
    rng = np.random.default_rng(n + off)
    valid = rng.random(n + off) < 0.7
    bits = np.packbits(valid, bitorder="little")
    mask, nn = ctx.arrow_validity_to_nulls(bits, n, off)
    want = ~valid[off:off + n]
    assert nn == int(want.sum())
    assert np.array_equal(mask, np.packbits(want, bitorder="little"))      # trailing bits of the last byte are 0
    mask, nn = ctx.arrow_validity_to_nulls(None, n)
    assert nn == 0 and not mask.any()


def _py_encode(strings):
    pool, ids = {}, []
    for s in strings:
        ids.append(pool.setdefault(s, len(pool)))
    return np.array(ids, np.uint32), list(pool)


@pytest.mark.parametrize("n,card,large", [(1, 1, False), (1000, 10, False), (50_000, 5000, True), (200_000, 150_000, False)])
def test_dict_encode_first_occurrence_ids(ctx, n, card, large):
    pa = pytest.importorskip("pyarrow")
    rng = np.random.default_rng(n)
    words = ["", "x", "NULL"] + ["w%d-%s" % (i, "é" * (i % 7)) for i in range(card)]
    vals = [words[j] for j in rng.integers(0, len(words), n)]
    isnull = rng.random(n) < 0.05
    arr = pa.array([None if nl else v for v, nl in zip(vals, isnull)], type=pa.large_string() if large else pa.string())
    for a in (arr, arr.slice(3) if n > 10 else arr):
        m = len(a)
        bufs = a.buffers()
        off = np.frombuffer(bufs[1], np.int64 if large else np.int32)[a.offset:a.offset + m + 1]
        data = np.frombuffer(bufs[2], np.uint8) if bufs[2] is not None else np.empty(0, np.uint8)
        validity = None if bufs[0] is None else np.frombuffer(bufs[0], np.uint8)
        enc = ctx.dict_encode(off, data, validity, a.offset, m)
        try:
            want_ids, want_pool = _py_encode(["" if v is None else v for v in a.to_pylist()])     # a NULL row is the empty string
            assert enc.n_unique == len(want_pool)
            assert np.array_equal(enc.ids(), want_ids)
            fr = enc.first_rows()
            py = a.to_pylist()
            assert [("" if py[r] is None else py[r]) for r in fr] == want_pool
            enc.remap(np.arange(len(want_pool), dtype=np.uint32)[::-1].copy())
            assert np.array_equal(enc.ids(), len(want_pool) - 1 - want_ids)
        finally:
            enc.close()


def test_dict_encode_rejects_corrupt_offsets(ctx):
    data = np.frombuffer(b"abcdef", np.uint8)
    for off in ([0, 2, 1, 6], [0, 2, 4, 9], [-1, 2, 4, 6]):
        with pytest.raises(pb.PandrsError) as e:
            ctx.dict_encode(np.array(off, np.int32), data)
        assert e.value.kind == "InvalidInput"
    enc = ctx.dict_encode(np.array([0, 2, 4, 6], np.int64), data)
    assert enc.n_unique == 3 and list(enc.ids()) == [0, 1, 2]
    enc.close()


def test_from_record_batch_then_groupby(fctx, oracle):
    pa = pytest.importorskip("pyarrow")
    from pandrs_b200 import frame as F
    ctx = fctx
    rng = np.random.default_rng(77)
    n = 20_000
    cat = [None if rng.random() < 0.03 else "c%d" % rng.integers(0, 12) for _ in range(n)]
    qty = [None if rng.random() < 0.05 else int(rng.integers(-100, 100)) for _ in range(n)]
    px = [None if rng.random() < 0.05 else float(rng.normal()) for _ in range(n)]
    flag = [None if rng.random() < 0.05 else bool(rng.random() < 0.5) for _ in range(n)]
    batch = pa.record_batch({"cat": pa.array(cat, pa.string()), "qty": pa.array(qty, pa.int64()), "px": pa.array(px, pa.float64()), "flag": pa.array(flag, pa.bool_())})
    for b, lo in ((batch, 0), (batch.slice(11, n - 20), 11)):
        df = F.OptimizedDataFrame.from_record_batch(b)
        m = b.num_rows
        assert df.row_count() == m and df.column_names() == ["cat", "qty", "px", "flag"]
        for r in list(range(0, 40)) + [m - 1]:
            assert df.column("cat").get(r) == cat[lo + r] and df.column("qty").get(r) == qty[lo + r]
            assert df.column("px").get(r) == px[lo + r] and df.column("flag").get(r) == flag[lo + r]
        # the ingested frame and a frame built value by value give the same groupby
        ref = F.OptimizedDataFrame()
        sl = slice(lo, lo + m)
        ref.add_column("cat", F.StringColumn(["" if v is None else v for v in cat[sl]], nulls=[v is None for v in cat[sl]]))
        ref.add_column("qty", F.Int64Column([0 if v is None else v for v in qty[sl]], nulls=[v is None for v in qty[sl]]))
        ref.add_column("px", F.Float64Column([0.0 if v is None else v for v in px[sl]], nulls=[v is None for v in px[sl]]))
        aggs = [("qty", F.AggregateOp.Sum, "s"), ("px", F.AggregateOp.Mean, "m"), ("px", F.AggregateOp.Count, "c"), ("qty", F.AggregateOp.Max, "mx")]
        a, c = df.group_by(["cat"]).aggregate(aggs), ref.group_by(["cat"]).aggregate(aggs)
        ka, kc = a.column("cat").to_list(), c.column("cat").to_list()
        assert sorted(ka) == sorted(kc) and len(ka) == 13
        oa, oc = np.argsort(ka), np.argsort(kc)
        for name in ("s", "c", "mx"):
            assert np.array_equal(a.column(name).values[oa], c.column(name).values[oc]), name
        assert np.allclose(a.column("m").values[oa], c.column("m").values[oc], rtol=1e-12, atol=0)      # f64 sums: the order of the partial sums varies
