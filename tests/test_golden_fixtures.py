"""The committed known-answer fixtures (tests/golden/pandrs_known_answers.json, transcribed from the pandrs
test-suite by tests/golden/make_golden.py) against the CPU oracle (CPU) and against the CUDA path through the
frame mirror of the reference API (GPU)."""
import json
import os

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pandrs_known_answers.json")))
VEC = {v["name"]: v for v in GOLD["vectors"]}
OPS = ["sum", "mean", "min", "max", "count", "std"]


def test_fixture_literals_were_verified_against_the_reference():
    # make_golden.py re-reads the cited reference files when they are mounted; a False here is a transcription error
    assert all(v is not False for v in GOLD["reference_literals_verified"].values())
    assert len(VEC) >= 6


def _dict(strings):
    pool, ids = [], []
    for s in strings:
        if s not in pool:
            pool.append(s)
        ids.append(pool.index(s))
    return ids, pool


def _oracle_groupby(o, keys, col, ops):
    ids, pool = _dict(keys)
    res = o.groupby([o.Col(o.DICT_U32, ids, pool=pool)], [col], [(0, getattr(o, op.upper())) for op in ops])
    return {k[0]: {op: float(res["aggs"][a][i]) for a, op in enumerate(ops)} for i, k in enumerate(res["key_strings"])}


def test_oracle_matches_pandas_compat_goldens(oracle):
    v = VEC["pandas_compat_groupby"]
    got = _oracle_groupby(oracle, v["keys"], oracle.Col(oracle.F64, v["values_f64"]), OPS)
    for g, exp in v["expect"].items():
        for op, want in exp.items():
            assert got[g][op] == want, (g, op)


def test_oracle_matches_ten_row_fixture(oracle):
    v = VEC["ten_row_fixture"]
    gv = _oracle_groupby(oracle, v["group"], oracle.Col(oracle.I64, v["value_i64"]), OPS)
    gf = _oracle_groupby(oracle, v["group"], oracle.Col(oracle.F64, v["float_f64"]), OPS)
    for g in "ABC":
        for op in OPS:
            assert gv[g][op] == v["expect_value"][g][op], (g, op)
            assert gf[g][op] == v["expect_float"][g][op], (g, op)
    assert sum(v["value_i64"]) == VEC["ten_row_total"]["total_value"]


def test_oracle_matches_join_goldens(oracle):
    v = VEC["optimized_join"]
    o = oracle
    L, R = o.Col(o.I64, v["left_ids"]), o.Col(o.I64, v["right_ids"])
    li, ri = o.join(L, R, o.INNER)
    assert [list(p) for p in zip(li.tolist(), ri.tolist())] == v["inner_pairs"]
    li, ri = o.join(L, R, o.LEFT)
    assert [list(p) for p in zip(li.tolist(), ri.tolist())] == v["left_pairs"]
    assert len(o.join(L, R, o.RIGHT)[0]) == v["right_rows"] and len(o.join(L, R, o.OUTER)[0]) == v["outer_rows"]
    assert len(o.join(L, o.Col(o.I64, v["right_ids_disjoint"]), o.INNER)[0]) == v["disjoint_inner_rows"]


# ---------------------------------------------------------------- the CUDA path through the mirror of the reference API
@pytest.fixture(scope="module")
def frame():
    import pandrs_b200.frame as fr
    yield fr
    ctx = fr._CTX
    fr.set_context(None)
    if ctx is not None:
        ctx.close()


@pytest.mark.gpu
def test_gpu_ten_row_fixture_through_group_by(frame):
    # tests/optimized_groupby_enhanced_test.rs:12-23 + SURVEY.md §9.6 (iv)
    fr, v = frame, VEC["ten_row_fixture"]
    df = fr.OptimizedDataFrame.new()
    df.add_string_column("group", v["group"])
    df.add_int_column("value", v["value_i64"])
    df.add_float_column("float", v["float_f64"])
    assert df.row_count() == 10 and df.column_count() == 3
    A = fr.AggregateOp
    res = df.group_by(["group"]).agg([("value", op) for op in (A.Sum, A.Mean, A.Min, A.Max, A.Count, A.Std)] +
                                     [("float", op) for op in (A.Sum, A.Mean, A.Min, A.Max, A.Count, A.Std)])
    assert res.row_count() == 3 and res.column_count() == 13
    keys = res.column("group").to_list()
    for op in OPS:
        cv, cf = res.column(f"value_{op}").values, res.column(f"float_{op}").values
        for i, g in enumerate(keys):
            wv, wf = v["expect_value"][g][op], v["expect_float"][g][op]
            if op in ("min", "max", "count") or op == "sum":
                assert cv[i] == wv, (g, op)
            else:
                assert abs(cv[i] - wv) <= 1e-12 * abs(wv), (g, op)
            if op in ("min", "max", "count"):
                assert cf[i] == wf, (g, op)
            else:
                assert abs(cf[i] - wf) <= 1e-12 * abs(wf), (g, op)
    # shortcuts name their column "<col>_<op>" (operations.rs:438-521)
    assert res.contains_column("value_sum") and df.group_by(["group"]).sum("value").column_names() == ["group", "value_sum"]


@pytest.mark.gpu
def test_gpu_pandas_compat_goldens(frame):
    fr, v = frame, VEC["pandas_compat_groupby"]
    df = fr.OptimizedDataFrame.new()
    df.add_string_column("category", v["keys"])
    df.add_float_column("value", v["values_f64"])
    A = fr.AggregateOp
    res = df.group_by(["category"]).agg([("value", op) for op in (A.Sum, A.Mean, A.Min, A.Max, A.Count, A.Std)])
    keys = res.column("category").to_list()
    for i, g in enumerate(keys):
        for op, want in v["expect"][g].items():
            assert res.column(f"value_{op}").values[i] == want, (g, op)


@pytest.mark.gpu
def test_gpu_join_tests_of_the_reference(frame):
    # tests/optimized_join_test.rs:6-237
    fr, v = frame, VEC["optimized_join"]
    left = fr.OptimizedDataFrame.new()
    left.add_column("id", fr.Int64Column(v["left_ids"]))
    left.add_column("name", fr.StringColumn(["Alice", "Bob", "Charlie", "Dave"]))
    right = fr.OptimizedDataFrame.new()
    right.add_column("id", fr.Int64Column(v["right_ids"]))
    right.add_column("value", fr.Int64Column([100, 200, 500, 600]))
    joined = left.inner_join(right, "id", "id")
    assert joined.row_count() == 2 and joined.column_count() == v["joined_columns"]
    assert joined.column_names() == ["name", "id", "value"]                      # join.rs:290-552 column order
    assert joined.column("id").values.tolist() == [1, 2] and joined.column("value").values.tolist() == [100, 200]
    assert joined.column("name").to_list() == ["Alice", "Bob"]
    joined = left.left_join(right, "id", "id")
    assert joined.row_count() == 4 and joined.column_count() == 3
    assert joined.column("value").values.tolist() == [100, 200, 0, 0]            # missing side -> type default, no null mask
    joined = left.right_join(right, "id", "id")                                  # test_right_join: ids 1, 2, 5, 6
    assert joined.row_count() == v["right_rows"] and joined.column_count() == 3
    assert sorted(joined.column("id").values.tolist()) == [1, 2, 5, 6] and sorted(joined.column("value").values.tolist()) == [100, 200, 500, 600]
    assert sorted(joined.column("name").to_list()) == ["", "", "Alice", "Bob"]   # right-only rows: left columns take the type default
    joined = left.outer_join(right, "id", "id")                                  # test_outer_join: ids 1..6
    assert joined.row_count() == v["outer_rows"] and sorted(joined.column("id").values.tolist()) == [1, 2, 3, 4, 5, 6]
    # test_empty_join: no matching ids -> 0 rows
    other = fr.OptimizedDataFrame.new()
    other.add_column("id", fr.Int64Column(v["right_ids_disjoint"]))
    other.add_column("value", fr.Int64Column([500, 600, 700, 800]))
    assert left.inner_join(other, "id", "id").row_count() == 0
    # test_join_different_column_names + "_right" suffix rule
    r2 = fr.OptimizedDataFrame.new()
    r2.add_column("right_id", fr.Int64Column(v["right_ids"]))
    r2.add_column("name", fr.StringColumn(["w", "x", "y", "z"]))
    j = left.inner_join(r2, "id", "right_id")
    assert j.row_count() == 2 and j.column_names() == ["name", "id", "name_right"]
    # schema errors stay on the host side (join.rs:98-104)
    with pytest.raises(fr.ColumnNotFound):
        left.inner_join(right, "nope", "id")
    fl = fr.OptimizedDataFrame.new()
    fl.add_column("id", fr.Float64Column([1.0, 2.0]))
    with pytest.raises(fr.ColumnTypeMismatch):
        left.inner_join(fl, "id", "id")


@pytest.mark.gpu
def test_gpu_lazyframe_and_concurrency_fixture(frame):
    # tests/optimized_groupby_test.rs:100-184 and tests/concurrency_test.rs:351-384
    fr = frame
    A = fr.AggregateOp
    df = fr.OptimizedDataFrame.new()
    df.add_column("values", fr.Int64Column([10, 20, 30, 40, 50, 60]))
    df.add_column("keys", fr.StringColumn(["A", "B", "A", "B", "A", "C"]))
    res = fr.LazyFrame.new(df).aggregate(["keys"], [("values", A.Count, "count"), ("values", A.Sum, "sum"), ("values", A.Mean, "mean"),
                                                    ("values", A.Min, "min"), ("values", A.Max, "max")]).execute()
    assert res.row_count() == 3 and res.column_names() == VEC["lazyframe_schema"]["single_key_columns"]
    got = dict(zip(res.column("keys").to_list(), res.column("sum").values.tolist()))
    assert got == {"A": 90.0, "B": 60.0, "C": 60.0}
    with pytest.raises(fr.OperationFailed):                                       # lazy.rs:267-383 has no Std
        fr.LazyFrame.new(df).aggregate(["keys"], [("values", A.Std, "std")]).execute()
    df2 = fr.OptimizedDataFrame.new()
    df2.add_column("values", fr.Int64Column([10, 20, 30, 40, 50, 60]))
    df2.add_column("category", fr.StringColumn(["X", "X", "Y", "Y", "X", "Y"]))
    df2.add_column("group", fr.StringColumn(["A", "B", "A", "B", "A", "B"]))
    res = fr.LazyFrame.new(df2).aggregate(["category", "group"], [("values", A.Sum, "sum")]).execute()
    assert res.row_count() == 4 and res.column_count() == VEC["lazyframe_schema"]["multi_key_column_count"]
    # multi-key group_by: keys live in the StringMultiIndex, the frame holds the aggregates only (aggregation.rs:812-853)
    mi = df2.group_by(["category", "group"]).sum("values")
    assert mi.column_names() == ["values_sum"] and sorted(mi.index) == [("X", "A"), ("X", "B"), ("Y", "A"), ("Y", "B")]
    flat = df2.group_by_with_options(["category", "group"], False).sum("values")
    assert flat.column_count() == 3
    c = VEC["concurrency_four_groups"]
    big = fr.OptimizedDataFrame.new()
    cats = ["A", "B", "C", "D"]
    big.add_string_column("category", [cats[i % c["modulus"]] for i in range(c["rows"])])
    big.add_int_column("value", list(range(c["rows"])))
    r = big.group_by(["category"]).count("value")
    assert r.row_count() == c["expect_groups"] and set(r.column("value_count").values.tolist()) == {float(c["expect_rows_per_group"])}
    # errors mirror Error::ColumnNotFound / OperationFailed
    with pytest.raises(fr.ColumnNotFound):
        big.group_by(["missing"])
    with pytest.raises(fr.OperationFailed):
        big.group_by(["value"]).aggregate([("category", A.Sum, "s")])             # aggregation.rs:748-752
    assert big.group_by(["value"]).par_aggregate([("category", A.Sum, "s")]).column("s").values.sum() == 0.0   # :114-117


@pytest.mark.gpu
def test_gpu_filter_then_aggregate(frame):
    # data_ops.rs:37-121 followed by grouping: NULLs of the kept rows become defaults
    fr = frame
    A = fr.AggregateOp
    df = fr.OptimizedDataFrame.new()
    df.add_column("k", fr.Int64Column([1, 1, 2, 2, 2, 3]))
    df.add_column("v", fr.Float64Column([1.0, 2.0, 3.0, 4.0, 5.0, 6.0], nulls=[False, True, False, False, True, False]))
    df.add_column("keep", fr.BooleanColumn([True, True, False, True, True, True], nulls=[False, False, False, False, False, True]))
    two_step = df.filter("keep").group_by(["k"]).agg([("v", A.Sum), ("v", A.Count), ("v", A.Mean)])
    fused = fr.LazyFrame.new(df).filter("keep").aggregate(["k"], [("v", A.Sum, "v_sum"), ("v", A.Count, "v_count"), ("v", A.Mean, "v_mean")]).execute()
    for res in (two_step, fused):
        got = {k: (s, c, m) for k, s, c, m in zip(res.column("k").to_list(), res.column("v_sum").values, res.column("v_count").values, res.column("v_mean").values)}
        assert got == {"1": (1.0, 2.0, 0.5), "2": (4.0, 2.0, 2.0)}
