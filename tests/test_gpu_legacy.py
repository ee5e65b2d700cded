"""The LEGACY API the north star names - DataFrame::groupby over string-materialised Series (src/dataframe/groupby.rs:188-532) and
optimize_dataframe (src/optimized/convert.rs:13-110) - served through the same C entry points, against a pure-Python restatement of
the reference (oracle.legacy_groupby)."""
import numpy as np
import pytest

from pandrs_b200 import frame as F
from pandrs_b200 import legacy as L

pytestmark = pytest.mark.gpu

FUNCS = [(L.AggFunc.Sum, "sum"), (L.AggFunc.Mean, "mean"), (L.AggFunc.Min, "min"), (L.AggFunc.Max, "max"), (L.AggFunc.Count, "count"),
         (L.AggFunc.Std, "std"), (L.AggFunc.Var, "var"), (L.AggFunc.Median, "median")]


def _close(a: str, b: str) -> bool:
    if a == b:
        return True
    x, y = float(a), float(b)
    return abs(x - y) <= 1e-12 * max(abs(x), abs(y))


def _compare(oracle, cols, by, value_cols):
    df = L.DataFrame()
    for name, vals in cols.items():
        df.add_column(name, L.Series(vals, name))
    aggs = [L.NamedAgg(v, f, f"{v}_{n}") for v in value_cols for f, n in FUNCS]
    got = df.groupby(by).agg(aggs)
    want = oracle.legacy_groupby(cols, by, [(v, n, f"{v}_{n}") for v in value_cols for _, n in FUNCS])
    assert got.column_names() == list(by) + [a.alias for a in aggs]          # key columns first, then the aliases (groupby.rs:270-297)
    keys = list(zip(*[got.get_column_string_values(c) for c in by]))
    assert sorted(keys) == sorted(want) and len(set(keys)) == len(keys)
    for a in aggs:
        col = got.get_column_string_values(a.alias)
        exact = a.func in (L.AggFunc.Min, L.AggFunc.Max, L.AggFunc.Count, L.AggFunc.Median)
        for k, g in zip(keys, col):
            w = want[k][a.alias]
            assert (g == w) if exact else _close(g, w), (k, a.alias, g, w)
    return got


def test_legacy_groupby_goldens(fctx, oracle):
    # the value fixture of src/dataframe/pandas_compat/groupby.rs:480-510 through DataFrame::groupby
    ctx = fctx
    cols = {"category": ["A", "B", "A", "B", "A"], "value": ["10", "20", "30", "40", "50"], "score": ["1", "2", "3", "4", "5"]}
    got = _compare(oracle, cols, ["category"], ["value", "score"])
    rows = dict(zip(got.get_column_string_values("category"), zip(got.get_column_string_values("value_sum"), got.get_column_string_values("value_mean"),
                                                                   got.get_column_string_values("value_count"), got.get_column_string_values("value_std"))))
    assert rows == {"A": ("90", "30", "3", "20"), "B": ("60", "30", "2", "14.142135623730951")}
    df = L.DataFrame().add_column("category", L.Series(cols["category"])).add_column("value", L.Series(cols["value"]))
    g = df.groupby_single("category")
    assert g.ngroups() == 2
    assert g.sum("value").column_names() == ["category", "value_sum"]
    size = g.size()
    assert dict(zip(size.get_column_string_values("group"), size.get_column_string_values("size"))) == {"A": "3", "B": "2"}
    with pytest.raises(F.ColumnNotFound):
        df.groupby(["nope"])
    with pytest.raises(F.OperationFailed):
        g.agg([L.NamedAgg("value", L.AggFunc.Nunique, "u")])


def test_legacy_groupby_random_strings(fctx, oracle):
    # unparseable cells are skipped (Count = parseable cells), empty strings, a literal "NULL" key is just a string, inf / nan / exponents
    ctx = fctx
    rng = np.random.default_rng(3)
    n = 4000
    pool = ["x", "NULL", "", "7", "a_b"]
    cells = ["1.5", "-2", "abc", "", "1e3", "inf", "-inf", " 4", "5_0", "+.5", "3.", "0x10", "1e", "12345678901234567890"]      # (NaN: below - its place in a sorted group is unspecified)
    cols = {"k1": [pool[i] for i in rng.integers(0, len(pool), n)], "k2": [str(i) for i in rng.integers(0, 7, n)],
            "v": [cells[i] if rng.random() < 0.3 else repr(float(np.round(rng.normal(0, 50), 3))) for i in rng.integers(0, len(cells), n)],
            "w": [str(int(i)) for i in rng.integers(-1000, 1000, n)]}
    cols["v"][:3] = ["abc", "", " 4"]
    _compare(oracle, cols, ["k1"], ["v", "w"])
    _compare(oracle, cols, ["k1", "k2"], ["v"])
    # NaN cells: sums / means / std turn NaN, min / max ignore them (f64::min / f64::max), Count counts them
    df = L.DataFrame().add_column("k", L.Series(["a", "a", "a"])).add_column("v", L.Series(["nan", "1", "3"]))
    r = df.groupby(["k"]).agg([L.NamedAgg("v", f, nm) for f, nm in FUNCS if nm != "median"])
    assert [r.get_column_string_values(nm)[0] for _, nm in FUNCS if nm != "median"] == ["NaN", "NaN", "1", "3", "3", "NaN", "NaN"]
    # a group without a parseable cell -> 0 for every function
    cols2 = {"k": ["a", "a", "b"], "v": ["x", "", "2.5"]}
    got = _compare(oracle, cols2, ["k"], ["v"])
    i = got.get_column_string_values("k").index("a")
    assert all(got.get_column_string_values(f"v_{nm}")[i] == "0" for _, nm in FUNCS)


def test_optimize_dataframe_type_inference(fctx):
    # optimized/convert.rs:13-110
    df = L.DataFrame()
    df.add_column("i", L.Series(["1", "", "-3"])).add_column("f", L.Series(["1.5", "2", ""])).add_column("b", L.Series(["true", "0", ""]))
    df.add_column("s", L.Series(["a", "1", ""]))
    o = L.optimize_dataframe(df)
    assert [o.column_type(c) for c in "ifbs"] == [F.ColumnType.Int64, F.ColumnType.Float64, F.ColumnType.Boolean, F.ColumnType.String]
    assert list(o.column("i").values) == [1, 0, -3] and list(o.column("f").values) == [1.5, 2.0, 0.0]
    assert list(o.column("b").values) == [True, False, False] and o.column("s").to_list() == ["a", "1", ""]
    ctx = fctx
    out = o.group_by(["s"]).aggregate([("i", F.AggregateOp.Sum, "t")])
    assert dict(zip(out.column("s").to_list(), out.column("t").values)) == {"a": 1.0, "1": 0.0, "": -3.0}
