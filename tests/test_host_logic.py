"""Host-side logic of the mirrors that needs no GPU: Rust's f64 / i64 parse grammar, optimize_dataframe's type inference
(src/optimized/convert.rs:13-110), cell formatting like `to_string()`."""
import numpy as np

from pandrs_b200 import frame as F
from pandrs_b200 import legacy as L


def test_rust_parse_grammar():
    ok = {"1": 1.0, "-2.5": -2.5, "+.5": 0.5, "3.": 3.0, "1e3": 1000.0, "1E-2": 0.01, "inf": float("inf"), "-Infinity": float("-inf"), "007": 7.0}
    for s, v in ok.items():
        assert L._parse_f64(s) == v, s
    assert np.isnan(L._parse_f64("NaN"))
    for s in ("", " 4", "4 ", "1_0", "0x10", "1e", ".", "e5", "--1", "1,5", "nan0"):
        assert L._parse_f64(s) is None, s
    assert L._parse_i64("-42") == -42 and L._parse_i64("+7") == 7
    for s in ("", "4.0", " 4", "9223372036854775808", "1e3"):
        assert L._parse_i64(s) is None, s
    assert L._parse_i64("-9223372036854775808") == -(1 << 63)


def test_optimize_dataframe_inference_order():
    df = L.DataFrame()
    df.add_column("ints", L.Series(["1", "", "-3"]))
    df.add_column("floats", L.Series(["1", "2.5", ""]))
    df.add_column("bools", L.Series(["TRUE", "false", ""]))
    df.add_column("zero_one", L.Series(["0", "1", "1"]))          # parses as Int64 first (convert.rs:33-47)
    df.add_column("strings", L.Series(["a", "1", ""]))
    o = L.optimize_dataframe(df)
    assert [o.column_type(c) for c in o.column_names()] == [F.ColumnType.Int64, F.ColumnType.Float64, F.ColumnType.Boolean, F.ColumnType.Int64, F.ColumnType.String]
    assert list(o.column("ints").values) == [1, 0, -3] and list(o.column("floats").values) == [1.0, 2.5, 0.0]
    assert list(o.column("bools").values) == [True, False, False]
    assert o.row_count() == 3


def test_cells_are_formatted_like_to_string():
    s = L.Series([1, 2.5, True, "x", 90.0, float("nan"), -0.0, 1e21])
    assert s.values == ["1", "2.5", "true", "x", "90", "NaN", "-0", "1000000000000000000000"]
