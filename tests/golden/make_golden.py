"""Writes tests/golden/pandrs_known_answers.json: every known-answer vector the pandrs test-suite holds for the
groupby-aggregate / join hot path, transcribed from the reference sources.

The reference is a Rust crate and cannot be built or imported in this image (no cargo / rustc), so the vectors
are transcribed by hand; when /root/reference is mounted, this script re-reads the cited files and checks that
every literal it transcribes is really there (so a typo here cannot silently become the "golden" value).
Vectors marked "derived" are computed here from the reference's formulas (aggregation.rs:500-754, 881-903)
with plain Python floats in the reference's evaluation order.

    python tests/golden/make_golden.py            # rewrites the JSON (run in the build container)
"""
import json
import math
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def ref_has(path, *literals):
    """True when every literal occurs in the reference file (whitespace-insensitive); None without the reference."""
    full = os.path.join(REF, path)
    if not os.path.exists(full):
        return None
    text = re.sub(r"\s+", "", open(full, encoding="utf-8").read())
    return all(re.sub(r"\s+", "", lit) in text for lit in literals)


def agg_reference(values, is_int):
    """aggregation.rs:500-754 on one group: row-order sums, two-pass variance with n-1."""
    n = len(values)
    if is_int:
        s = sum(int(v) for v in values)
        total = float(s)
        mean = float(s) / float(n)
    else:
        total = 0.0
        for v in values:
            total += v
        mean = total / n
    fv = [float(v) for v in values]
    m = 0.0
    for v in fv:
        m += v
    m /= n
    var = 0.0
    if n > 1:
        acc = 0.0
        for v in fv:
            acc += (v - m) ** 2
        var = acc / (n - 1)
    return {"count": float(n), "sum": total, "mean": mean, "min": float(min(values)), "max": float(max(values)), "std": math.sqrt(var), "var": var}


def main():
    checks = {}
    out = {"_about": "known-answer vectors of the pandrs tests for the groupby / join hot path; see make_golden.py", "vectors": []}

    # (i) src/dataframe/pandas_compat/groupby.rs:480-617
    src = "src/dataframe/pandas_compat/groupby.rs"
    checks[src] = ref_has(src, '"A".to_string(),"B".to_string(),"A".to_string(),"B".to_string(),"A".to_string(),', "vec![10.0,20.0,30.0,40.0,50.0]")
    out["vectors"].append({
        "name": "pandas_compat_groupby", "source": src + ":480-617", "kind": "groupby",
        "keys": ["A", "B", "A", "B", "A"], "values_f64": [10.0, 20.0, 30.0, 40.0, 50.0],
        "expect": {"A": {"sum": 90.0, "mean": 30.0, "min": 10.0, "max": 50.0, "count": 3.0, "std": 20.0},
                   "B": {"sum": 60.0, "mean": 30.0, "min": 20.0, "max": 40.0, "count": 2.0}}})

    # (ii) tests/optimized_join_test.rs:6-237
    src = "tests/optimized_join_test.rs"
    checks[src] = ref_has(src, "Int64Column::new(vec![1,2,3,4])", "Int64Column::new(vec![1,2,5,6])", "assert_eq!(joined.row_count(),2)")
    out["vectors"].append({
        "name": "optimized_join", "source": src + ":6-237", "kind": "join",
        "left_ids": [1, 2, 3, 4], "right_ids": [1, 2, 5, 6], "right_ids_disjoint": [5, 6, 7, 8],
        "inner_pairs": [[0, 0], [1, 1]], "left_pairs": [[0, 0], [1, 1], [2, -1], [3, -1]],
        "right_rows": 4, "outer_rows": 6, "disjoint_inner_rows": 0, "joined_columns": 3})

    # (iii) tests/concurrency_test.rs:351-384
    src = "tests/concurrency_test.rs"
    checks[src] = ref_has(src, "let data_size = 1000;", "categories[i % categories.len()]", "assert_eq!(grouped.len(), 4)")
    out["vectors"].append({"name": "concurrency_four_groups", "source": src + ":351-384", "kind": "groupby_counts",
                           "rows": 1000, "modulus": 4, "expect_groups": 4, "expect_rows_per_group": 250})

    # (iv) tests/optimized_groupby_enhanced_test.rs:12-23 (fixture) + formulas of aggregation.rs:500-754 (derived)
    src = "tests/optimized_groupby_enhanced_test.rs"
    checks[src] = ref_has(src, '["A","B","A","B","A","C","B","C","C","A"]', "vec![10,25,15,30,22,18,24,12,16,20]",
                          "vec![1.1,2.2,3.3,4.4,5.5,6.6,7.7,8.8,9.9,10.0]")
    group = "A B A B A C B C C A".split()
    value = [10, 25, 15, 30, 22, 18, 24, 12, 16, 20]
    flt = [1.1, 2.2, 3.3, 4.4, 5.5, 6.6, 7.7, 8.8, 9.9, 10.0]
    exp_v, exp_f = {}, {}
    for g in "ABC":
        exp_v[g] = agg_reference([v for k, v in zip(group, value) if k == g], True)
        exp_f[g] = agg_reference([v for k, v in zip(group, flt) if k == g], False)
    out["vectors"].append({"name": "ten_row_fixture", "source": src + ":12-23", "kind": "groupby", "derived": "aggregation.rs:500-754, 881-903",
                           "group": group, "value_i64": value, "float_f64": flt, "expect_value": exp_v, "expect_float": exp_f})

    # (v) tests/optimized_custom_aggregation_test.rs:49 - total of `value`
    src = "tests/optimized_custom_aggregation_test.rs"
    checks[src] = ref_has(src, "192")
    out["vectors"].append({"name": "ten_row_total", "source": src + ":49", "kind": "scalar", "total_value": 192})

    # (vi) tests/optimized_groupby_test.rs:100-184 - LazyFrame aggregate schema
    src = "tests/optimized_groupby_test.rs"
    checks[src] = ref_has(src, "assert_eq!(result.column_count(),6)", "assert_eq!(result.column_count(),3)")
    out["vectors"].append({"name": "lazyframe_schema", "source": src + ":100-184", "kind": "schema",
                           "single_key_columns": ["keys", "count", "sum", "mean", "min", "max"], "multi_key_column_count": 3})

    out["reference_literals_verified"] = checks
    with open(os.path.join(HERE, "pandrs_known_answers.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(checks, indent=1))
    bad = [k for k, v in checks.items() if v is False]
    if bad:
        raise SystemExit(f"literals not found in the reference: {bad}")


if __name__ == "__main__":
    main()
