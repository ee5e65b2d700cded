"""Pins the CPU oracle against every known-answer vector the pandrs tests hold for the hot path
(SURVEY.md §8c / §9.6).  CPU only."""
import numpy as np
import pytest

A = "A B A B A C B C C A".split()
VALUE = [10, 25, 15, 30, 22, 18, 24, 12, 16, 20]
FLOAT = [1.1, 2.2, 3.3, 4.4, 5.5, 6.6, 7.7, 8.8, 9.9, 10.0]


def _dict(strings):
    pool = []
    ids = []
    for s in strings:
        if s not in pool:
            pool.append(s)
        ids.append(pool.index(s))
    return ids, pool


def _by_key(res):
    return {k[0] if len(k) == 1 else k: i for i, k in enumerate(res["key_strings"])}


def test_pandas_compat_groupby_goldens(oracle):
    # src/dataframe/pandas_compat/groupby.rs:480-617
    o = oracle
    ids, pool = _dict(["A", "B", "A", "B", "A"])
    key = o.Col(o.DICT_U32, ids, pool=pool)
    val = o.Col(o.F64, [10.0, 20.0, 30.0, 40.0, 50.0])
    ops = [o.SUM, o.MEAN, o.MIN, o.MAX, o.COUNT, o.STD]
    res = o.groupby([key], [val], [(0, op) for op in ops])
    assert res["n_groups"] == 2
    g = _by_key(res)
    want = {"A": [90.0, 30.0, 10.0, 50.0, 3.0, 20.0], "B": [60.0, 30.0, 20.0, 40.0, 2.0, None]}
    for k, w in want.items():
        for a, x in enumerate(w):
            if x is not None:
                assert res["aggs"][a][g[k]] == x, (k, a)
    assert abs(res["aggs"][5][g["B"]] - np.std([20.0, 40.0], ddof=1)) < 1e-12


def test_ten_row_fixture(oracle):
    # tests/optimized_groupby_enhanced_test.rs:12-23, goldens in SURVEY.md §9.6 (iv);
    # the total 192 is asserted by tests/optimized_custom_aggregation_test.rs:49
    o = oracle
    ids, pool = _dict(A)
    key = o.Col(o.DICT_U32, ids, pool=pool)
    vi = o.Col(o.I64, VALUE)
    vf = o.Col(o.F64, FLOAT)
    ops = [o.COUNT, o.SUM, o.MEAN, o.MIN, o.MAX, o.STD]
    res = o.groupby([key], [vi, vf], [(0, op) for op in ops] + [(1, op) for op in ops])
    g = _by_key(res)
    want_i = {"A": [4, 67, 16.75, 10, 22, 5.377421934967226],
              "B": [3, 79, 26.333333333333332, 24, 30, 3.2145502536643185],
              "C": [3, 46, 15.333333333333334, 12, 18, 3.055050463303893]}
    want_f = {"A": [4, 19.9, 4.975, 1.1, 10.0, 3.8012059489939065],
              "B": [3, 14.3, 4.766666666666667, 2.2, 7.7, 2.768272626265942],
              "C": [3, 25.3, 8.433333333333334, 6.6, 9.9, 1.680277754817142]}
    for k in "ABC":
        for a in range(6):
            assert res["aggs"][a][g[k]] == pytest.approx(want_i[k][a], rel=1e-15, abs=0), (k, a)
            assert res["aggs"][6 + a][g[k]] == pytest.approx(want_f[k][a], rel=2e-15, abs=0), (k, a)
    assert sum(res["aggs"][1]) == 192


def test_concurrency_fixture_four_groups(oracle):
    # tests/concurrency_test.rs:351-384 — 1,000 rows, category i % 4 -> 4 groups of 250
    o = oracle
    key = o.Col(o.I64, np.arange(1000) % 4)
    res = o.groupby([key], [o.Col(o.F64, np.arange(1000, dtype=np.float64))], [(0, o.COUNT)])
    assert res["n_groups"] == 4
    assert list(res["aggs"][0]) == [250.0] * 4


def test_join_goldens(oracle):
    # tests/optimized_join_test.rs:6-237
    o = oracle
    L = o.Col(o.I64, [1, 2, 3, 4])
    R = o.Col(o.I64, [1, 2, 5, 6])
    li, ri = o.join(L, R, o.INNER)
    assert list(zip(li, ri)) == [(0, 0), (1, 1)]
    li, ri = o.join(L, R, o.LEFT)
    assert list(zip(li, ri)) == [(0, 0), (1, 1), (2, -1), (3, -1)]
    li, ri = o.join(L, R, o.RIGHT)
    assert list(zip(li, ri)) == [(0, 0), (1, 1), (-1, 2), (-1, 3)]
    li, ri = o.join(L, R, o.OUTER)
    assert len(li) == 6 and list(zip(li, ri))[4:] == [(-1, 2), (-1, 3)]
    li, ri = o.join(L, o.Col(o.I64, [5, 6, 7, 8]), o.INNER)
    assert len(li) == 0


def test_join_merge_ordering_golden(oracle):
    # src/dataframe/pandas_compat/merge.rs:320-410 — left keys A,B,C,D / right B,C,D,E
    o = oracle
    pool = list("ABCDE")
    L = o.Col(o.DICT_U32, [0, 1, 2, 3], pool=pool)
    R = o.Col(o.DICT_U32, [1, 2, 3, 4], pool=pool)
    li, ri = o.join(L, R, o.INNER)
    assert [pool[L.data[i]] for i in li] == ["B", "C", "D"]
    li, ri = o.join(L, R, o.LEFT)
    assert [pool[L.data[i]] for i in li] == ["A", "B", "C", "D"] and ri[0] == -1
    li, ri = o.join(L, R, o.OUTER)
    keys = [pool[L.data[i]] if i >= 0 else pool[R.data[r]] for i, r in zip(li, ri)]
    assert keys == ["A", "B", "C", "D", "E"]


def test_null_semantics(oracle):
    # NULL keys form one "NULL" group; Count counts NULL values; all-NULL -> 0.0;
    # NULL join keys are dropped on both sides, also for Left.
    o = oracle
    key = o.Col(o.I64, [7, 7, 0, 0, 9], nulls=o.pack_bits([0, 0, 1, 1, 0]))
    val = o.Col(o.F64, [1.0, 2.0, 3.0, 4.0, 5.0], nulls=o.pack_bits([0, 1, 0, 0, 1]))
    ops = [o.SUM, o.MEAN, o.MIN, o.MAX, o.COUNT, o.STD]
    res = o.groupby([key], [val], [(0, op) for op in ops])
    g = _by_key(res)
    assert set(g) == {"7", "NULL", "9"}
    assert [res["aggs"][a][g["7"]] for a in range(6)] == [1.0, 1.0, 1.0, 1.0, 2.0, 0.0]
    assert [res["aggs"][a][g["9"]] for a in range(6)] == [0.0, 0.0, 0.0, 0.0, 1.0, 0.0]
    assert res["aggs"][0][g["NULL"]] == 7.0 and res["aggs"][4][g["NULL"]] == 2.0
    L = o.Col(o.I64, [1, 2, 3], nulls=o.pack_bits([0, 1, 0]))
    R = o.Col(o.I64, [1, 2, 3], nulls=o.pack_bits([0, 0, 1]))
    li, ri = o.join(L, R, o.LEFT)
    assert list(zip(li, ri)) == [(0, 0), (2, -1)]
    # short null masks: bytes past the end read as "not NULL" (int64_column.rs:77)
    k2 = o.Col(o.I64, np.arange(20) % 2, nulls=np.array([0xFF], np.uint8))
    res = o.groupby([k2], [o.Col(o.F64, np.ones(20))], [(0, o.COUNT)])
    g = _by_key(res)
    assert res["aggs"][0][g["NULL"]] == 8 and res["aggs"][0][g["0"]] == 6 and res["aggs"][0][g["1"]] == 6


def test_int_sentinels_and_wrapping(oracle):
    # aggregation.rs:531-556 sentinel collapse; :508-513 wrapping i64 sum in release builds
    o = oracle
    key = o.Col(o.I64, [1, 1, 2, 2])
    val = o.Col(o.I64, [2**63 - 1, 1, -(2**63), -(2**63)])
    res = o.groupby([key], [val], [(0, o.SUM), (0, o.MIN), (0, o.MAX)])
    g = _by_key(res)
    assert res["aggs"][0][g["1"]] == float(-(2**63))       # wrapped
    assert res["aggs"][2][g["1"]] == float(2**63 - 1)      # i64::MAX is a sentinel for Min only
    assert res["aggs"][1][g["1"]] == 1.0
    assert res["aggs"][2][g["2"]] == 0.0                   # max == i64::MIN -> 0.0
    assert res["aggs"][0][g["2"]] == 0.0                   # MIN+MIN wraps to 0


def test_lazy_mode_rejects_std(oracle):
    # lazy.rs:377-382
    o = oracle
    res = o.groupby([o.Col(o.I64, [1, 2])], [o.Col(o.F64, [1.0, 2.0])], [(0, o.STD)], mode=o.MODE_LAZY)
    assert res["error"] == 1
    res = o.groupby([o.Col(o.I64, [1, 2])], [o.Col(o.F64, [1.0, 2.0])], [(0, o.SUM)], mode=o.MODE_LAZY)
    assert res["error"] == 0


def test_par_aggregate_equals_serial(oracle):
    o = oracle
    n = 20000
    key = o.Col(o.I64, o.synth_keys(n, card=37))
    val = o.Col(o.F64, o.synth_vals(n), nulls=o.synth_nulls(n))
    aggs = [(0, op) for op in (o.SUM, o.MEAN, o.MIN, o.MAX, o.COUNT, o.STD)]
    a = o.groupby([key], [val], aggs)
    b = o.groupby([key], [val], aggs, mode=o.MODE_PAR_AGGREGATE, nthreads=4)
    assert a["key_strings"] == b["key_strings"]
    for x, y in zip(a["aggs"], b["aggs"]):
        assert np.array_equal(x, y)


def test_against_pandas(oracle):
    import pandas as pd
    o = oracle
    n = 5000
    k1 = o.synth_keys(n, seed=1, card=13)
    k2 = o.synth_keys(n, seed=2, card=7)
    v = o.synth_vals(n, seed=3)
    res = o.groupby([o.Col(o.I64, k1), o.Col(o.I64, k2)], [o.Col(o.F64, v)],
                    [(0, o.SUM), (0, o.MEAN), (0, o.MIN), (0, o.MAX), (0, o.COUNT), (0, o.STD)])
    df = pd.DataFrame(dict(k1=k1, k2=k2, v=v)).groupby(["k1", "k2"])["v"].agg(["sum", "mean", "min", "max", "count", "std"])
    assert res["n_groups"] == len(df)
    for g, ks in enumerate(res["key_strings"]):
        row = df.loc[(int(ks[0]), int(ks[1]))]
        got = [res["aggs"][a][g] for a in range(6)]
        assert np.allclose(got, row.values, rtol=1e-12, atol=0)
    # join vs pandas merge
    lk = o.synth_keys(300, seed=5, card=50)
    rk = o.synth_keys(200, seed=6, card=50)
    li, ri = o.join(o.Col(o.I64, lk), o.Col(o.I64, rk), o.INNER)
    m = pd.DataFrame(dict(k=lk, l=np.arange(300))).merge(pd.DataFrame(dict(k=rk, r=np.arange(200))), on="k")
    assert sorted(zip(li, ri)) == sorted(zip(m.l, m.r))


def test_gather_and_filter(oracle):
    o = oracle
    col = o.Col(o.F64, [1.5, 2.5, 3.5], nulls=o.pack_bits([0, 1, 0]))
    assert list(o.gather(col, [2, -1, 1, 0])) == [3.5, 0.0, 0.0, 1.5]
    mask = o.Col(o.BOOL_BITS, o.pack_bits([1, 0, 1, 1, 0]), nulls=o.pack_bits([0, 0, 1, 0, 0]), length=5)
    assert list(o.filter_indices(mask)) == [0, 3]


def test_synth_matches_numpy_restatement(oracle):
    # the generators are shared with the CUDA side; pin them with an independent numpy restatement
    o = oracle

    def sm(x):
        x = (x + np.uint64(0x9E3779B97F4A7C15))
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))
    with np.errstate(over="ignore"):
        rows = np.arange(1000, dtype=np.uint64)
        k = sm(np.uint64(42) * np.uint64(0x100000001B3) + rows) % np.uint64(1000)
        assert np.array_equal(o.synth_keys(1000), k.astype(np.int64))
        r = sm(np.uint64(43) * np.uint64(0x100000001B3) + rows)
        assert np.array_equal(o.synth_vals(1000), (r >> np.uint64(11)).astype(np.float64) * (1000.0 / 2**53))


# ---------------------------------------------------------------- typed-key oracle (oracle/typed_oracle.cpp)
# The checker for BASELINE.json's full sizes.  It is pinned by the string oracle above: identical results, bit for bit,
# on inputs that exercise every rule of SURVEY.md §9.1-9.3.
def _typed_key_string(o, dtype, v, isnull, pool):
    if isnull:
        return "NULL"
    if dtype in (o.I64, o.I32):
        return str(int(np.uint64(v).astype(np.int64)))
    if dtype == o.F64:
        d = float(np.uint64(v).view(np.float64))
        if np.isnan(d):
            return "NaN"
        if np.isinf(d):
            return "-inf" if d < 0 else "inf"
        return np.format_float_positional(d, trim="-")
    if dtype == o.DICT_U32:
        return pool[int(v)] if pool is not None else "#" + str(int(v))
    return "true" if v else "false"


def _assert_typed_equals_string(o, keys, val, null_alias=None, nthreads=3):
    ops = [o.SUM, o.MEAN, o.MIN, o.MAX, o.COUNT, o.STD, o.VAR]
    want = o.groupby(keys, [val], [(0, op) for op in ops])
    got = o.typed_groupby(keys, val, null_alias=null_alias, nthreads=nthreads)
    assert got["n_groups"] == want["n_groups"]
    index = {kt: g for g, kt in enumerate(want["key_strings"])}
    for g in range(got["n_groups"]):
        kt = tuple(_typed_key_string(o, k.dtype, got["keys"][c][0][g], got["keys"][c][1][g], k.pool) for c, k in enumerate(keys))
        j = index[kt]
        assert got["group_rows"][g] == want["group_rows"][j] and got["first_row"][g] == want["first_row"][j]
        for a, name in enumerate(["sum", "mean", "min", "max", None, "std", "var"]):
            w = want["aggs"][a][j]
            x = got["group_rows"][g] if name is None else got[name][g]
            assert x == w or (np.isnan(x) and np.isnan(w)), (kt, name, x, w)      # bit-identical, not "close"


def test_typed_oracle_is_bit_identical_to_the_string_oracle(oracle):
    o = oracle
    rng = np.random.default_rng(21)
    n = 40_000
    pool = [f"s{i}" for i in range(40)] + ["NULL"]
    ki = o.Col(o.I64, rng.integers(-30, 30, n) * 10**12, o.pack_bits(rng.random(n) < 0.05))
    k32 = o.Col(o.I32, rng.integers(-5, 5, n).astype(np.int32))
    kd = o.Col(o.DICT_U32, rng.integers(0, 41, n).astype(np.uint32), o.pack_bits(rng.random(n) < 0.03), pool=pool)
    kb = o.Col(o.BOOL_BITS, o.pack_bits(rng.random(n) < 0.5), length=n)
    fvals = rng.integers(0, 6, n).astype(np.float64) / 4
    fvals[rng.random(n) < 0.02] = np.nan
    fvals[rng.random(n) < 0.02] = -0.0
    kf = o.Col(o.F64, fvals)
    v = rng.normal(1e6, 3.0, n)
    v[rng.random(n) < 0.01] = np.nan
    v[rng.random(n) < 0.005] = np.inf
    vf = o.Col(o.F64, v, o.pack_bits(rng.random(n) < 0.1))
    vi = o.Col(o.I64, rng.integers(-2**62, 2**62, n), o.pack_bits(rng.random(n) < 0.1))
    _assert_typed_equals_string(o, [ki], vf)
    _assert_typed_equals_string(o, [ki], vi)
    _assert_typed_equals_string(o, [k32, kd], vf, null_alias=[-1, 40])
    _assert_typed_equals_string(o, [kb, kf], vf)
    _assert_typed_equals_string(o, [ki, k32, kd, kb], vi, null_alias=[-1, -1, 40, -1], nthreads=5)
    # all-NULL values, a single row, the i64 sentinels
    _assert_typed_equals_string(o, [o.Col(o.I64, [7, 7, 8])], o.Col(o.F64, [1.0, 2.0, 3.0], o.pack_bits([True, True, True])))
    _assert_typed_equals_string(o, [o.Col(o.I64, [1])], o.Col(o.I64, [np.iinfo(np.int64).max]))
    _assert_typed_equals_string(o, [o.Col(o.I64, [1, 1, 2, 2])], o.Col(o.I64, [np.iinfo(np.int64).max, 5, np.iinfo(np.int64).min, np.iinfo(np.int64).min]))


def test_typed_oracle_filter_and_synthetic_source(oracle):
    o = oracle
    rng = np.random.default_rng(22)
    n = 30_000
    k = rng.integers(0, 50, n)
    kn = rng.random(n) < 0.1
    v = rng.random(n) * 100
    vn = rng.random(n) < 0.1
    f = rng.random(n) < 0.8
    fn = rng.random(n) < 0.05
    keep = np.nonzero(f & ~fn)[0]
    fcol = o.Col(o.BOOL_BITS, o.pack_bits(f), o.pack_bits(fn), length=n)
    ops = [o.SUM, o.MEAN, o.MIN, o.MAX, o.COUNT, o.STD]
    for compat in (False, True):
        # the reference filters first (data_ops.rs:37-121; compat: NULLs of the kept rows become defaults, keys included)
        kk, vv = k[keep].copy(), v[keep].copy()
        knn, vnn = kn[keep], vn[keep]
        if compat:
            kk[knn] = 0
            vv[vnn] = 0.0
            want = o.groupby([o.Col(o.I64, kk)], [o.Col(o.F64, vv)], [(0, op) for op in ops])
        else:
            want = o.groupby([o.Col(o.I64, kk, o.pack_bits(knn))], [o.Col(o.F64, vv, o.pack_bits(vnn))], [(0, op) for op in ops])
        got = o.typed_groupby([o.Col(o.I64, k, o.pack_bits(kn))], o.Col(o.F64, v, o.pack_bits(vn)), filter=fcol, compat_nulls=compat, nthreads=4)
        assert got["n_groups"] == want["n_groups"]
        index = {kt[0]: g for g, kt in enumerate(want["key_strings"])}
        for g in range(got["n_groups"]):
            ks = "NULL" if got["keys"][0][1][g] else str(int(got["keys"][0][0][g].astype(np.int64)))
            j = index[ks]
            for a, name in enumerate(["sum", "mean", "min", "max", None, "std"]):
                x = got["group_rows"][g] if name is None else got[name][g]
                assert x == want["aggs"][a][j], (compat, ks, name)
    # generator-backed rows == the same rows materialised (BASELINE.json configs[1] shape, 5% NULLs)
    m = 200_000
    a = o.typed_groupby_synth(m, card=1000, null_per_million=50_000, nthreads=3)
    b = o.typed_groupby([o.Col(o.I64, o.synth_keys(m, card=1000))], o.Col(o.F64, o.synth_vals(m), o.synth_nulls(m)), nthreads=2)
    oa, ob = np.argsort(a["keys"][0][0]), np.argsort(b["keys"][0][0])
    for name in ("group_rows", "valid_n", "sum", "mean", "min", "max", "std", "var"):
        assert np.array_equal(a[name][oa], b[name][ob]), name


def test_typed_join_equals_string_join(oracle):
    o = oracle
    rng = np.random.default_rng(23)
    L = o.Col(o.I64, rng.integers(0, 400, 3000), o.pack_bits(rng.random(3000) < 0.05))
    R = o.Col(o.I64, rng.integers(0, 400, 2000), o.pack_bits(rng.random(2000) < 0.05))
    for how in (o.INNER, o.LEFT):
        wl, wr = o.join(L, R, how)
        t = o.typed_join(L, R, how, want_pairs=True, nthreads=3)
        assert np.array_equal(t["left"], wl) and np.array_equal(t["right"], wr)          # same pairs in the same order
        assert t["n"] == len(wl) and t["checksum"] == o.pair_checksum(wl, wr)
        assert t["sum_left"] == int(wl.sum()) and t["sum_right"] == int(wr[wr >= 0].sum()) and t["unmatched_left"] == int((wr < 0).sum())
    # generator-backed sides (configs[2] shape) == the same keys materialised
    nb, npr = 5_000, 60_000
    t = o.typed_join(left_synth=dict(n=npr, domain=2 * nb), right_synth=dict(n=nb, unique=True), how=o.LEFT)
    wl, wr = o.join(o.Col(o.I64, o.synth_join_keys(npr, domain=2 * nb)), o.Col(o.I64, o.synth_join_keys(nb, unique=True)), o.LEFT)
    assert t["n"] == len(wl) == npr and t["checksum"] == o.pair_checksum(wl, wr)


def test_exact_mode_measures_the_reference_rounding_error(oracle):
    # ORC_MODE_EXACT (80-bit accumulation) is what tests/_util.py uses to tell the reference's own rounding error from a
    # defect of the CUDA path.  On offset data (values ~1e10, spread ~1) the reference's sequential f64 sum leaves its mean
    # off by ~1e-5, and its two-pass Std is therefore only good to ~1e-9 RELATIVE (measured below) - a CUDA result that is
    # closer to the truth than that cannot also be within 1e-12 of the reference, which is why compare_groupby() allows
    # rtol * |w| + 2 * |w - exact| and records both errors.
    o = oracle
    rng = np.random.default_rng(3)
    n = 50_000
    g = rng.integers(0, 20, n)
    v = 1e9 * (g + 1) + rng.normal(0, 1.0, n)
    ops = [(0, o.STD), (0, o.VAR), (0, o.MEAN), (0, o.SUM)]
    a = o.groupby([o.Col(o.I64, g)], [o.Col(o.F64, v)], ops)
    b = o.groupby([o.Col(o.I64, g)], [o.Col(o.F64, v)], ops, mode=o.MODE_EXACT)
    assert np.allclose(a["aggs"][2], b["aggs"][2], rtol=1e-12, atol=0) and np.allclose(a["aggs"][3], b["aggs"][3], rtol=1e-12, atol=0)
    rel = np.abs(a["aggs"][0] / b["aggs"][0] - 1).max()
    assert rel < 1e-7, rel
    assert np.abs(b["aggs"][0] - 1.0).max() < 0.05
    # well-conditioned data: the two modes agree to the last digits
    w = rng.random(n) * 1000
    a = o.groupby([o.Col(o.I64, g)], [o.Col(o.F64, w)], ops)
    b = o.groupby([o.Col(o.I64, g)], [o.Col(o.F64, w)], ops, mode=o.MODE_EXACT)
    for i in range(4):
        assert np.allclose(a["aggs"][i], b["aggs"][i], rtol=1e-13, atol=0)


def test_compare_groupby_typed_helper_on_a_mock_result(oracle):
    # tests/_util.py: compare_groupby_typed (the checker of tests/test_gpu_scale.py) fed with a permuted copy of a typed
    # result standing in for the CUDA result: passes as is, fails when one sum moves by 1e-11 relative or one key changes
    import pandrs_b200._native as pbn
    from _util import compare_groupby_typed
    o = oracle
    rng = np.random.default_rng(5)
    n = 20_000
    k1 = rng.integers(-20, 20, n).astype(np.int32)
    k2 = rng.integers(0, 9, n) * -(10**15)
    v = rng.random(n) * 1000
    tg = o.typed_groupby([o.Col(o.I32, k1, o.pack_bits(rng.random(n) < 0.1)), o.Col(o.I64, k2)], o.Col(o.F64, v, o.pack_bits(rng.random(n) < 0.05)))
    perm = rng.permutation(tg["n_groups"])
    ops = [pbn.SUM, pbn.MEAN, pbn.MIN, pbn.MAX, pbn.COUNT, pbn.STD, pbn.VAR]

    class Mock:
        n_groups = tg["n_groups"]
        sums = tg["sum"][perm].copy()
        k0 = tg["keys"][0][0][perm].astype(np.int64).astype(np.int32)
        def key(self, k): return (self.k0 if k == 0 else tg["keys"][1][0][perm].view(np.int64)), tg["keys"][k][1][perm]
        def group_rows(self): return tg["group_rows"][perm]
        def valid_n(self, v): return tg["valid_n"][perm]
        def agg(self, a):
            return [self.sums, tg["mean"][perm], tg["min"][perm], tg["max"][perm], tg["group_rows"][perm].astype(np.float64), tg["std"][perm], tg["var"][perm]][a]
    m = Mock()
    compare_groupby_typed(pbn, m, tg, [pbn.I32, pbn.I64], ops)
    m.sums[3] *= 1 + 1e-11
    with pytest.raises(AssertionError):
        compare_groupby_typed(pbn, m, tg, [pbn.I32, pbn.I64], ops)
    m.sums[3] = tg["sum"][perm][3]
    m.k0 = m.k0.copy()
    m.k0[5] += 1000
    with pytest.raises(AssertionError):
        compare_groupby_typed(pbn, m, tg, [pbn.I32, pbn.I64], ops)
    import _util
    _util.MEASURED.clear()          # a mock is not a measurement of the CUDA path


# ---------------------------------------------------------------- par_groupby / Median / First / Last (pinned by restatement)
def test_par_groupby_labels_and_row_lists(oracle):
    # grouping.rs:124-331: label = parts joined with "_", NULL -> "NA"; rows ascending.  The fixtures of
    # tests/optimized_groupby_test.rs:6-31, 138-171 (the reference asserts only "not empty" there)
    pool = ["A", "B", "C"]
    g = oracle.par_groupby([oracle.Col(oracle.DICT_U32, np.array([0, 1, 0, 1, 2], np.uint32), pool=pool)])
    assert {k: list(v) for k, v in g.items()} == {"A": [0, 2], "B": [1, 3], "C": [4]}
    pool = ["X", "Y", "A", "B"]
    g = oracle.par_groupby([oracle.Col(oracle.DICT_U32, np.array([0, 0, 1, 1, 0, 1], np.uint32), pool=pool),
                            oracle.Col(oracle.DICT_U32, np.array([2, 3, 2, 3, 2, 3], np.uint32), pool=pool)])
    assert {k: list(v) for k, v in g.items()} == {"X_A": [0, 4], "X_B": [1], "Y_A": [2], "Y_B": [3, 5]}
    # NULL -> "NA"; a literal "NULL" string stays its own group; ("a_b", "c") and ("a", "b_c") collide
    pool = ["a_b", "c", "a", "b_c", "NULL"]
    k0 = oracle.Col(oracle.DICT_U32, np.array([0, 2, 4, 4], np.uint32), oracle.pack_bits([False, False, False, True]), pool=pool)
    k1 = oracle.Col(oracle.DICT_U32, np.array([1, 3, 1, 1], np.uint32), pool=pool)
    g = oracle.par_groupby([k0, k1])
    assert {k: list(v) for k, v in g.items()} == {"a_b_c": [0, 1], "NULL_c": [2], "NA_c": [3]}


def test_median_first_last_semantics(oracle):
    # aggregation.rs:585-624 (Int64), 703-742 (Float64): NULLs are skipped by Median, First / Last look at the row itself
    k = oracle.Col(oracle.I64, np.array([1, 1, 1, 1, 2, 2, 3], np.int64))
    v = oracle.Col(oracle.I64, np.array([7, 1, 5, 3, 10, 4, 9], np.int64), oracle.pack_bits([False, False, False, True, True, False, True]))
    f = oracle.Col(oracle.F64, np.array([0.5, 2.5, 1.5, 9.0, 4.0, 8.0, 1.0]))
    r = oracle.groupby([k], [v, f], [(0, oracle.MEDIAN), (0, oracle.FIRST), (0, oracle.LAST), (1, oracle.MEDIAN), (1, oracle.FIRST), (1, oracle.LAST)])
    assert r["error"] == 0 and r["key_strings"] == [("1",), ("2",), ("3",)]
    assert [list(a) for a in r["aggs"]] == [[5.0, 4.0, 0.0], [7.0, 0.0, 0.0], [0.0, 4.0, 0.0], [2.0, 6.0, 1.0], [0.5, 4.0, 1.0], [9.0, 8.0, 1.0]]
    # even count: (values[mid - 1] + values[mid]) as f64 / 2.0 with the i64 sum computed first
    big = np.iinfo(np.int64).max
    r = oracle.groupby([oracle.Col(oracle.I64, np.zeros(2, np.int64))], [oracle.Col(oracle.I64, np.array([big, big], np.int64))], [(0, oracle.MEDIAN)])
    assert r["aggs"][0][0] == -1.0          # wrapping add (release build), then / 2.0
    s = oracle.Col(oracle.DICT_U32, np.zeros(2, np.uint32))
    assert oracle.groupby([oracle.Col(oracle.I64, np.zeros(2, np.int64))], [s], [(0, oracle.MEDIAN)])["error"] == 1     # aggregation.rs:748-752


def test_legacy_dataframe_groupby_restatement(oracle):
    # src/dataframe/groupby.rs:443-532 on the value fixture of src/dataframe/pandas_compat/groupby.rs:480-510 (A: 10, 30, 50; B: 20, 40):
    # sums 90 / 60, means 30 / 30, std 20 / 14.14..., Count = parseable cells, results formatted like f64::to_string()
    cols = {"category": ["A", "B", "A", "B", "A"], "value": ["10", "20", "30", "40", "50"]}
    r = oracle.legacy_groupby(cols, ["category"], [("value", f, f) for f in ("sum", "mean", "min", "max", "count", "std", "var", "median")])
    assert r[("A",)] == {"sum": "90", "mean": "30", "min": "10", "max": "50", "count": "3", "std": "20", "var": "400", "median": "30"}
    assert r[("B",)] == {"sum": "60", "mean": "30", "min": "20", "max": "40", "count": "2", "std": "14.142135623730951", "var": "200", "median": "30"}
    # unparseable cells are skipped; a group without a parseable cell gives 0; the parse grammar is Rust's, not Python's
    r = oracle.legacy_groupby({"k": ["a", "a", "a", "b"], "v": [" 4", "1_0", "+.5", "x"]}, ["k"], [("v", "sum", "s"), ("v", "count", "c")])
    assert r == {("a",): {"s": "0.5", "c": "1"}, ("b",): {"s": "0", "c": "0"}}
