"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Needs a B200: `pytest -m gpu`.

Cases follow the reference's own tests for this path plus the edge cases of §9:
empty and ragged inputs, NULL keys / values, sentinel collapse, duplicate-key fan-out, filters."""
import numpy as np
import pytest

import pandrs_b200 as pb
from _util import Spec, compare_groupby, compare_join, gpu_groupby_dict

pytestmark = pytest.mark.gpu

ALL6 = [pb.SUM, pb.MEAN, pb.MIN, pb.MAX, pb.COUNT, pb.STD]
A = "A B A B A C B C C A".split()
VALUE = [10, 25, 15, 30, 22, 18, 24, 12, 16, 20]
FLOAT = [1.1, 2.2, 3.3, 4.4, 5.5, 6.6, 7.7, 8.8, 9.9, 10.0]


def _dict(strings):
    pool, ids = [], []
    for s in strings:
        if s not in pool:
            pool.append(s)
        ids.append(pool.index(s))
    return ids, pool


# ---------------------------------------------------------------- reference known-answer vectors
def test_pandas_compat_groupby_goldens(ctx, oracle):
    # src/dataframe/pandas_compat/groupby.rs:480-617
    ids, pool = _dict(["A", "B", "A", "B", "A"])
    key = Spec(pb.DICT_U32, ids, pool=pool)
    val = Spec(pb.F64, [10.0, 20.0, 30.0, 40.0, 50.0])
    got = compare_groupby(pb, oracle, ctx, [key], [val], [(0, op) for op in ALL6])
    assert got[("A",)] == (3, [90.0, 30.0, 10.0, 50.0, 3.0, 20.0])
    assert got[("B",)][1][:5] == [60.0, 30.0, 20.0, 40.0, 2.0]


def test_ten_row_fixture(ctx, oracle):
    # tests/optimized_groupby_enhanced_test.rs:12-23; total 192 (optimized_custom_aggregation_test.rs:49)
    ids, pool = _dict(A)
    key = Spec(pb.DICT_U32, ids, pool=pool)
    ops = [pb.COUNT, pb.SUM, pb.MEAN, pb.MIN, pb.MAX, pb.STD]
    got = compare_groupby(pb, oracle, ctx, [key], [Spec(pb.I64, VALUE), Spec(pb.F64, FLOAT)], [(0, op) for op in ops] + [(1, op) for op in ops])
    want_i = {"A": [4, 67, 16.75, 10, 22, 5.377421934967226], "B": [3, 79, 26.333333333333332, 24, 30, 3.2145502536643185],
              "C": [3, 46, 15.333333333333334, 12, 18, 3.055050463303893]}
    want_f = {"A": [4, 19.9, 4.975, 1.1, 10.0, 3.8012059489939065], "B": [3, 14.3, 4.766666666666667, 2.2, 7.7, 2.768272626265942],
              "C": [3, 25.3, 8.433333333333334, 6.6, 9.9, 1.680277754817142]}
    for k in "ABC":
        vals = got[(k,)][1]
        assert vals[:6] == pytest.approx(want_i[k], rel=1e-14)
        assert vals[6:] == pytest.approx(want_f[k], rel=1e-14)
    assert sum(got[(k,)][1][1] for k in "ABC") == 192


def test_concurrency_fixture_four_groups(ctx, oracle):
    # tests/concurrency_test.rs:351-384 — 1,000 rows, category i % 4 -> 4 groups of 250
    got = compare_groupby(pb, oracle, ctx, [Spec(pb.I64, np.arange(1000) % 4)], [Spec(pb.F64, np.arange(1000, dtype=np.float64))], [(0, pb.COUNT)])
    assert sorted(v[0] for v in got.values()) == [250] * 4


def test_join_goldens(ctx, oracle):
    # tests/optimized_join_test.rs:6-237
    L, R = Spec(pb.I64, [1, 2, 3, 4]), Spec(pb.I64, [1, 2, 5, 6])
    li, ri = compare_join(pb, oracle, ctx, L, R, pb.INNER)
    assert list(zip(li, ri)) == [(0, 0), (1, 1)]
    li, ri = compare_join(pb, oracle, ctx, L, R, pb.LEFT)
    assert list(zip(li, ri)) == [(0, 0), (1, 1), (2, -1), (3, -1)]
    li, ri = compare_join(pb, oracle, ctx, L, Spec(pb.I64, [5, 6, 7, 8]), pb.INNER)
    assert len(li) == 0
    # src/dataframe/pandas_compat/merge.rs:320-410 — keys A,B,C,D / B,C,D,E, left order kept
    pool = list("ABCDE")
    li, ri = compare_join(pb, oracle, ctx, Spec(pb.DICT_U32, [0, 1, 2, 3], pool=pool), Spec(pb.DICT_U32, [1, 2, 3, 4], pool=pool), pb.LEFT)
    assert list(li) == [0, 1, 2, 3] and list(ri) == [-1, 0, 1, 2]


def test_join_type_mismatch_and_unsupported(ctx):
    # join.rs:98-104 -> Error::ColumnTypeMismatch
    with pytest.raises(pb.PandrsError) as e:
        ctx.join_pairs(pb.Column.int64([1]), pb.Column.float64([1.0]))
    assert e.value.kind == "ColumnTypeMismatch"
    with pytest.raises(pb.PandrsError) as e:
        ctx.join_pairs(pb.Column.int64([1]), pb.Column.int64([1]), how=7)
    assert e.value.kind == "InvalidInput"
    # aggregation.rs:748-752 -> Error::OperationFailed for Sum on a string column; Count is fine
    with pytest.raises(pb.PandrsError) as e:
        ctx.groupby_agg([pb.Column.int64([1, 2])], [pb.Column.dict_ids([0, 1])], [(0, pb.SUM)])
    assert e.value.kind == "OperationFailed"
    r = ctx.groupby_agg([pb.Column.int64([1, 1])], [pb.Column.dict_ids([0, 1])], [(0, pb.COUNT)])
    assert r.n_groups == 1 and r.agg(0)[0] == 2.0
    r.close()


# ---------------------------------------------------------------- NULL semantics / sentinels
def test_null_semantics(ctx, oracle):
    key = Spec(pb.I64, [7, 7, 0, 0, 9], nulls=[0, 0, 1, 1, 0])
    val = Spec(pb.F64, [1.0, 2.0, 3.0, 4.0, 5.0], nulls=[0, 1, 0, 0, 1])
    got = compare_groupby(pb, oracle, ctx, [key], [val], [(0, op) for op in ALL6])
    assert got[("7",)][1] == [1.0, 1.0, 1.0, 1.0, 2.0, 0.0]
    assert got[("9",)][1] == [0.0, 0.0, 0.0, 0.0, 1.0, 0.0]
    assert got[("NULL",)][1][0] == 7.0 and got[("NULL",)][1][4] == 2.0
    li, ri = compare_join(pb, oracle, ctx, Spec(pb.I64, [1, 2, 3], nulls=[0, 1, 0]), Spec(pb.I64, [1, 2, 3], nulls=[0, 0, 1]), pb.LEFT)
    assert list(zip(li, ri)) == [(0, 0), (2, -1)]


def test_short_null_mask(ctx, oracle):
    # int64_column.rs:77: bytes past the end of the mask read as "not NULL"
    k = pb.Column(pb.I64, np.arange(20) % 2, nulls=np.array([0xFF], np.uint8))
    r = ctx.groupby_agg([k], [pb.Column.float64(np.ones(20))], [(0, pb.COUNT)])
    keys, isnull = r.key(0)
    cnt = r.agg(0)
    got = {("NULL" if isnull[g] else int(keys[g])): cnt[g] for g in range(r.n_groups)}
    r.close()
    assert got == {"NULL": 8.0, 0: 6.0, 1: 6.0}


def test_int_sentinels_and_wrapping(ctx, oracle):
    key = Spec(pb.I64, [1, 1, 2, 2])
    val = Spec(pb.I64, [2**63 - 1, 1, -(2**63), -(2**63)])
    got = compare_groupby(pb, oracle, ctx, [key], [val], [(0, pb.SUM), (0, pb.MIN), (0, pb.MAX), (0, pb.MEAN), (0, pb.STD)])
    assert got[("1",)][1][0] == float(-(2**63)) and got[("2",)][1][2] == 0.0


def test_float_specials(ctx, oracle):
    inf, nan = np.inf, np.nan
    key = Spec(pb.I64, [1, 1, 1, 2, 2, 3, 3, 4, 5, 5])
    val = Spec(pb.F64, [1.0, nan, 3.0, inf, 2.0, -inf, -inf, nan, -0.0, 0.0])
    compare_groupby(pb, oracle, ctx, [key], [val], [(0, op) for op in ALL6])


def test_dict_null_alias_and_bool_f64_keys(ctx, oracle):
    # a string literally "NULL" merges with the NULL group (grouping.rs:69-98)
    pool = ["x", "NULL", "y"]
    key = Spec(pb.DICT_U32, [0, 1, 2, 1, 0, 2], nulls=[0, 0, 0, 0, 1, 0], pool=pool, null_alias=1)
    val = Spec(pb.F64, np.arange(6, dtype=np.float64))
    got = compare_groupby(pb, oracle, ctx, [key], [val], [(0, pb.SUM), (0, pb.COUNT)])
    assert got[("NULL",)] == (3, [8.0, 3.0])
    rng = np.random.default_rng(5)
    kb = Spec(pb.BOOL_BITS, rng.integers(0, 2, 500), nulls=rng.random(500) < 0.1)
    kf = Spec(pb.F64, rng.choice([0.0, -0.0, 1.5, np.nan, 2.0, np.inf], 500))
    v = Spec(pb.F64, rng.random(500))
    compare_groupby(pb, oracle, ctx, [kb], [v], [(0, op) for op in ALL6])
    compare_groupby(pb, oracle, ctx, [kf], [v], [(0, op) for op in ALL6])
    compare_groupby(pb, oracle, ctx, [kb, kf], [v], [(0, op) for op in ALL6])


# ---------------------------------------------------------------- randomized parity, both kernels
def _synth(oracle, n, card, nulls=True, seed=42, scramble=False):
    k = oracle.synth_keys(n, seed=seed, card=card, scramble=scramble)
    v = oracle.synth_vals(n, seed=seed)
    vn = None
    if nulls:
        vn = np.unpackbits(oracle.synth_nulls(n, seed=seed), bitorder="little")[:n].astype(bool)
    return Spec(pb.I64, k), Spec(pb.F64, v, vn)


def test_config1_shape(ctx, oracle):
    # BASELINE.json configs[0]: 1M rows, i64 key (1K distinct), sum/mean/max of f64
    k, v = _synth(oracle, 1_000_000, 1000, nulls=False)
    got = compare_groupby(pb, oracle, ctx, [k], [v], [(0, pb.SUM), (0, pb.MEAN), (0, pb.MAX)], device=True)
    assert len(got) == 1000
    assert ctx.stats()["groupby_algo_used"] in (pb.GB_SHARED, pb.GB_TILESORT)


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 511, 512, 513, 1025, 100_003])
def test_ragged_sizes(ctx, oracle, n):
    k, v = _synth(oracle, n, 17)
    compare_groupby(pb, oracle, ctx, [k], [v], [(0, op) for op in ALL6])
    compare_groupby(pb, oracle, ctx, [k], [v], [(0, op) for op in ALL6], device=True)


def test_empty(ctx):
    r = ctx.groupby_agg([pb.Column.int64([])], [pb.Column.float64([])], [(0, pb.SUM)])
    assert r.n_groups == 0 and len(r.agg(0)) == 0
    r.close()
    j = ctx.join_pairs(pb.Column.int64([]), pb.Column.int64([1, 2]))
    assert j.n == 0
    j.close()
    j = ctx.join_pairs(pb.Column.int64([1, 2]), pb.Column.int64([]), how=pb.LEFT)
    assert list(zip(*j.indices())) == [(0, -1), (1, -1)]
    j.close()


@pytest.mark.parametrize("card", [1, 3, 100, 1000, 3000, 20000, 150000])
@pytest.mark.parametrize("algo", [pb.GB_AUTO, pb.GB_GLOBAL])
def test_cardinalities_all_aggs(ctx, oracle, card, algo):
    # BASELINE.json configs[1] at oracle-sized n: six aggregates, 5% NULL values
    n = 200_000
    k, v = _synth(oracle, n, card, scramble=card > 100)
    ctx.set_option("groupby_algo", algo)
    try:
        got = compare_groupby(pb, oracle, ctx, [k], [v], [(0, op) for op in ALL6], device=True)
    finally:
        ctx.set_option("groupby_algo", pb.GB_AUTO)
    assert len(got) == len(np.unique(k.values))


@pytest.mark.parametrize("dense", [1, 0])
def test_skewed_keys_take_the_warp_cooperative_reduce(ctx, oracle, dense):
    # Zipf-like skew (BASELINE.json configs[3] style) and a group that owns 90% of the rows: segments of the sorted
    # tile longer than 64 rows are reduced by the owner's whole warp (gb_tsort.cu); int and float values, NULLs
    rng = np.random.default_rng(5)
    n = 300_000
    w = 1.0 / np.arange(1, 401) ** 1.1
    zipf = rng.choice(400, size=n, p=w / w.sum())
    heavy = np.where(rng.random(n) < 0.9, 7, rng.integers(0, 40, n))
    for keys in (zipf, heavy):
        kv = keys.astype(np.int64) * (1 if dense else 1_000_003) - (0 if dense else 17)
        k = Spec(pb.I64, kv)
        v = Spec(pb.F64, 50.0 + rng.normal(0, 5.0, n), nulls=rng.random(n) < 0.05)
        vi = Spec(pb.I64, rng.integers(-10**6, 10**6, n))
        ctx.set_option("dense", dense)
        try:
            got = compare_groupby(pb, oracle, ctx, [k], [v, vi], [(0, op) for op in ALL6] + [(1, op) for op in ALL6], device=True)
        finally:
            ctx.set_option("dense", 1)
        assert ctx.stats()["groupby_algo_used"] == pb.GB_TILESORT and len(got) == len(np.unique(kv))


def test_partitioned_high_cardinality_path(ctx, oracle):
    # thousands of groups, >= 2 tiles per partition: one- and two-level hash partition + tile sort per partition
    # (gb_part.cu); NULL values travel as flag bytes, a row filter is applied by the first partition pass
    rng = np.random.default_rng(6)
    for n, card in ((3_000_000, 60_000), (3_000_000, 5_000), (9_000_000, 200_000)):     # the last one needs two levels
        kv = rng.integers(0, card, n) * 7_919 - 1_000_000
        k = Spec(pb.I64, kv)
        v = Spec(pb.F64, rng.normal(10.0, 3.0, n), nulls=rng.random(n) < 0.05)
        f = Spec(pb.BOOL_BITS, rng.random(n) < 0.9)
        got = compare_groupby(pb, oracle, ctx, [k], [v], [(0, op) for op in ALL6], device=True)
        assert ctx.stats()["groupby_algo_used"] == pb.GB_PARTITIONED and len(got) == len(np.unique(kv))
        got = compare_groupby(pb, oracle, ctx, [k], [v], [(0, op) for op in ALL6], filter_spec=f, device=True)
        assert ctx.stats()["groupby_algo_used"] == pb.GB_PARTITIONED
        ctx.set_option("part", 0)
        try:
            compare_groupby(pb, oracle, ctx, [k], [v], [(0, pb.SUM), (0, pb.COUNT)], device=True)
            assert ctx.stats()["groupby_algo_used"] != pb.GB_PARTITIONED
        finally:
            ctx.set_option("part", 1)


@pytest.mark.parametrize("part_hash", [1, 2])
def test_partitioned_groups_written_straight_to_the_result(ctx, oracle, part_hash):
    # gb_part.cu: complete hash partitions are aggregated in a shared-memory table and their groups are written to the result
    # directly (part_hash = 1); incomplete partitions (a hot key parked part of them in the side area) and part_hash = 2 go
    # through the tile-sort kernel and the global table.  Few rows per group, the all-ones key word, Int64 values, sum-only,
    # a moderately hot key, partial states (the multi-GPU path merges them).
    rng = np.random.default_rng(66)
    n = 2_500_000
    kv = rng.integers(0, 300_000, n) * 1_000_003 - 7
    kv[::1001] = -1                                     # the key whose bits are all ones has a reserved table slot
    hot = rng.random(n) < 0.08
    kv[hot] = 424_242_424_242                           # its partition overflows into the side area: incomplete -> table
    k = Spec(pb.I64, kv)
    v = Spec(pb.F64, rng.normal(1e6, 3.0, n), nulls=rng.random(n) < 0.05)
    vi = Spec(pb.I64, rng.integers(-2**61, 2**61, n), nulls=rng.random(n) < 0.05)
    ctx.set_option("part_hash", part_hash)
    try:
        got = compare_groupby(pb, oracle, ctx, [k], [v], [(0, op) for op in ALL6] + [(0, pb.VAR)], device=True)
        assert ctx.stats()["groupby_algo_used"] == pb.GB_PARTITIONED and len(got) == len(np.unique(kv))
        compare_groupby(pb, oracle, ctx, [k], [vi], [(0, op) for op in ALL6], device=True)
        assert ctx.stats()["groupby_algo_used"] == pb.GB_PARTITIONED
        compare_groupby(pb, oracle, ctx, [k], [v], [(0, pb.SUM), (0, pb.MEAN), (0, pb.COUNT)], device=True)
        k2 = Spec(pb.I32, rng.integers(0, 70, n).astype(np.int32))
        k3 = Spec(pb.DICT_U32, rng.integers(0, 900, n).astype(np.uint32), pool=[f"s{i}" for i in range(900)])
        compare_groupby(pb, oracle, ctx, [k2, k3], [v], [(0, op) for op in ALL6], device=True)          # packed two-key tuples, ~40 rows per group
        assert ctx.stats()["groupby_algo_used"] == pb.GB_PARTITIONED
        # partial states of the same groups merge to the same result (pdrs_groupby_partial + pdrs_groupby_merge)
        kc, vc = ctx.upload(k.gpu(pb)), ctx.upload(v.gpu(pb))
        whole = ctx.groupby_agg([kc], [vc], [(0, op) for op in ALL6])
        part = ctx.groupby_partial([kc], [vc], all_stats=True)
        G = part.n_groups
        assert G == whole.n_groups
        kcol = pb.Column(pb.I64, device_ptr=part.key_dev(0), length=G)
        merged = ctx.groupby_merge([kcol], [part.states_dev(0)], [False], G, [(0, op) for op in ALL6])
        wk, mk = whole.key(0)[0], merged.key(0)[0]
        wo, mo = np.argsort(wk), np.argsort(mk)
        assert np.array_equal(wk[wo], mk[mo])
        for a in range(6):
            assert np.allclose(whole.agg(a)[wo], merged.agg(a)[mo], rtol=1e-12, atol=0)
        for r in (whole, part, merged):
            r.close()
        ctx.free(kc); ctx.free(vc)
    finally:
        ctx.set_option("part_hash", 0)


def test_partition_overflow_falls_back(ctx, oracle):
    # heavily skewed high-cardinality keys: one key owns half of the rows, so its hash bucket overflows the padded
    # range of the one-pass partition; the call must restart on the global-table path and still match the oracle
    rng = np.random.default_rng(8)
    n = 3_000_000
    kv = np.where(rng.random(n) < 0.5, 123_456_789, rng.integers(0, 40_000, n) * 104_729)
    k = Spec(pb.I64, kv)
    v = Spec(pb.F64, rng.normal(1.0, 1.0, n), nulls=rng.random(n) < 0.05)
    ctx.set_option("part_hot", 0)            # without the hot-key routing
    try:
        got = compare_groupby(pb, oracle, ctx, [k], [v], [(0, op) for op in ALL6], device=True)
        assert len(got) == len(np.unique(kv)) and ctx.stats()["groupby_algo_used"] != pb.GB_PARTITIONED
    finally:
        ctx.set_option("part_hot", 1)
    # default: the cardinality sample has seen the hot key; its rows bypass the hash partitions (side area, one tile-sort pass),
    # nothing overflows and the partitioned path applies
    got = compare_groupby(pb, oracle, ctx, [k], [v], [(0, op) for op in ALL6], device=True)
    assert len(got) == len(np.unique(kv)) and ctx.stats()["groupby_algo_used"] == pb.GB_PARTITIONED
    # a milder hot key (12% of the rows): its bucket fills up and the excess runs are parked in the side area behind
    # the buckets, which is aggregated as extra partitions - the partitioned path still applies
    kv = np.where(rng.random(n) < 0.12, 123_456_789, rng.integers(0, 40_000, n) * 104_729)
    got = compare_groupby(pb, oracle, ctx, [Spec(pb.I64, kv)], [v], [(0, op) for op in ALL6], device=True)
    assert len(got) == len(np.unique(kv)) and ctx.stats()["groupby_algo_used"] == pb.GB_PARTITIONED


@pytest.mark.parametrize("radix", [1, 0])
def test_high_cardinality_radix_path(ctx, oracle, radix):
    # BASELINE.json configs[1] "10M distinct" shape at oracle-sized n: the table (slots x 80 B) no longer fits L2, so
    # rows are radix-partitioned first (gb_radix.cu); radix=0 forces the plain global-table kernel on the same data
    n = 2_500_000
    k, v = _synth(oracle, n, 700_000, scramble=True)
    kn = np.zeros(n, bool)
    kn[::1009] = True
    k = Spec(pb.I64, k.values, nulls=kn)
    ctx.set_option("radix", radix)
    try:
        got = compare_groupby(pb, oracle, ctx, [k], [v], [(0, op) for op in ALL6], device=True)
        st = ctx.stats()
    finally:
        ctx.set_option("radix", 1)
    assert st["groupby_algo_used"] == pb.GB_GLOBAL and len(got) > 600_000


@pytest.mark.parametrize("ops", [[pb.SUM], [pb.SUM, pb.MEAN, pb.COUNT], [pb.MIN, pb.MAX], [pb.STD], [pb.VAR, pb.MEAN], [pb.COUNT]])
def test_agg_subsets_and_int_values(ctx, oracle, ops):
    n = 50_000
    k, v = _synth(oracle, n, 257)
    rng = np.random.default_rng(1)
    vi = Spec(pb.I64, rng.integers(-10**12, 10**12, n), nulls=rng.random(n) < 0.05)
    compare_groupby(pb, oracle, ctx, [k], [v, vi], [(0, op) for op in ops] + [(1, op) for op in ops])


def test_key_nulls_and_spill(ctx, oracle):
    n = 120_000
    rng = np.random.default_rng(2)
    k = Spec(pb.I64, rng.integers(-50, 50, n), nulls=rng.random(n) < 0.03)
    v = Spec(pb.F64, rng.normal(0, 1e6, n), nulls=rng.random(n) < 0.05)
    compare_groupby(pb, oracle, ctx, [k], [v], [(0, op) for op in ALL6])
    # under-estimated cardinality: the CTA tables overflow and rows spill to the global table
    k2 = Spec(pb.I64, rng.integers(0, 1500, n))
    ctx.set_option("groups_hint", 64)
    try:
        got = compare_groupby(pb, oracle, ctx, [k2], [v], [(0, op) for op in ALL6])
        st = ctx.stats()
    finally:
        ctx.set_option("groups_hint", 0)
    assert len(got) == 1500 and st["spilled_rows"] > 0


@pytest.mark.parametrize("dense", [1, 0])
def test_dense_and_hashed_key_paths(ctx, oracle, dense):
    # small dense integer keys take the direct-mapped path (no key table); keys outside the sampled range
    # (here: rows the sampler never visits) must still be grouped correctly through the spill path
    n = 600_000
    rng = np.random.default_rng(12)
    k = rng.integers(-500, 500, n)
    out = np.arange(400, n, 585 * 40)          # r % 585 == 400: outside every sampled run of 256 rows
    k[out] = 10**15 + (np.arange(len(out)) % 3)
    v = Spec(pb.F64, rng.random(n) * 1000, nulls=rng.random(n) < 0.05)
    vi = Spec(pb.I64, rng.integers(-2**62, 2**62, n))
    ctx.set_option("dense", dense)
    try:
        got = compare_groupby(pb, oracle, ctx, [Spec(pb.I64, k)], [v, vi], [(0, op) for op in ALL6] + [(1, op) for op in ALL6], device=True)
        st = ctx.stats()
    finally:
        ctx.set_option("dense", 1)
    assert len(got) == 1003 and st["groupby_algo_used"] in (pb.GB_SHARED, pb.GB_TILESORT)
    if dense:   # two value columns = two passes, every outlier row spills in each of them
        assert st["spilled_rows"] == 2 * len(out)


def test_variance_is_stable_for_offset_groups(ctx, oracle):
    # group means far apart and far from zero: a single global pivot would lose all digits
    n = 100_000
    rng = np.random.default_rng(3)
    g = rng.integers(0, 50, n)
    v = 1e9 * (g + 1) + rng.normal(0, 1.0, n)
    compare_groupby(pb, oracle, ctx, [Spec(pb.I64, g)], [Spec(pb.F64, v)], [(0, pb.STD), (0, pb.VAR), (0, pb.MEAN), (0, pb.SUM)])


def test_multi_key_and_dictionary(ctx, oracle):
    # BASELINE.json configs[3] shape: (i32, i64) + dictionary-encoded string key, skewed keys
    n = 150_000
    rng = np.random.default_rng(4)
    z = lambda a, dom: np.minimum(rng.zipf(a, n) - 1, dom - 1)
    k1 = Spec(pb.I32, z(1.1, 1000).astype(np.int32) - 500)
    k2 = Spec(pb.I64, z(1.1, 100000).astype(np.int64) * 7_000_000_007)
    pool = [f"s{i}" for i in range(10000)]
    k3 = Spec(pb.DICT_U32, z(1.1, 10000).astype(np.uint32), pool=pool)
    v = Spec(pb.F64, rng.random(n) * 1000, nulls=rng.random(n) < 0.05)
    compare_groupby(pb, oracle, ctx, [k1, k2, k3], [v], [(0, op) for op in ALL6], device=True)
    compare_groupby(pb, oracle, ctx, [k1, k3], [v], [(0, pb.SUM), (0, pb.COUNT)])
    # with NULL key parts: every distinct NULL pattern is its own group
    k1n = Spec(pb.I32, k1.values, nulls=rng.random(n) < 0.2)
    k3n = Spec(pb.DICT_U32, k3.values % 5, nulls=rng.random(n) < 0.2, pool=pool)
    k2s = Spec(pb.I64, k2.values % 3, nulls=rng.random(n) < 0.2)
    compare_groupby(pb, oracle, ctx, [k1n, k2s, k3n], [v], [(0, op) for op in ALL6])


def test_packed_keys_through_the_tile_sort_kernel(ctx, oracle):
    # string (dictionary) keys, i32 keys and key pairs that pack into one 64-bit word take the tile-sort kernel
    # through load_key_generic; NULL keys, a literal "NULL" string, a row filter, int and float values
    n = 180_000
    rng = np.random.default_rng(12)
    pool = [f"name{i}" for i in range(300)] + ["NULL"]
    ids = rng.integers(0, 301, n).astype(np.uint32)
    ks = Spec(pb.DICT_U32, ids, nulls=rng.random(n) < 0.03, pool=pool, null_alias=300)
    k32 = Spec(pb.I32, rng.integers(-40, 40, n).astype(np.int32), nulls=rng.random(n) < 0.1)
    kb = Spec(pb.BOOL_BITS, rng.random(n) < 0.5)
    v = Spec(pb.F64, rng.normal(5.0, 2.0, n), nulls=rng.random(n) < 0.05)
    vi = Spec(pb.I64, rng.integers(-1000, 1000, n))
    f = Spec(pb.BOOL_BITS, rng.random(n) < 0.8, nulls=rng.random(n) < 0.05)
    aggs = [(0, op) for op in ALL6] + [(1, op) for op in ALL6]
    k32d = Spec(pb.I32, rng.integers(-40, 40, n).astype(np.int32))                       # 32 + 32 bits: exactly one word
    ksd = Spec(pb.DICT_U32, rng.integers(0, 20, n).astype(np.uint32), pool=pool)
    for keys in ([ks], [k32], [k32d, ksd], [kb, k32]):
        compare_groupby(pb, oracle, ctx, keys, [v, vi], aggs, device=True)
        assert ctx.stats()["groupby_algo_used"] == pb.GB_TILESORT, [k.dtype for k in keys]
    compare_groupby(pb, oracle, ctx, [kb, k32], [v, vi], aggs, filter_spec=f, device=True)
    assert ctx.stats()["groupby_algo_used"] == pb.GB_TILESORT
    ctx.set_option("compat_filter_nulls", 1)
    try:
        compare_groupby(pb, oracle, ctx, [ks], [v], [(0, pb.SUM), (0, pb.STD)], filter_spec=f, device=True, compat_nulls=True)
    finally:
        ctx.set_option("compat_filter_nulls", 0)


@pytest.mark.parametrize("compat", [False, True])
def test_fused_filter(ctx, oracle, compat):
    # BASELINE.json configs[4] shape: Boolean-mask filter -> groupby(returnflag, linestatus)
    n = 80_000
    rng = np.random.default_rng(6)
    rf = Spec(pb.DICT_U32, rng.integers(0, 3, n).astype(np.uint32), pool=["A", "N", "R"])
    ls = Spec(pb.DICT_U32, rng.integers(0, 2, n).astype(np.uint32), pool=["F", "O"])
    qty = Spec(pb.F64, rng.integers(1, 51, n).astype(np.float64), nulls=rng.random(n) < 0.02)
    price = Spec(pb.F64, rng.random(n) * 1e5)
    mask = Spec(pb.BOOL_BITS, rng.random(n) < 0.98, nulls=rng.random(n) < 0.01)
    aggs = [(0, pb.SUM), (1, pb.SUM), (0, pb.MEAN), (1, pb.MEAN), (0, pb.COUNT), (0, pb.MIN), (1, pb.STD)]
    ctx.set_option("compat_filter_nulls", int(compat))
    try:
        got = compare_groupby(pb, oracle, ctx, [rf, ls], [qty, price], aggs, filter_spec=mask, compat_nulls=compat)
        got2 = compare_groupby(pb, oracle, ctx, [rf, ls], [qty, price], aggs, filter_spec=mask, compat_nulls=compat, device=True)
    finally:
        ctx.set_option("compat_filter_nulls", 0)
    assert len(got) == 6 and got.keys() == got2.keys()


def test_few_groups_many_value_columns_in_one_scan(ctx, oracle):
    # BASELINE.json configs[4] shape: Boolean-mask filter -> groupby(returnflag, linestatus) -> sums / means / count of FIVE value
    # columns: the few-groups kernel (gb_few.cu) reads keys, mask and all value columns once.  Every key layout (one Int64 key,
    # raw 4-byte columns, generic tuples), int and float values, NULL values, filter NULLs, compat_filter_nulls, a ragged tail,
    # keys the cardinality sample never saw (they spill to the global table), and a typed predicate evaluated in the scan.
    n = 300_003
    rng = np.random.default_rng(91)
    rf = Spec(pb.DICT_U32, rng.integers(0, 3, n).astype(np.uint32), pool=["A", "N", "R"])
    ls = Spec(pb.DICT_U32, rng.integers(0, 2, n).astype(np.uint32), pool=["F", "O"])
    qty = Spec(pb.F64, rng.integers(1, 51, n).astype(np.float64), nulls=rng.random(n) < 0.02)
    price = Spec(pb.F64, rng.random(n) * 1e5)
    disc = Spec(pb.F64, rng.random(n) * 0.1, nulls=rng.random(n) < 0.3)
    cnt_i = Spec(pb.I64, rng.integers(-2**60, 2**60, n), nulls=rng.random(n) < 0.05)
    tax = Spec(pb.F64, rng.random(n) * 0.08)
    ship = Spec(pb.I64, rng.integers(8000, 10600, n), nulls=rng.random(n) < 0.01)
    mask = Spec(pb.BOOL_BITS, rng.random(n) < 0.98, nulls=rng.random(n) < 0.01)
    vals = [qty, price, disc, cnt_i, tax]
    aggs = [(0, pb.SUM), (1, pb.SUM), (2, pb.SUM), (3, pb.SUM), (4, pb.SUM), (0, pb.MEAN), (1, pb.MEAN), (2, pb.MEAN), (3, pb.MEAN), (0, pb.COUNT)]
    k64 = Spec(pb.I64, rng.integers(-2, 3, n) * 10**15)
    kb = Spec(pb.BOOL_BITS, rng.random(n) < 0.5)
    k32 = Spec(pb.I32, rng.integers(-1, 2, n).astype(np.int32))
    for keys in ([rf, ls], [k64], [rf], [kb, k32]):
        for filt in (None, mask):
            got = compare_groupby(pb, oracle, ctx, keys, vals, aggs, filter_spec=filt, device=True)
            assert ctx.stats()["groupby_algo_used"] == pb.GB_FEW, ([k.dtype for k in keys], ctx.stats())
    assert len(got) == 6
    compare_groupby(pb, oracle, ctx, [rf, ls], vals, aggs, filter_spec=mask)                       # host columns
    compare_groupby(pb, oracle, ctx, [rf, ls], vals, aggs, device=True, pred=(ship, pb.CMP_LE, 10_500))
    assert ctx.stats()["groupby_algo_used"] == pb.GB_FEW
    compare_groupby(pb, oracle, ctx, [rf, ls], vals, aggs, filter_spec=mask, device=True, pred=(price, pb.CMP_GT, 5e4))
    ctx.set_option("compat_filter_nulls", 1)
    try:
        compare_groupby(pb, oracle, ctx, [rf, ls], vals, aggs, filter_spec=mask, compat_nulls=True, device=True)
        assert ctx.stats()["groupby_algo_used"] == pb.GB_FEW
        compare_groupby(pb, oracle, ctx, [k64], vals, aggs, compat_nulls=True, device=True, pred=(ship, pb.CMP_GE, 8100))
    finally:
        ctx.set_option("compat_filter_nulls", 0)
    # min / max / std need the full statistics: not this kernel; the predicate then becomes a mask in front of the other kernels
    compare_groupby(pb, oracle, ctx, [rf, ls], [qty, price], [(0, pb.SUM), (1, pb.STD), (0, pb.MIN)], device=True, pred=(ship, pb.CMP_LT, 9000))
    assert ctx.stats()["groupby_algo_used"] != pb.GB_FEW
    # keys the sample never saw: the sample reads runs of 256 rows every n // 1024 rows; rows at offset 270 of a stride lie between runs
    k = rng.integers(0, 4, n)
    stride = n // 1024
    assert stride > 280
    out = stride * np.arange(3, 1000, 37) + 270
    k[out] = 1000 + (np.arange(len(out)) % 3)
    got = compare_groupby(pb, oracle, ctx, [Spec(pb.I64, k)], vals, aggs, device=True)
    st = ctx.stats()
    assert len(got) == 7 and st["groupby_algo_used"] == pb.GB_FEW and st["spilled_rows"] == len(out)


def test_fused_filter_turns_null_keys_into_defaults(ctx, oracle):
    # compat_filter_nulls: the reference filters first, and its filter() defaults the NULLs of every column of the kept
    # rows - key columns included (data_ops.rs:64-108, parallel.rs:177-231): a NULL Int64 key joins group "0", a NULL f64
    # key group "0", a NULL bool key "false", a NULL string key the group of "" - never the "NULL" group
    n = 60_000
    rng = np.random.default_rng(61)
    ki = Spec(pb.I64, rng.integers(-3, 4, n), nulls=rng.random(n) < 0.2)
    kf = Spec(pb.F64, rng.integers(0, 3, n).astype(np.float64), nulls=rng.random(n) < 0.2)
    kb = Spec(pb.BOOL_BITS, rng.random(n) < 0.5, nulls=rng.random(n) < 0.2)
    pool = ["a", "b", "NULL", "c"]
    kd = Spec(pb.DICT_U32, rng.integers(0, 4, n).astype(np.uint32), nulls=rng.random(n) < 0.2, pool=pool, null_alias=2)
    k32 = Spec(pb.I32, rng.integers(5, 9, n).astype(np.int32), nulls=rng.random(n) < 0.2)
    v = Spec(pb.F64, rng.random(n) * 10, nulls=rng.random(n) < 0.1)
    f = Spec(pb.BOOL_BITS, rng.random(n) < 0.7, nulls=rng.random(n) < 0.05)
    aggs = [(0, op) for op in ALL6]
    ctx.set_option("compat_filter_nulls", 1)
    try:
        for keys in ([ki], [kf], [kb], [kd], [k32, kd], [ki, kb, kd]):
            got = compare_groupby(pb, oracle, ctx, keys, [v], aggs, filter_spec=f, compat_nulls=True, device=len(keys) == 1)
            if len(keys) == 1 and keys[0] is not kd:
                assert ("NULL",) not in got
        got = compare_groupby(pb, oracle, ctx, [kd], [v], aggs, filter_spec=f, compat_nulls=True)
        assert ("",) in got and ("NULL",) in got          # "" = the former NULLs, "NULL" = the literal string
    finally:
        ctx.set_option("compat_filter_nulls", 0)
    # the frame mirror: LazyFrame filter -> aggregate (fused) == filter() then group_by (unfused), NULL keys included
    from pandrs_b200 import frame as fr
    df = fr.OptimizedDataFrame()
    df.add_column("k", fr.StringColumn([pool[i] for i in kd.values[:5000]], nulls=kd.nulls[:5000]))
    df.add_column("i", fr.Int64Column(ki.values[:5000], nulls=ki.nulls[:5000]))
    df.add_column("v", fr.Float64Column(v.values[:5000], nulls=v.nulls[:5000]))
    df.add_column("keep", fr.BooleanColumn(f.values[:5000], nulls=f.nulls[:5000]))
    spec = [("v", fr.AggregateOp.Sum, "s"), ("v", fr.AggregateOp.Count, "c")]
    fused = fr.LazyFrame.new(df).filter("keep").aggregate(["k", "i"], spec).execute()
    plain = fr.LazyFrame.new(df.filter("keep")).aggregate(["k", "i"], spec).execute()
    def rows(d):
        lst = lambda name: [d.column(name).get(i) for i in range(len(d.column(name)))]
        return sorted(zip(lst("k"), lst("i"), lst("c"), np.round(lst("s"), 9).tolist()))
    assert rows(fused) == rows(plain) and any(r[0] == "" for r in rows(fused)) and any(r[1] == "0" for r in rows(fused))


# ---------------------------------------------------------------- joins
@pytest.mark.parametrize("how", [pb.INNER, pb.LEFT])
def test_join_random_duplicates_and_nulls(ctx, oracle, how):
    rng = np.random.default_rng(7)
    L = Spec(pb.I64, rng.integers(0, 400, 3000), nulls=rng.random(3000) < 0.05)
    R = Spec(pb.I64, rng.integers(0, 400, 2000), nulls=rng.random(2000) < 0.05)
    compare_join(pb, oracle, ctx, L, R, how)
    compare_join(pb, oracle, ctx, L, R, how, device=True)


@pytest.mark.parametrize("how", [pb.INNER, pb.LEFT])
def test_join_low_cardinality_build_keys(ctx, oracle, how):
    # A build key with k rows: its rows are collected into one CSR segment and sorted once (join.cu jdup_*), so the join
    # costs O(k log^2 k + output) - not O(k^2) per probe row.  Segment lengths here cover every sort path: <= 32 rows (one
    # thread), <= 8192 (bitonic in shared memory), 40 000 (bitonic in global scratch); matches must come out in ascending
    # right row, left-row-major (join.rs:150-163) - compared with the oracle in ORDER, not after a sort.
    rng = np.random.default_rng(71)
    rk = np.concatenate([np.repeat(np.arange(5), 40_000), np.repeat(np.arange(100, 140), 3000), np.repeat(np.arange(1000, 3000), 20), np.arange(10_000, 12_000)])
    rk = rk[rng.permutation(len(rk))]
    R = Spec(pb.I64, rk, nulls=rng.random(len(rk)) < 0.01)
    L = Spec(pb.I64, np.concatenate([rng.integers(0, 5, 20), rng.integers(100, 140, 50), rng.integers(1000, 3100, 400), rng.integers(9_000, 12_500, 2000)]),
             nulls=None)
    compare_join(pb, oracle, ctx, L, R, how)
    compare_join(pb, oracle, ctx, L, R, how, device=True)
    ctx.set_option("join_algo", 2)              # the radix-partitioned path: same multiset of pairs in bucket order
    try:
        compare_join(pb, oracle, ctx, L, R, how, device=True, check_order=False)
    finally:
        ctx.set_option("join_algo", 0)
    # the all-ones key (the table's empty marker lives in a reserved slot) with duplicates
    R2 = Spec(pb.I64, np.array([-1, 5, -1, 7, -1, 5] * 20))
    L2 = Spec(pb.I64, np.array([-1, 5, 9, -1]))
    compare_join(pb, oracle, ctx, L2, R2, how)


@pytest.mark.parametrize("how", [pb.RIGHT, pb.OUTER])
@pytest.mark.parametrize("size", ["small", "radix"])
def test_right_and_outer_joins(ctx, oracle, how, size):
    # join.rs:211-224: unmatched right rows are appended as (None, r), NULL-key right rows included
    rng = np.random.default_rng(11)
    if size == "small":
        nl, nr, dom = 5000, 3000, 2500
    else:
        nl, nr, dom = 400_000, 120_000, 200_000
    L = Spec(pb.I64, rng.integers(0, dom, nl), nulls=rng.random(nl) < 0.02)
    R = Spec(pb.I64, rng.integers(0, dom, nr), nulls=rng.random(nr) < 0.02)
    if size == "radix":
        ctx.set_option("join_algo", 2)
    try:
        li, ri = compare_join(pb, oracle, ctx, L, R, how, device=True, check_order=False)
    finally:
        ctx.set_option("join_algo", 0)
    assert (li < 0).sum() > 0 and ((ri < 0).sum() > 0) == (how == pb.OUTER)
    # unique build keys take the single-pass probe + emit on the radix path
    Ru = Spec(pb.I64, rng.permutation(dom)[:nr])
    if size == "radix":
        ctx.set_option("join_algo", 2)
    try:
        compare_join(pb, oracle, ctx, L, Ru, how, device=True, check_order=False)
    finally:
        ctx.set_option("join_algo", 0)


@pytest.mark.parametrize("how", [pb.INNER, pb.LEFT])
def test_join_unique_build_config3_shape(ctx, oracle, how):
    # BASELINE.json configs[2] at oracle-sized n: unique i64 build keys, ~50% hit rate
    nb, npr = 100_000, 1_000_000
    R = Spec(pb.I64, oracle.synth_join_keys(nb, unique=True))
    L = Spec(pb.I64, oracle.synth_join_keys(npr, domain=2 * nb))
    li, ri = compare_join(pb, oracle, ctx, L, R, how, device=True)
    if how == pb.INNER:
        assert 0.45 * npr < len(li) < 0.55 * npr
    else:
        assert len(li) == npr


@pytest.mark.parametrize("how", [pb.INNER, pb.LEFT])
def test_join_radix_partitioned_path(ctx, oracle, how):
    # large tables take the radix-partitioned path (L2-resident table regions); its pairs come out in bucket
    # order, so they are compared after the canonical sort (north star: "compared after a canonical sort")
    rng = np.random.default_rng(14)
    ctx.set_option("join_algo", 2)
    try:
        L = Spec(pb.I64, rng.integers(0, 5000, 40_000), nulls=rng.random(40_000) < 0.05)
        R = Spec(pb.I64, rng.integers(0, 5000, 9_000), nulls=rng.random(9_000) < 0.05)
        compare_join(pb, oracle, ctx, L, R, how, check_order=False)
        nb, npr = 300_000, 2_000_003
        Ru = Spec(pb.I64, oracle.synth_join_keys(nb, unique=True))
        Lu = Spec(pb.I64, oracle.synth_join_keys(npr, domain=2 * nb))
        compare_join(pb, oracle, ctx, Lu, Ru, how, device=True, check_order=False)
        assert ctx.stats()["groupby_algo_used"] == 2          # one table for all buckets
        ctx.set_option("join_bucketwise", 2)
        try:
            compare_join(pb, oracle, ctx, Lu, Ru, how, device=True, check_order=False)
            assert ctx.stats()["groupby_algo_used"] == 3      # bucket-at-a-time, one reused table region (default with payload columns)
            compare_join(pb, oracle, ctx, L, R, how, check_order=False)     # duplicate build keys: detected, falls back to the one-table path
            assert ctx.stats()["groupby_algo_used"] == 2
        finally:
            ctx.set_option("join_bucketwise", 1)
        pool = [f"k{i}" for i in range(300)]
        Ld = Spec(pb.DICT_U32, rng.integers(0, 300, 5000).astype(np.uint32), pool=pool)
        Rd = Spec(pb.DICT_U32, rng.integers(0, 300, 700).astype(np.uint32), pool=pool)
        compare_join(pb, oracle, ctx, Ld, Rd, how, check_order=False)
    finally:
        ctx.set_option("join_algo", 0)


def _compare_join_gather(ctx, oracle, L, R, how, cols, device=True):
    """pdrs_join_gather against oracle join + oracle gather (join.rs:290-552), pair by pair after the canonical sort."""
    lc, rc, cc = L.gpu(pb), R.gpu(pb), [c.gpu(pb) for c in cols]
    ups = []
    if device:
        lc, rc = ctx.upload(lc), ctx.upload(rc)
        cc = [ctx.upload(c) for c in cc]
        ups = [lc, rc] + cc
    j = ctx.join_gather(lc, rc, how, cc)
    try:
        gl, gr = j.indices()
        got = [j.right_col(k) for k in range(len(cols))]
    finally:
        j.close()
        for c in ups:
            ctx.free(c)
    wl, wr = oracle.join(L.cpu(oracle), R.cpu(oracle), how)
    assert len(gl) == len(wl)
    order = np.lexsort((gr, gl))
    worder = np.lexsort((wr, wl))
    assert np.array_equal(gl[order], wl[worder]) and np.array_equal(gr[order], wr[worder])
    for k, c in enumerate(cols):
        want = oracle.gather(c.cpu(oracle), wr[worder])
        g = got[k][order]
        if c.dtype == pb.F64:
            assert np.array_equal(g.view(np.uint64), want.view(np.uint64)), k
        else:
            assert np.array_equal(g, want), k


@pytest.mark.parametrize("how", [pb.INNER, pb.LEFT])
def test_join_gather_right_columns(ctx, oracle, how):
    # BASELINE.json configs[2]: unique i64 build keys + 2 payload columns of the build side.  Large inputs: the columns travel
    # with the build rows through the partition and the 32-byte table slots; everything else (small inputs, other dtypes,
    # more than two columns, duplicate keys, Right / Outer) gathers by right row.  Same results either way.
    rng = np.random.default_rng(81)
    nb, npr = 200_000, 1_500_007
    Ru = Spec(pb.I64, oracle.synth_join_keys(nb, unique=True))
    Lu = Spec(pb.I64, oracle.synth_join_keys(npr, domain=2 * nb), nulls=rng.random(npr) < 0.01)
    p_i = Spec(pb.I64, rng.integers(-2**62, 2**62, nb), nulls=rng.random(nb) < 0.1)
    p_f = Spec(pb.F64, rng.normal(0, 1e6, nb))
    p_f.values[::97] = np.nan
    p_32 = Spec(pb.I32, rng.integers(-1000, 1000, nb).astype(np.int32))
    p_b = Spec(pb.BOOL_BITS, rng.random(nb) < 0.5, nulls=rng.random(nb) < 0.1)
    _compare_join_gather(ctx, oracle, Lu, Ru, how, [p_i, p_f])                      # small table: direct path + gathers
    ctx.set_option("join_algo", 2)
    try:
        for cols in ([p_f], [p_i, p_f]):
            _compare_join_gather(ctx, oracle, Lu, Ru, how, cols)
            assert ctx.stats()["groupby_algo_used"] == 3
        _compare_join_gather(ctx, oracle, Lu, Ru, how, [p_i, p_f], device=False)    # host columns
        _compare_join_gather(ctx, oracle, Lu, Ru, how, [p_i, p_f, p_32, p_b])      # more than two columns / other dtypes: gathered
        keys_with_ones = Ru.values.copy()
        keys_with_ones[7] = -1                                                       # the all-ones key lives in a reserved slot
        Lones = Lu.values.copy()
        Lones[::1000] = -1
        _compare_join_gather(ctx, oracle, Spec(pb.I64, Lones), Spec(pb.I64, keys_with_ones), how, [p_i, p_f])
        Rd = Spec(pb.I64, rng.integers(0, 50_000, nb))                               # duplicate build keys: falls back, same answer
        Ld = Spec(pb.I64, rng.integers(0, 60_000, 300_000))
        _compare_join_gather(ctx, oracle, Ld, Rd, how, [p_i, p_f])
    finally:
        ctx.set_option("join_algo", 0)
    _compare_join_gather(ctx, oracle, Spec(pb.I64, [1, 2, 3]), Spec(pb.I64, [3, 1, 1]), pb.OUTER if how == pb.LEFT else pb.RIGHT,
                         [Spec(pb.F64, [0.5, 1.5, 2.5]), Spec(pb.I64, [7, 8, 9])], device=False)


@pytest.mark.parametrize("dtype", ["f64", "i32", "dict", "bool"])
def test_join_other_key_types(ctx, oracle, dtype):
    rng = np.random.default_rng(8)
    if dtype == "f64":
        mk = lambda n: Spec(pb.F64, rng.choice([0.0, -0.0, 1.5, np.nan, 2.0, 7.25], n))
    elif dtype == "i32":
        mk = lambda n: Spec(pb.I32, rng.integers(-20, 20, n).astype(np.int32))
    elif dtype == "dict":
        pool = [f"k{i}" for i in range(30)]
        mk = lambda n: Spec(pb.DICT_U32, rng.integers(0, 30, n).astype(np.uint32), pool=pool)
    else:
        mk = lambda n: Spec(pb.BOOL_BITS, rng.integers(0, 2, n), nulls=rng.random(n) < 0.3)
    compare_join(pb, oracle, ctx, mk(300), mk(40), pb.INNER)
    compare_join(pb, oracle, ctx, mk(300), mk(40), pb.LEFT)


def test_gather_and_filter(ctx, oracle):
    rng = np.random.default_rng(9)
    n = 10_000
    idx = rng.integers(-1, n, 25_000)
    for spec in (Spec(pb.F64, rng.random(n), nulls=rng.random(n) < 0.1), Spec(pb.I64, rng.integers(-5, 5, n), nulls=rng.random(n) < 0.1),
                 Spec(pb.DICT_U32, rng.integers(0, 9, n).astype(np.uint32), nulls=rng.random(n) < 0.1), Spec(pb.I32, rng.integers(-5, 5, n).astype(np.int32)),
                 Spec(pb.BOOL_BITS, rng.integers(0, 2, n), nulls=rng.random(n) < 0.1)):
        got = ctx.gather(spec.gpu(pb), idx)
        want = oracle.gather(spec.cpu(oracle), idx)
        assert np.array_equal(got, want)
    for m in (5, 63, 64, 65, 100_001):
        mask = Spec(pb.BOOL_BITS, rng.random(m) < 0.6, nulls=rng.random(m) < 0.1)
        assert np.array_equal(ctx.filter_indices(mask.gpu(pb)), oracle.filter_indices(mask.cpu(oracle)))


# ---------------------------------------------------------------- partial states, merge, partitioning (multi-GPU building blocks)
def test_partial_merge_equals_whole(ctx, oracle):
    n = 90_000
    rng = np.random.default_rng(10)
    k = Spec(pb.I64, rng.integers(0, 700, n), nulls=rng.random(n) < 0.01)
    v = Spec(pb.F64, 1e6 + rng.normal(0, 3.0, n), nulls=rng.random(n) < 0.05)
    vi = Spec(pb.I64, rng.integers(-10**9, 10**9, n))
    aggs = [(0, op) for op in ALL6] + [(1, op) for op in ALL6]
    cuts = [0, 20_000, 20_001, 55_555, n]
    parts = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        sl = lambda s: Spec(s.dtype, s.values[a:b], None if s.nulls is None else s.nulls[a:b])
        parts.append(ctx.groupby_partial([sl(k).gpu(pb)], [sl(v).gpu(pb), sl(vi).gpu(pb)], all_stats=True))
    # concatenate the partial results on the device side of the ABI: keys + null flags + states
    tot = sum(p.n_groups for p in parts)
    keys = np.concatenate([p.key(0)[0] for p in parts])
    knull = np.concatenate([p.key(0)[1] for p in parts])
    sts = []
    for vcol in range(2):
        buf = ctx.dev_alloc(tot * 64)
        off = 0
        for p in parts:
            ctx.memcpy(buf + off * 64, p.states_dev(vcol), p.n_groups * 64, 2)
            off += p.n_groups
        sts.append(buf)
    merged = ctx.groupby_merge([pb.Column(pb.I64, keys, pb.pack_bits(knull))], sts, [False, True], tot, aggs)
    got = gpu_groupby_dict(pb, merged, [k], len(aggs))
    whole = compare_groupby(pb, oracle, ctx, [k], [v, vi], aggs)
    assert got.keys() == whole.keys()
    for kt in whole:
        assert got[kt][0] == whole[kt][0]
        for a, (g, w) in enumerate(zip(got[kt][1], whole[kt][1])):
            if aggs[a][1] in (pb.COUNT, pb.MIN, pb.MAX) or a >= 6 and aggs[a][1] == pb.SUM:
                assert g == w
            else:
                assert abs(g - w) <= 1e-12 * max(abs(w), 1e6 if a < 6 else 1e9), (kt, a, g, w)
    merged.close()
    for p in parts:
        p.close()
    for b in sts:
        ctx.dev_free(b)


def test_hash_partition(ctx):
    n, parts = 50_000, 8
    rng = np.random.default_rng(11)
    k = rng.integers(0, 3000, n)
    kn = rng.random(n) < 0.02
    col = pb.Column.int64(k, kn)
    perm = ctx.dev_alloc(n * 8)
    counts = ctx.hash_partition([col], parts, perm)
    p = ctx.to_host(perm, n, np.int64)
    ctx.dev_free(perm)
    assert counts.sum() == n and np.array_equal(np.sort(p), np.arange(n))
    dest = np.repeat(np.arange(parts), counts)
    d_of_row = np.empty(n, np.int64)
    d_of_row[p] = dest
    for key in np.unique(k[~kn])[:500]:
        assert len(np.unique(d_of_row[(k == key) & ~kn])) == 1
    assert np.all(d_of_row[kn] == 0)
    assert counts.min() > n / parts * 0.7


# ---------------------------------------------------------------- full-size properties (BASELINE.json sizes)
def test_two_kernels_agree_at_scale(ctx):
    # 2^27 rows: the tile-sort, shared-memory and global-table kernels are independent code paths and must
    # agree: bit-exact keys / counts / min / max, 1e-12 on sum / mean / std; counts sum to n
    n = 1 << 27
    keys = ctx.synth_keys(n, card=1000)
    vals = ctx.synth_vals(n, null_per_million=50_000)
    aggs = [(0, op) for op in ALL6]
    out = []
    for algo in (pb.GB_TILESORT, pb.GB_SHARED, pb.GB_GLOBAL):
        ctx.set_option("groupby_algo", algo)
        r = ctx.groupby_agg([keys], [vals], aggs)
        ctx.set_option("groupby_algo", pb.GB_AUTO)
        k, isnull = r.key(0)
        order = np.argsort(k)
        out.append((k[order], r.group_rows()[order], r.valid_n(0)[order], [r.agg(a)[order] for a in range(6)]))
        r.close()
    a = out[0]
    assert len(a[0]) == 1000 and a[1].sum() == n and abs(a[2].sum() / n - 0.95) < 1e-3
    for b in out[1:]:
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        for i, op in enumerate(ALL6):
            if op in (pb.MIN, pb.MAX, pb.COUNT):
                assert np.array_equal(a[3][i], b[3][i])
            else:
                assert np.allclose(a[3][i], b[3][i], rtol=1e-12, atol=0)
    assert np.array_equal(a[3][4], a[1].astype(np.float64))


def test_full_size_groupby_properties(ctx):
    # BASELINE.json configs[1] at its full size (1e9 rows, 5% NULL values) through size-independent properties:
    # counts sum to n, valid counts to the non-NULL rows, sum = mean * n, min <= mean <= max, std >= 0, and the
    # 1K-group result agrees with the sum over a finer grouping (key and key // 1000 of the same rows)
    n = 1_000_000_000
    vals = ctx.synth_vals(n, null_per_million=50_000)
    fine = ctx.synth_keys(n, card=1_000_000)                 # hash-partitioned path
    aggs = [(0, op) for op in ALL6]
    res = {}
    for name, card in (("1k", 1000), ("10m", 10_000_000)):
        keys = ctx.synth_keys(n, card=card)
        r = ctx.groupby_agg([keys], [vals], aggs)
        k, _ = r.key(0)
        rows, nv = r.group_rows(), r.valid_n(0)
        a = [r.agg(i) for i in range(6)]
        r.close()
        del keys
        assert len(k) == card and len(np.unique(k)) == card
        assert rows.sum() == n and abs(nv.sum() / n - 0.95) < 1e-4
        assert np.array_equal(a[4], rows.astype(np.float64))
        ok = nv > 0
        assert np.allclose(a[0][ok], a[1][ok] * nv[ok], rtol=1e-12)
        assert (a[2][ok] <= a[1][ok]).all() and (a[1][ok] <= a[3][ok]).all() and (a[5] >= 0).all()
        assert (a[2][ok] >= 0).all() and (a[3] < 1000.0).all()
        res[name] = (rows.sum(), nv.sum(), a[0].sum())
    r = ctx.groupby_agg([fine], [vals], [(0, pb.SUM), (0, pb.COUNT)])
    s_fine, c_fine = r.agg(0).sum(), r.agg(1).sum()
    r.close()
    assert c_fine == n
    for name in res:
        assert abs(res[name][2] - s_fine) <= 1e-9 * abs(s_fine)      # the same 1e9 values summed along three different groupings


def test_full_size_join_properties(ctx):
    # BASELINE.json configs[2] at its full size (1e9-row probe x 1e8-row build, unique keys, ~50% hits):
    # |left| = n_probe, |inner| + unmatched = n_probe, every pair joins equal keys, no probe row appears twice
    nb, npr = 100_000_000, 1_000_000_000
    build = ctx.synth_join_keys(nb, unique=True)
    probe = ctx.synth_join_keys(npr, domain=2 * nb)
    j = ctx.join_pairs(probe, build, pb.INNER)
    m_inner = j.n
    assert 0.49 * npr < m_inner < 0.51 * npr
    # key equality on the device for a sample window of the pairs + the whole left index column is a set
    w = min(m_inner, 20_000_000)
    for off in (0, m_inner - w):
        lk = ctx.gather(probe, j.left_dev() + 8 * off, n=w, idx_dev=True, out_dev=ctx.dev_alloc(w * 8))
        rk = ctx.gather(build, j.right_dev() + 8 * off, n=w, idx_dev=True, out_dev=ctx.dev_alloc(w * 8))
        a, b = ctx.to_host(lk, w, np.int64), ctx.to_host(rk, w, np.int64)
        ctx.dev_free(lk); ctx.dev_free(rk)
        assert np.array_equal(a, b)
    li = ctx.to_host(j.left_dev(), min(m_inner, 50_000_000), np.int64)
    assert len(np.unique(li)) == len(li) and li.min() >= 0 and li.max() < npr
    j.close()
    j = ctx.join_pairs(probe, build, pb.LEFT)
    assert j.n == npr
    ri = ctx.to_host(j.right_dev(), 50_000_000, np.int64)
    assert ((ri >= -1) & (ri < nb)).all()
    j.close()
    j = ctx.join_pairs(probe, build, pb.RIGHT)          # unique build keys: inner pairs + the build rows nobody asked for
    assert j.n >= m_inner and j.n - m_inner < nb
    j.close()


# ---------------------------------------------------------------- multi-GPU plumbing on one GPU (world size 1)
class _OneRankDist:
    """torch.distributed look-alike for a single rank: every collective is a copy."""

    @staticmethod
    def get_world_size(): return 1
    @staticmethod
    def get_rank(): return 0
    @staticmethod
    def all_gather(outs, t): outs[0].copy_(t)
    @staticmethod
    def all_to_all_single(out, inp, output_split_sizes=None, input_split_sizes=None): out.copy_(inp)


def test_dist_operators_single_rank(ctx, oracle):
    # the CUDA backend of pandrs_b200/dist.py (device pointers <-> tensors, state rows, NULL re-packing, permuted
    # gathers) with identity collectives must reproduce the single-GPU result; world size 2 runs under gloo on CPU
    # (tests/test_dist_gloo.py) and under NCCL in bench.py --gpus 2
    import torch
    from pandrs_b200.dist import CudaBackend, DistGroupBy, DistJoin
    n = 60_000
    rng = np.random.default_rng(13)
    k = Spec(pb.I64, rng.integers(0, 900, n), nulls=rng.random(n) < 0.02)
    v = Spec(pb.F64, rng.normal(10, 3, n), nulls=rng.random(n) < 0.05)
    aggs = [(0, op) for op in ALL6]
    want = compare_groupby(pb, oracle, ctx, [k], [v], aggs)
    kc, vc = ctx.upload(k.gpu(pb)), ctx.upload(v.gpu(pb))
    for method in ("groupby_agg_lowcard", "groupby_agg_shuffle"):
        r = getattr(DistGroupBy(ctx, _OneRankDist), method)([kc], [vc], aggs)
        got = gpu_groupby_dict(pb, r, [k], len(aggs))
        r.close()
        assert got.keys() == want.keys()
        for kt in want:
            assert got[kt][0] == want[kt][0]
            assert np.allclose(got[kt][1], want[kt][1], rtol=1e-12, atol=0), (method, kt)
    L = Spec(pb.I64, rng.integers(0, 500, 5000), nulls=rng.random(5000) < 0.03)
    R = Spec(pb.I64, rng.integers(0, 500, 1500), nulls=rng.random(1500) < 0.03)
    lc, rc = ctx.upload(L.gpu(pb)), ctx.upload(R.gpu(pb))
    for how in (pb.INNER, pb.LEFT):
        gl, gr = DistJoin(ctx, _OneRankDist).join_pairs(lc, rc, how, 1000, 7000)
        wl, wr = oracle.join(L.cpu(oracle), R.cpu(oracle), how)
        wr = np.where(wr >= 0, wr + 7000, -1)
        assert sorted(zip(gl.cpu().tolist(), gr.cpu().tolist())) == sorted(zip((wl + 1000).tolist(), wr.tolist()))
    for c in (kc, vc, lc, rc):
        ctx.free(c)


# ---------------------------------------------------------------- Zipf-skewed keys (BASELINE.json configs[3])
def _zipf(rng, n, domain, s=1.1):
    w = 1.0 / np.arange(1, domain + 1) ** s
    return rng.choice(domain, size=n, p=w / w.sum())


@pytest.mark.gpu
def test_zipf_dictionary_key_takes_the_skew_fallback(ctx, oracle):
    # ~10^4 groups: too many for one tile-sort table, and the hot keys overflow their hash partition -> the tile-sort
    # kernel keeps the hot keys in registers and spills the tail row by row to the global table
    n = 1_300_000
    rng = np.random.default_rng(31)
    pool = [f"s{i}" for i in range(10_000)]
    k = Spec(pb.DICT_U32, _zipf(rng, n, 10_000).astype(np.uint32), pool=pool)
    v = Spec(pb.F64, rng.normal(50.0, 20.0, n), nulls=rng.random(n) < 0.05)
    compare_groupby(pb, oracle, ctx, [k], [v], [(0, op) for op in ALL6], device=True)
    assert ctx.stats()["groupby_algo_used"] in (pb.GB_TILESORT, pb.GB_PARTITIONED)
    ku = Spec(pb.DICT_U32, rng.integers(0, 10_000, n).astype(np.uint32), pool=pool)      # uniform: the partitioned path, packed keys
    compare_groupby(pb, oracle, ctx, [ku], [v], [(0, op) for op in ALL6], device=True)
    assert ctx.stats()["groupby_algo_used"] == pb.GB_PARTITIONED
    k32 = Spec(pb.I32, rng.integers(-3000, 3000, n).astype(np.int32))                   # (i32, bool) pair in one word
    kb = Spec(pb.BOOL_BITS, rng.random(n) < 0.3)
    compare_groupby(pb, oracle, ctx, [k32, kb], [v], [(0, pb.SUM), (0, pb.STD), (0, pb.COUNT)], device=True)
    assert ctx.stats()["groupby_algo_used"] == pb.GB_PARTITIONED


@pytest.mark.gpu
def test_zipf_high_cardinality_two_level_partition_with_side_areas(ctx, oracle):
    # hundreds of thousands of groups AND hot keys (the top key owns ~13% of the rows): two partition levels, both
    # with a side area for the runs of full buckets; level 1's parked rows are carried over to level 2's area
    n = 3_000_000
    rng = np.random.default_rng(33)
    k = Spec(pb.I64, _zipf(rng, n, 400_000).astype(np.int64) * 1_000_003 - 77)
    v = Spec(pb.F64, rng.normal(10.0, 4.0, n), nulls=rng.random(n) < 0.05)
    got = compare_groupby(pb, oracle, ctx, [k], [v], [(0, op) for op in ALL6], device=True)
    assert len(got) == len(np.unique(k.values))
    assert ctx.stats()["groupby_algo_used"] in (pb.GB_PARTITIONED, pb.GB_TILESORT)       # (tile-sort + spill buffer if a side area overflowed)
    k32 = Spec(pb.I32, _zipf(rng, n, 700).astype(np.int32) - 350)                        # the same through a packed (i32, i64) tuple
    k64 = Spec(pb.I64, _zipf(rng, n, 3000).astype(np.int64) * 11 + 5_000_000_000)
    compare_groupby(pb, oracle, ctx, [k32, k64], [v], [(0, pb.SUM), (0, pb.MEAN), (0, pb.STD), (0, pb.COUNT)], device=True)


@pytest.mark.gpu
def test_zipf_multi_key_underestimated_cardinality_retries_fast(ctx, oracle):
    # (i32, i64) and (i32, i64, dictionary) tuples are wider than one word: global table.  A tiny sample of Zipf keys
    # underestimates the number of groups by far; the table must report the overflow quickly (not degenerate into
    # whole-table probe sequences) and the retry must give the right answer.
    n = 400_000
    rng = np.random.default_rng(32)
    pool = [f"p{i}" for i in range(10_000)]
    k1 = Spec(pb.I32, _zipf(rng, n, 1000).astype(np.int32))
    k2 = Spec(pb.I64, _zipf(rng, n, 100_000).astype(np.int64) * 1_000_003)
    k3 = Spec(pb.DICT_U32, _zipf(rng, n, 10_000).astype(np.uint32), pool=pool)
    v = Spec(pb.F64, rng.random(n) * 1000.0)
    ctx.set_option("sample_rows", 4096)
    ctx.set_option("key_compress", 0)        # keep the tuples two / three words wide
    try:
        for keys in ([k1, k2], [k1, k2, k3]):
            compare_groupby(pb, oracle, ctx, keys, [v], [(0, op) for op in ALL6], device=True)
            assert ctx.stats()["retries"] >= 1
    finally:
        ctx.set_option("sample_rows", 1 << 18)
        ctx.set_option("key_compress", 1)


@pytest.mark.gpu
def test_multi_key_tuples_are_packed_by_value_range(ctx, oracle):
    # (i32, i64[, dictionary, bool]) needs 2-3 words at natural widths; the exact value ranges fit one word, so the tuple
    # is packed as (value - column minimum) fields and takes the one-word kernels.  Negative minima, a large i64
    # offset, NULL parts, a "NULL" dictionary alias and a row filter must survive the round trip through the packing.
    n = 500_000
    rng = np.random.default_rng(41)
    pool = [f"d{i}" for i in range(5000)] + ["NULL"]
    k1 = Spec(pb.I32, rng.integers(-700, 300, n).astype(np.int32), nulls=rng.random(n) < 0.03)
    k2 = Spec(pb.I64, rng.integers(0, 900, n).astype(np.int64) * 7 + 9_000_000_000_000)
    k3 = Spec(pb.DICT_U32, rng.integers(3, 5001, n).astype(np.uint32), pool=pool, null_alias=5000)
    kb = Spec(pb.BOOL_BITS, rng.random(n) < 0.5)
    v = Spec(pb.F64, rng.normal(3.0, 1.0, n), nulls=rng.random(n) < 0.05)
    f = Spec(pb.BOOL_BITS, rng.random(n) < 0.9, nulls=rng.random(n) < 0.02)
    aggs = [(0, op) for op in ALL6]
    k1z = Spec(pb.I32, _zipf(rng, n, 40).astype(np.int32) - 20)
    k2z = Spec(pb.I64, _zipf(rng, n, 30).astype(np.int64) * 1000 - 5)
    compare_groupby(pb, oracle, ctx, [k1z, k2z], [v], aggs, device=True)             # ~1000 groups: tile-sort
    assert ctx.stats()["groupby_algo_used"] == pb.GB_TILESORT
    compare_groupby(pb, oracle, ctx, [k1, k2], [v], aggs, device=True)               # NULL parts in the tuple (one word, global table)
    compare_groupby(pb, oracle, ctx, [k2, k3, kb, k1], [v], aggs, filter_spec=f, device=True)
    k1n = Spec(pb.I32, rng.integers(-700, 300, n).astype(np.int32))
    m = 1_200_000                                                                    # 10^4 tuples, no NULLs: partitioned path on packed words
    k1p = Spec(pb.I32, rng.integers(-70, 30, m).astype(np.int32))
    k2p = Spec(pb.I64, rng.integers(0, 100, m).astype(np.int64) * 7 + 9_000_000_000_000)
    compare_groupby(pb, oracle, ctx, [k1p, k2p], [Spec(pb.F64, rng.normal(3.0, 1.0, m))], aggs, device=True)
    assert ctx.stats()["groupby_algo_used"] == pb.GB_PARTITIONED
    wide = Spec(pb.I64, rng.integers(-2**62, 2**62, n))                              # range does not fit: natural layout
    compare_groupby(pb, oracle, ctx, [k1n, wide], [v], [(0, pb.SUM), (0, pb.COUNT)], device=True)


# ---------------------------------------------------------------- fused partition + shuffle join (pdrs_xjoin_*)
def _xjoin_simulated(ctx, oracle, world, L, R, how, opt_log_nb=0, mode=0):
    """All `world` ranks live in this process on one GPU: rank r's receive area is handed to the others as a plain
    device pointer (attach_ptrs), so the partition kernel's peer stores, the sub-bucket layout, the count
    publication and the global row numbers are exercised exactly as with CUDA IPC between processes."""
    nl, nr = len(L), len(R)
    lcut = [nl * r // world for r in range(world + 1)]
    rcut = [nr * r // world for r in range(world + 1)]
    if opt_log_nb:
        ctx.set_option("join_log_nb", opt_log_nb)
    ctx.set_option("xjoin_mode", mode)        # 1 = fused (rank x bucket in one pass), 2 = staged (by rank, then the local radix partition)
    xs = [pb.XJoin(ctx, r, world, max(lcut[i + 1] - lcut[i] for i in range(world)), max(rcut[i + 1] - rcut[i] for i in range(world)), nr)
          for r in range(world)]
    try:
        for x in xs:
            x.attach_ptrs([y.base for y in xs])
        cols = []
        for r, x in enumerate(xs):
            lspec = Spec(L.dtype, L.values[lcut[r]:lcut[r + 1]], None if L.nulls is None else L.nulls[lcut[r]:lcut[r + 1]])
            rspec = Spec(R.dtype, R.values[rcut[r]:rcut[r + 1]], None if R.nulls is None else R.nulls[rcut[r]:rcut[r + 1]])
            lc, rc = ctx.upload(lspec.gpu(pb)), ctx.upload(rspec.gpu(pb))
            cols += [lc, rc]
            x.shuffle(lc, rc, rcut[r])
        got = []
        for x in xs:
            j = x.local(how, lcut[:world])
            li, ri = j.indices()
            j.close()
            got += list(zip(li.tolist(), ri.tolist()))
        for c in cols:
            ctx.free(c)
    finally:
        for x in xs:
            x.close()
        if opt_log_nb:
            ctx.set_option("join_log_nb", 0)
        ctx.set_option("xjoin_mode", 0)
    wl, wr = oracle.join(L.cpu(oracle), R.cpu(oracle), how)
    assert sorted(got) == sorted(zip(wl.tolist(), wr.tolist())), (world, how)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("world", [1, 2, 8])
@pytest.mark.parametrize("how", [pb.INNER, pb.LEFT])
def test_xjoin_fused_shuffle_unique_build(ctx, oracle, world, how, mode):
    rng = np.random.default_rng(21 + world)
    nb, npr = 40_000, 300_001
    bk = rng.permutation(2 * nb)[:nb].astype(np.int64) * 7919 - 100_000          # unique build keys, ~50% hits
    pk = (rng.integers(0, 2 * nb, npr) * 7919 - 100_000).astype(np.int64)
    _xjoin_simulated(ctx, oracle, world, Spec(pb.I64, pk), Spec(pb.I64, bk), how, opt_log_nb=3, mode=mode)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_xjoin_fused_shuffle_duplicates_nulls_and_extremes(ctx, oracle, world):
    # duplicate build keys take the count / scan / write path with chains indexed by GLOBAL right rows; NULL keys
    # never travel; the all-ones key lives in the reserved slot
    rng = np.random.default_rng(5)
    nl, nr = 50_000, 9_000
    lk = rng.integers(-300, 300, nl).astype(np.int64)
    rk = rng.integers(-300, 300, nr).astype(np.int64)
    lk[::97] = -1
    rk[::53] = -1
    lk[5], rk[7] = np.iinfo(np.int64).min, np.iinfo(np.int64).min
    ln, rn = rng.random(nl) < 0.02, rng.random(nr) < 0.02
    for how in (pb.INNER, pb.LEFT):
        for mode, log_nb in ((1, 0), (2, 0), (2, 2)):
            _xjoin_simulated(ctx, oracle, world, Spec(pb.I64, lk, nulls=ln), Spec(pb.I64, rk, nulls=rn), how, opt_log_nb=log_nb, mode=mode)


@pytest.mark.gpu
def test_xjoin_overflow_is_reported_not_hidden(ctx):
    # one key repeated: every row goes to one sub-bucket, which overflows its padded range -> UNSUPPORTED (the host falls back)
    n = 2_000_000
    lc, rc = ctx.upload(pb.Column.int64(np.full(n, 12345))), ctx.upload(pb.Column.int64(np.full(n, 12345)))
    ctx.set_option("join_log_nb", 6)                  # 64 sub-buckets of ~n / 64 rows
    x = pb.XJoin(ctx, 0, 1, n, n, n)
    ctx.set_option("join_log_nb", 0)
    try:
        with pytest.raises(pb.PandrsError) as e:
            x.shuffle(lc, rc, 0)
        assert e.value.kind == "OperationFailed"
    finally:
        x.close()
        ctx.free(lc); ctx.free(rc)
