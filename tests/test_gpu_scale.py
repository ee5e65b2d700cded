"""The CUDA path against an INDEPENDENT CPU result at BASELINE.json's sizes (1e8 - 1e9 rows): the typed-key oracle
(oracle/typed_oracle.cpp, bit-identical to the string oracle - tests/test_oracle_golden.py) computes the reference's
results on the box's host cores from the same counter-based generators the device columns come from.  These are the
paths that only switch on at scale: the tile-sort kernel over 120 000 tiles, the two-level hash partition at 1e7
groups, the radix join with a multi-GB table, the side areas under Zipf skew.

PDRS_SCALE=<float> scales every row count (debugging); rows are also scaled down when the host has little memory."""
import os

import numpy as np
import pytest

import pandrs_b200 as pb
from _util import canon_pairs, compare_groupby_typed, device_pair_stats, mem_available_gb

pytestmark = pytest.mark.gpu

SCALE = float(os.environ.get("PDRS_SCALE", "1"))
ALL7 = [pb.SUM, pb.MEAN, pb.MIN, pb.MAX, pb.COUNT, pb.STD, pb.VAR]


def _rows(n):
    return max(1 << 20, int(n * SCALE))


@pytest.mark.parametrize("card,rows,scramble", [(1000, 1_000_000_000, False), (1000, 250_000_000, True), (10_000_000, 300_000_000, False)])
def test_config1_groupby_vs_typed_oracle(ctx, oracle, card, rows, scramble):
    # BASELINE.json configs[1]: i64 key (1K / 10M distinct), f64 value with 5% NULLs, sum / mean / min / max / count / std (+ var)
    n = _rows(rows)
    keys = ctx.synth_keys(n, card=card, scramble=scramble)
    vals = ctx.synth_vals(n, null_per_million=50_000)
    r = ctx.groupby_agg([keys], [vals], [(0, op) for op in ALL7])
    algo = ctx.stats()["groupby_algo_used"]
    try:
        tg = oracle.typed_groupby_synth(n, card=card, scramble=scramble, null_per_million=50_000)
        assert tg["group_rows"].sum() == n
        worst = compare_groupby_typed(pb, r, tg, [pb.I64], ALL7)
    finally:
        r.close()
    assert algo == (pb.GB_TILESORT if card <= 2000 else pb.GB_PARTITIONED), algo
    print(f"\n[scale] groupby {n} rows, {card} groups{' (hashed keys)' if scramble else ''}: algo {algo}, worst relative error vs reference / vs exact: {worst}")


def _join_case(ctx, oracle, npr, nb, full_pairs):
    build = ctx.synth_join_keys(nb, unique=True)
    probe = ctx.synth_join_keys(npr, domain=2 * nb)
    want = oracle.typed_join(left_synth=dict(n=npr, domain=2 * nb), right_synth=dict(n=nb, unique=True), how=oracle.LEFT, want_pairs=full_pairs)
    m_inner = want["n"] - want["unmatched_left"]
    mask = (1 << 64) - 1
    for how in (pb.INNER, pb.LEFT):
        j = ctx.join_pairs(probe, build, how)
        try:
            n, cs, sl, sr, un = device_pair_stats(ctx, j)
            if how == pb.LEFT:
                assert (n, cs, sr, un) == (want["n"], want["checksum"], want["sum_right"], want["unmatched_left"])
                assert sl == want["sum_left"]
            else:   # unique build keys: the Inner pairs are the Left pairs with a right row
                assert (n, un, sr) == (m_inner, 0, want["sum_right"])
                assert cs == (want["checksum"] - want["checksum_unmatched"]) & mask
            if full_pairs:
                gl, gr = canon_pairs(*j.indices())
                wl, wr = want["left"], want["right"]
                if how == pb.INNER:
                    keep = wr >= 0
                    wl, wr = wl[keep], wr[keep]
                assert np.array_equal(gl, wl) and np.array_equal(gr, wr)      # the oracle's pairs are already in canonical order
        finally:
            j.close()
    return want["n"], m_inner


def test_config2_join_full_size_vs_typed_oracle(ctx, oracle):
    # BASELINE.json configs[2]: 1e9-row probe x 1e8-row build, unique i64 keys, ~50% hits; the multiset of pairs is compared
    # through its count, an order-independent 64-bit checksum and the index sums (sorting 5e8 pairs on the host is not needed)
    n, m = _join_case(ctx, oracle, _rows(1_000_000_000), _rows(100_000_000), full_pairs=False)
    print(f"\n[scale] join: left {n} pairs, inner {m} pairs: count / checksum / index sums match the typed oracle")


def test_config2_join_1e8_pairs_after_canonical_sort(ctx, oracle):
    # the same join at 1e8 x 1e7: every pair compared after the canonical sort (north star)
    _join_case(ctx, oracle, _rows(100_000_000), _rows(10_000_000), full_pairs=True)


def _zipf_torch(torch, gen, dev, n, domain, dtype, s=1.1):
    w = torch.arange(1, domain + 1, device=dev, dtype=torch.float64).pow(-s)
    cdf = (w.cumsum(0) / w.sum()).to(torch.float32)
    out = torch.empty(n, dtype=dtype, device=dev)
    CH = 1 << 26
    for a in range(0, n, CH):
        b = min(n, a + CH)
        out[a:b] = torch.searchsorted(cdf, torch.rand(b - a, device=dev, generator=gen)).clamp_(max=domain - 1).to(dtype)
    return out


@pytest.mark.parametrize("nkeys,rows", [(2, 500_000_000), (3, 100_000_000)])
def test_config3_zipf_multi_key_vs_typed_oracle(ctx, oracle, nkeys, rows):
    # BASELINE.json configs[3]: (i32, i64) [+ dictionary id] keys, Zipf(1.1) over 1e3 / 1e5 / 1e4 values, f64 value.
    # Columns are generated on the device (bench.py does the same), copied to the host once, and grouped there by the
    # typed oracle.  5e8 rows = 24 M groups for two keys; the three-key tuple (one group per ~3 rows) runs at 1e8 rows.
    import torch
    n = _rows(rows)
    need_gb = n * (20 + (4 if nkeys == 3 else 0)) / 1e9 + (n / (20 if nkeys == 2 else 2.5)) * 330 / 1e9 + 8
    while need_gb > 0.6 * mem_available_gb() and n > (1 << 22):
        n //= 2
        need_gb /= 2
    dev = torch.device("cuda", ctx.device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(4242)
    k1 = _zipf_torch(torch, gen, dev, n, 1000, torch.int32)
    k2 = _zipf_torch(torch, gen, dev, n, 100_000, torch.int64)
    k3 = _zipf_torch(torch, gen, dev, n, 10_000, torch.int32) if nkeys == 3 else None
    v = torch.rand(n, device=dev, generator=gen, dtype=torch.float64) * 1000.0
    torch.cuda.synchronize(dev)

    def col(dtype, t):
        return pb.Column(dtype, device_ptr=t.data_ptr(), length=n, owner=t)
    gkeys = [col(pb.I32, k1), col(pb.I64, k2)] + ([col(pb.DICT_U32, k3)] if nkeys == 3 else [])
    ops = [pb.SUM, pb.MEAN, pb.MIN, pb.MAX, pb.COUNT, pb.STD]
    r = ctx.groupby_agg(gkeys, [col(pb.F64, v)], [(0, op) for op in ops])
    algo, retries = ctx.stats()["groupby_algo_used"], ctx.stats()["retries"]
    try:
        okeys = [oracle.Col(oracle.I32, k1.cpu().numpy()), oracle.Col(oracle.I64, k2.cpu().numpy())]
        if nkeys == 3:
            okeys.append(oracle.Col(oracle.DICT_U32, k3.cpu().numpy().view(np.uint32)))
        tg = oracle.typed_groupby(okeys, oracle.Col(oracle.F64, v.cpu().numpy()))
        worst = compare_groupby_typed(pb, r, tg, [pb.I32, pb.I64, pb.DICT_U32][:nkeys], ops)
    finally:
        r.close()
    print(f"\n[scale] configs[3] {nkeys} keys, {n} rows: {tg['n_groups']} groups, algo {algo}, retries {retries}, worst relative error {worst}")


def test_chunked_host_groupby_vs_typed_oracle(ctx, oracle):
    # out-of-core path (pdrs_groupby_agg on HOST columns, 2^26-row chunks through the staging engine) against the typed oracle on
    # the whole input: 2e8 rows = 3 chunks, pageable numpy memory
    n = _rows(200_000_000)
    if mem_available_gb() < 8 * n * 16 / 1e9 / 4:
        pytest.skip("not enough host memory")
    hk = np.empty(n, np.int64)
    hv = np.empty(n, np.float64)
    hn = np.empty((n + 7) // 8, np.uint8)
    keys = ctx.synth_keys(n, card=1000)
    vals = ctx.synth_vals(n, null_per_million=50_000)
    ctx.memcpy(hk.ctypes.data, keys.ptr, 8 * n, 1)
    ctx.memcpy(hv.ctypes.data, vals.ptr, 8 * n, 1)
    ctx.memcpy(hn.ctypes.data, vals.nulls_ptr, (n + 7) // 8, 1)
    ctx.free(keys); ctx.free(vals)
    launches0 = ctx.stats()["kernel_launches"]
    r = ctx.groupby_agg([pb.Column(pb.I64, hk)], [pb.Column(pb.F64, hv, hn)], [(0, op) for op in ALL7])
    try:
        assert ctx.stats()["kernel_launches"] - launches0 >= 3 * ((n + (1 << 26) - 1) >> 26)     # one partial aggregation per chunk
        tg = oracle.typed_groupby_synth(n, card=1000, null_per_million=50_000)
        worst = compare_groupby_typed(pb, r, tg, [pb.I64], ALL7)
    finally:
        r.close()
    print(f"\n[scale] chunked host groupby {n} rows: worst relative error vs reference / vs exact: {worst}")


@pytest.mark.parametrize("card", [1000, 3_000_000])
def test_row_lists_at_scale(ctx, oracle, card):
    # pdrs_groupby_rows (par_groupby, grouping.rs:124-331) at 2e8 rows: a permutation of the rows, ascending inside every group,
    # every group holds one key only, sizes = the typed oracle's group sizes
    import torch
    n = _rows(200_000_000)
    keys = ctx.synth_keys(n, card=card)
    res = ctx.groupby_rows([keys])
    try:
        G = res.n_groups
        tg = oracle.typed_groupby_synth(n, card=card, null_per_million=0)
        assert G == tg["n_groups"] and res.n_rows == n
        kv, kn = res.key(0)
        off = res.offsets()
        assert not kn.any() and off[0] == 0 and off[-1] == n
        o, to = np.argsort(kv), np.argsort(tg["keys"][0][0].view(np.int64))
        assert np.array_equal(kv[o], tg["keys"][0][0].view(np.int64)[to]) and np.array_equal(np.diff(off)[o], tg["group_rows"][to])
        dev = torch.device("cuda", ctx.device)
        ids = torch.empty(n, dtype=torch.int64, device=dev)
        k = torch.empty(n, dtype=torch.int64, device=dev)
        torch.cuda.synchronize(dev)
        ctx.memcpy(ids.data_ptr(), res.ids_dev(), 8 * n, 2)
        ctx.memcpy(k.data_ptr(), keys.ptr, 8 * n, 2)
        assert int(ids.sum().item()) == n * (n - 1) // 2 and int(ids.min().item()) == 0 and int(ids.max().item()) == n - 1
        gk = k[ids]                                           # keys in grouped order
        offs = torch.from_numpy(off).to(dev)
        seg = torch.repeat_interleave(torch.arange(G, device=dev), offs[1:] - offs[:-1])
        assert bool((gk == torch.from_numpy(kv).to(dev)[seg]).all()), "a group holds rows of another key"
        d = ids[1:] - ids[:-1]
        inside = seg[1:] == seg[:-1]
        assert bool((d[inside] > 0).all()), "rows of a group must ascend"
    finally:
        res.close()
        ctx.free(keys)
