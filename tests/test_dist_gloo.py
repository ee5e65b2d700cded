"""world_size-2 tests of the multi-GPU host logic (pandrs_b200/dist.py) on CPU with the gloo backend.

The collective plumbing (variable-length all_gather of partial states, count exchange, all_to_all of
hash-partitioned columns, NULL bitmap re-packing, global row ids of distributed joins) is the code under test;
the per-rank compute steps are supplied by an oracle/numpy test backend with the same interface as CudaBackend."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

I64, F64 = 0, 1
SUM, MEAN, MIN, MAX, COUNT, STD = 0, 1, 2, 3, 4, 5


class TCol:
    """Host column of the test backend (same attributes dist.py reads from pandrs_b200.Column)."""

    def __init__(self, dtype, data, nulls=None, null_alias=-1):
        self.dtype, self.data, self.null_alias = dtype, np.asarray(data), null_alias
        self.nulls = None if nulls is None else np.asarray(nulls, dtype=bool)
        self.len = len(self.data)
        self.nulls_ptr = 1 if nulls is not None else 0


def _mix(x):
    x = x.astype(np.uint64)
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15))
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


class OracleBackend:
    """numpy stand-in for CudaBackend: partial states are (rows, n, mean, M2, min, max, isum, 0) as f64 bit patterns."""

    def __init__(self):
        self.device = torch.device("cpu")

    def col(self, dtype, data, nulls=None, length=None, null_alias=-1):
        d = data.numpy()
        nb = None
        if nulls is not None:
            nb = np.unpackbits(nulls.numpy(), bitorder="little")[: len(d)].astype(bool)
        return TCol(dtype, d, nb, null_alias)

    def _groups(self, keys):
        n = keys[0].len
        kt = []
        for r in range(n):
            kt.append(tuple(None if (k.nulls is not None and k.nulls[r]) else k.data[r].item() for k in keys))
        order = {}
        for r, t in enumerate(kt):
            order.setdefault(t, []).append(r)
        return order

    def partial(self, keys, vals, filter, all_stats):
        g = self._groups(keys)
        G = len(g)
        kt = [torch.tensor([0 if t[i] is None else t[i] for t in g], dtype=torch.int64) for i in range(len(keys))]
        kn = [torch.tensor([1 if t[i] is None else 0 for t in g], dtype=torch.uint8) for i in range(len(keys))]
        st = []
        for v in vals:
            s = np.zeros((G, 8), np.float64)
            for gi, rows in enumerate(g.values()):
                rows = np.array(rows)
                ok = rows if v.nulls is None else rows[~v.nulls[rows]]
                x = v.data[ok].astype(np.float64)
                s[gi, 0], s[gi, 1] = len(rows), len(x)
                if len(x):
                    s[gi, 2] = x.mean(); s[gi, 3] = ((x - x.mean()) ** 2).sum(); s[gi, 4] = x.min(); s[gi, 5] = x.max(); s[gi, 6] = x.sum()
            st.append(torch.from_numpy(s.view(np.int64).copy()))
        return kt, kn, st, torch.tensor([len(r) for r in g.values()], dtype=torch.int64)

    def merge(self, key_dtypes, kt, kn, st, val_is_int, aggs, null_alias=None):
        n = len(kt[0])
        groups = {}
        for r in range(n):
            t = tuple(None if kn[i][r] else int(kt[i][r]) for i in range(len(kt)))
            groups.setdefault(t, []).append(r)
        out = {}
        for t, rows in groups.items():
            res = []
            for v, op in aggs:
                s = st[v].numpy().view(np.float64)[rows]
                N_, n_ = s[:, 0].sum(), s[:, 1].sum()
                mean = (s[:, 1] * s[:, 2]).sum() / n_ if n_ else 0.0
                m2 = (s[:, 3] + s[:, 1] * (s[:, 2] - mean) ** 2).sum()      # Chan's parallel update
                has = s[:, 1] > 0
                res.append({SUM: s[:, 6].sum(), MEAN: mean, MIN: s[has, 4].min() if has.any() else 0.0, MAX: s[has, 5].max() if has.any() else 0.0,
                            COUNT: N_, STD: np.sqrt(m2 / (n_ - 1)) if n_ > 1 else 0.0}[op])
            out[t] = res
        return out

    def groupby(self, keys, vals, aggs, filter=None):
        kt, kn, st, _ = self.partial(keys, vals, filter, True)
        return self.merge([k.dtype for k in keys], kt, kn, st, None, aggs)

    def hash_partition(self, keys, nparts):
        h = np.zeros(keys[0].len, np.uint64)
        isnull = np.zeros(keys[0].len, bool)
        for k in keys:
            h = _mix(h ^ _mix(k.data.astype(np.int64).view(np.uint64)))
            if k.nulls is not None:
                isnull |= k.nulls
        dest = (h % np.uint64(nparts)).astype(np.int64)
        if len(keys) == 1:
            dest[isnull] = 0
        perm = np.argsort(dest, kind="stable")
        return torch.from_numpy(perm), np.bincount(dest, minlength=nparts).astype(np.int64)

    def gather(self, col, perm):
        return torch.from_numpy(col.data[perm.numpy()].copy())

    def null_flags(self, col, perm):
        return None if col.nulls is None else torch.from_numpy(col.nulls[perm.numpy()].astype(np.uint8))

    def join(self, left, right, how):
        import oracle as o
        li, ri = o.join(o.Col(o.I64, left.data, None if left.nulls is None else o.pack_bits(left.nulls)),
                        o.Col(o.I64, right.data, None if right.nulls is None else o.pack_bits(right.nulls)), how)
        return torch.from_numpy(li), torch.from_numpy(ri)


def _worker(rank, world, port, case):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pandrs_b200.dist import DistGroupBy, DistJoin, all_gather_varlen, pack_null_bits
        rng = np.random.default_rng(100)           # same stream on every rank: global data, then slice
        n = 4000
        k = rng.integers(0, 37, n)
        kn = rng.random(n) < 0.03
        v = rng.normal(50, 10, n)
        vn = rng.random(n) < 0.05
        cuts = [0, 1500, n] if world == 2 else np.linspace(0, n, world + 1).astype(int)
        a, b = cuts[rank], cuts[rank + 1]
        be = OracleBackend()
        aggs = [(0, op) for op in (SUM, MEAN, MIN, MAX, COUNT, STD)]
        whole = be.groupby([TCol(I64, k, kn)], [TCol(F64, v, vn)], aggs)
        keys, vals = [TCol(I64, k[a:b], kn[a:b])], [TCol(F64, v[a:b], vn[a:b])]
        if case == "varlen":
            t = torch.arange(rank + 2, dtype=torch.int64) + 10 * rank
            g, sizes = all_gather_varlen(dist, t)
            assert sizes == [r + 2 for r in range(world)]
            assert g.tolist() == [x + 10 * r for r in range(world) for x in range(r + 2)]
            bits = pack_null_bits(torch.tensor([1, 0, 0, 1, 0, 0, 0, 0, 1], dtype=torch.uint8))
            assert bits.numel() % 8 == 0 and bits[0] == 9 and bits[1] == 1
        elif case == "lowcard":
            got = DistGroupBy(be, dist).groupby_agg_lowcard(keys, vals, aggs)
            assert got.keys() == whole.keys()
            for t in whole:
                assert np.allclose(got[t], whole[t], rtol=1e-11, atol=1e-9), (t, got[t], whole[t])
        elif case == "shuffle":
            got = DistGroupBy(be, dist).groupby_agg_shuffle(keys, vals, aggs)
            # results stay sharded: the union over ranks is the whole result, the shards are disjoint
            mine = torch.tensor([(-1 if t[0] is None else t[0]) for t in got], dtype=torch.int64)
            allk, _ = all_gather_varlen(dist, mine)
            assert sorted(allk.tolist()) == sorted((-1 if t[0] is None else t[0]) for t in whole)
            for t in got:
                assert np.allclose(got[t], whole[t], rtol=1e-11, atol=1e-9), (t, got[t], whole[t])
        elif case == "join":
            import oracle as o
            lk = rng.integers(0, 300, 3000)
            ln = rng.random(3000) < 0.04
            rk = rng.integers(0, 300, 800)
            rn = rng.random(800) < 0.04
            lc, rc = [0, 1000, 3000], [0, 500, 800]
            for how in (0, 1):
                gl, gr = DistJoin(be, dist).join_pairs_auto(TCol(I64, lk[lc[rank]:lc[rank + 1]], ln[lc[rank]:lc[rank + 1]]),
                                                       TCol(I64, rk[rc[rank]:rc[rank + 1]], rn[rc[rank]:rc[rank + 1]]), how, lc[rank], rc[rank])
                pl, _ = all_gather_varlen(dist, gl)
                pr, _ = all_gather_varlen(dist, gr)
                wl, wr = o.join(o.Col(o.I64, lk, o.pack_bits(ln)), o.Col(o.I64, rk, o.pack_bits(rn)), how)
                assert sorted(zip(pl.tolist(), pr.tolist())) == sorted(zip(wl.tolist(), wr.tolist()))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("case", ["varlen", "lowcard", "shuffle", "join"])
def test_world_size_2(case):
    import oracle
    oracle.build()
    mp.spawn(_worker, args=(2, _free_port(), case), nprocs=2, join=True)
