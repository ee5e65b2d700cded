"""Run under torchrun with >= 2 GPUs (tests/test_gpu_multi.py; `gpurun --gpus N`): the multi-GPU operators behind the C ABI
(pdrs_comm_init / pdrs_groupby_agg_dist / pdrs_join_pairs_dist, NCCL bound inside libpandrs_b200.so) on `world` ranks against the
oracle on the UNION of the ranks' rows:
  * low cardinality, replicated result (fixed-size all-gather of per-group states + merge) - the path bench.py --gpus N times
  * high cardinality, sharded result (all-to-all of per-group states by hash(key)), forced and chosen automatically
  * multi-key with NULL key parts, Int64 values, NULL values, a Boolean filter with NULLs
  * configs[4]-style filter -> groupby(returnflag, linestatus) with five value columns (few-groups kernel) and a typed predicate
  * the sharded Inner / Left join (peer-store shuffle) in global row numbers
torch.distributed only broadcasts the 128-byte NCCL id and collects the results for the check."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as orc  # noqa: E402  (tests may use the oracle as the checker)
import pandrs_b200 as pb  # noqa: E402
from _util import Spec, gpu_groupby_dict, oracle_groupby_dict  # noqa: E402

ALL6 = [pb.SUM, pb.MEAN, pb.MIN, pb.MAX, pb.COUNT, pb.STD]


def gather_dicts(d, world):
    out = [None] * world
    dist.all_gather_object(out, d)
    return out


def check(name, rank, world, comm, ctx, keys, vals, aggs, filt=None, pred=None, mode=0, expect_sharded=None, compat=False):
    """keys / vals / filt: Specs of the UNION; every rank aggregates its contiguous slice."""
    n = len(keys[0])
    lo, hi = n * rank // world, n * (rank + 1) // world

    def cut(s):
        return Spec(s.dtype, s.values[lo:hi], None if s.nulls is None else s.nulls[lo:hi], s.pool, s.null_alias)
    kc = [ctx.upload(cut(s).gpu(pb)) for s in keys]
    vc = [ctx.upload(cut(s).gpu(pb)) for s in vals]
    fc = None if filt is None else ctx.upload(cut(filt).gpu(pb))
    pc = None if pred is None else (ctx.upload(cut(pred[0]).gpu(pb)), pred[1], pred[2])
    r = comm.groupby_agg(kc, vc, aggs, filter=fc, pred=pc, result_mode=mode)
    got = gpu_groupby_dict(pb, r, keys, len(aggs))
    r.close()
    for c in kc + vc + ([fc] if fc is not None else []) + ([pc[0]] if pc is not None else []):
        ctx.free(c)
    alld = gather_dicts(got, world)
    if rank != 0:
        return
    fspec = filt
    if pred is not None:
        cmpf = {pb.CMP_LE: np.less_equal, pb.CMP_GT: np.greater, pb.CMP_LT: np.less, pb.CMP_GE: np.greater_equal}[pred[1]]
        keep = cmpf(pred[0].values, pred[2])
        if pred[0].nulls is not None:
            keep &= ~pred[0].nulls
        if filt is not None:
            keep &= filt.values.astype(bool) & (~filt.nulls if filt.nulls is not None else True)
        fspec = Spec(pb.BOOL_BITS, keep)
    want = oracle_groupby_dict(orc, keys, vals, aggs, fspec, compat)
    sizes = [len(d) for d in alld]
    sharded = not all(set(d) == set(want) for d in alld)
    if sharded:       # every group lives on exactly one rank
        union = {}
        for d in alld:
            assert not (set(d) & set(union)), f"{name}: a group was returned by two ranks"
            union.update(d)
        results = [union]
    else:
        results = alld
    if expect_sharded is not None:
        assert sharded == expect_sharded, (name, sharded, sizes)
    for d in results:
        assert set(d) == set(want), (name, len(d), len(want))
        for kt, (rows, wv) in want.items():
            grows, gv = d[kt]
            assert grows == rows, (name, kt, grows, rows)
            for a, (v, op) in enumerate(aggs):
                w, g = wv[a], gv[a]
                if op in (pb.COUNT, pb.MIN, pb.MAX) or (vals[v].dtype == pb.I64 and op in (pb.SUM, pb.MEAN)):
                    assert g == w, (name, kt, op, g, w)
                else:
                    assert abs(g - w) <= 1e-12 * max(abs(w), 1e-300) + 1e-9 * (op in (pb.STD, pb.VAR)) * 0, (name, kt, op, g, w)
    print(f"  {name}: {len(want)} groups, {'sharded ' + str(sizes) if sharded else 'replicated'} ok", flush=True)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = pb.Context(device=local)
    comm = pb.Comm(ctx, rank, world, pb.torch_broadcast_id(dist, torch.device("cuda", local)))
    rng = np.random.default_rng(123)                       # the same union on every rank
    n = 400_000
    k1k = Spec(pb.I64, rng.integers(0, 1000, n))
    v = Spec(pb.F64, rng.random(n) * 1000, nulls=rng.random(n) < 0.05)
    vi = Spec(pb.I64, rng.integers(-10**12, 10**12, n), nulls=rng.random(n) < 0.05)
    check("1K groups, replicated", rank, world, comm, ctx, [k1k], [v], [(0, op) for op in ALL6], mode=0, expect_sharded=False)
    check("1K groups, int + float values", rank, world, comm, ctx, [k1k], [v, vi], [(0, pb.SUM), (1, pb.SUM), (1, pb.MEAN), (1, pb.MIN), (1, pb.STD), (0, pb.VAR)], mode=1)
    khi = Spec(pb.I64, rng.integers(0, 60_000, n) * 7_000_003 - 11, nulls=rng.random(n) < 0.01)
    check("60K groups + NULL keys, sharded (forced)", rank, world, comm, ctx, [khi], [v], [(0, op) for op in ALL6], mode=2, expect_sharded=world > 1)
    check("60K groups, auto -> sharded", rank, world, comm, ctx, [khi], [v], [(0, pb.SUM), (0, pb.COUNT)], mode=0, expect_sharded=world > 1)
    pool = [f"s{i}" for i in range(50)]
    kd = Spec(pb.DICT_U32, rng.integers(0, 50, n).astype(np.uint32), nulls=rng.random(n) < 0.05, pool=pool)
    k32 = Spec(pb.I32, rng.integers(-20, 20, n).astype(np.int32), nulls=rng.random(n) < 0.05)
    f = Spec(pb.BOOL_BITS, rng.random(n) < 0.9, nulls=rng.random(n) < 0.02)
    check("multi-key with NULL parts + filter", rank, world, comm, ctx, [kd, k32], [v, vi], [(0, pb.SUM), (1, pb.MAX), (0, pb.STD), (0, pb.COUNT)], filt=f, mode=0)
    comm.set_option("groups_cap", 64)
    check("multi-key, groups_cap 64 -> sharded", rank, world, comm, ctx, [kd, k32], [v], [(0, pb.MEAN), (0, pb.MIN)], mode=0, expect_sharded=world > 1)
    comm.set_option("groups_cap", 4096)
    # configs[4] shape
    rf = Spec(pb.DICT_U32, rng.integers(0, 3, n).astype(np.uint32), pool=["A", "N", "R"])
    ls = Spec(pb.DICT_U32, rng.integers(0, 2, n).astype(np.uint32), pool=["F", "O"])
    cols = [Spec(pb.F64, rng.random(n) * s) for s in (50, 1e5, 1e5, 1e5, 0.1)]
    ship = Spec(pb.I64, rng.integers(8000, 10600, n))
    q1 = [(0, pb.SUM), (1, pb.SUM), (2, pb.SUM), (3, pb.SUM), (0, pb.MEAN), (1, pb.MEAN), (4, pb.MEAN), (0, pb.COUNT)]
    check("Q1-style filter -> groupby, 5 value columns", rank, world, comm, ctx, [rf, ls], cols, q1, filt=f, mode=1)
    assert ctx.stats()["groupby_algo_used"] == pb.GB_FEW
    check("Q1-style with the date predicate in the scan", rank, world, comm, ctx, [rf, ls], cols, q1, pred=(ship, pb.CMP_LE, 10_500), mode=1)
    # ---- sharded join in global row numbers
    nb, npr = 200_000, 1_500_000
    bk = rng.permutation(2 * nb)[:nb].astype(np.int64) * 1_000_003 - 5
    pk = (rng.integers(0, 2 * nb, npr) * 1_000_003 - 5).astype(np.int64)
    pnull = rng.random(npr) < 0.01
    lcut = [npr * r // world for r in range(world + 1)]
    rcut = [nb * r // world for r in range(world + 1)]
    lc = ctx.upload(pb.Column.int64(pk[lcut[rank]:lcut[rank + 1]], pnull[lcut[rank]:lcut[rank + 1]]))
    rc = ctx.upload(pb.Column.int64(bk[rcut[rank]:rcut[rank + 1]]))
    ml, mr = max(lcut[i + 1] - lcut[i] for i in range(world)), max(rcut[i + 1] - rcut[i] for i in range(world))
    # the last two joins run in ROUNDS of at most 100 000 left rows per rank (what shards beyond the 32-bit row encoding of
    # the staged exchange do): same pairs
    for how, round_rows in ((pb.INNER, 0), (pb.LEFT, 0), (pb.INNER, 0), (pb.LEFT, 100_000), (pb.INNER, 100_000)):
        ctx.set_option("xjoin_round_rows", round_rows)
        j = comm.join_pairs(lc, rc, how, lcut[rank], rcut[rank], ml, mr, nb)
        li, ri = j.indices()
        j.close()
        parts = gather_dicts(np.stack([li, ri], 1), world)
        if rank == 0:
            got = np.concatenate(parts)
            wl, wr = orc.join(orc.Col(orc.I64, pk, orc.pack_bits(pnull)), orc.Col(orc.I64, bk), how)
            want = np.stack([wl, wr], 1)
            got = got[np.lexsort((got[:, 1], got[:, 0]))]
            want = want[np.lexsort((want[:, 1], want[:, 0]))]
            assert got.shape == want.shape and np.array_equal(got, want), (how, got.shape, want.shape)
            ms, nbytes = comm.last_exchange()
            print(f"  join how={how}{' in rounds' if round_rows else ''}: {len(want)} pairs ok (shuffle {ms:.3f} ms, {nbytes / 1e6:.1f} MB to peers)", flush=True)
    ctx.set_option("xjoin_round_rows", 0)
    comm.close()
    ctx.close()
    dist.barrier()
    if rank == 0:
        print(f"dist groupby / join parity ok on {world} GPUs")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
