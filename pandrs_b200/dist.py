"""Multi-GPU groupby / join: one process per GPU, torch.distributed (NCCL over NVLink) for the plumbing.

pandrs has no communication backend: its `distributed` module lowers SQL to an in-process
DataFusion context and `PartitionStrategy::Hash` (src/distributed/core/partition.rs:11-18) is only an enum.
The sharded semantics implemented here are those of the single-frame reference applied to the union of the
ranks' rows:

  groupby, low cardinality   local partial aggregate (pdrs_groupby_partial) -> all_gather of the per-group state
                             rows (64 B per group and value column) -> merge + finalise (pdrs_groupby_merge).
                             No row ever crosses NVLink.
  groupby, high cardinality  rows are hash-partitioned by key (pdrs_hash_partition + pdrs_gather), exchanged with
                             one all_to_all per column, and aggregated locally; every rank owns the groups whose
                             hash maps to it.
  join                       both sides are hash-partitioned by key and exchanged the same way, then joined
                             locally; row ids travel with the keys so the result is in GLOBAL row numbers.
  join, exchange path        (DistJoin.setup_fused / join_pairs_fused, pdrs_xjoin_*) the join's partition kernel
                             stores its runs straight into the destination GPU's receive area (CUDA IPC mappings,
                             NVLink stores): no send buffers, no all_to_all; torch.distributed only carries the
                             64-byte IPC handles once and one small all_gather per call (the barrier).

The compute steps are methods of a `backend` object (default: the CUDA Context); the collective steps only see
tensors.  tests/test_dist_gloo.py drives the same code on CPU tensors with the gloo backend and an oracle-based
test backend, so the N > 1 logic is covered without GPUs.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import _native as N
from .core import Column, Context, NP_DTYPE

_TORCH_DTYPE = {N.I64: torch.int64, N.F64: torch.float64, N.DICT_U32: torch.int32, N.I32: torch.int32, N.BOOL_BITS: torch.uint8}


# ---------------------------------------------------------------- collective plumbing (backend independent)
def all_gather_varlen(dist, t: torch.Tensor):
    """Concatenation of every rank's 1-D/2-D tensor along dim 0 (+ the per-rank lengths)."""
    world = dist.get_world_size()
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes) if sizes else 0
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0), sizes


def exchange_counts(dist, send_counts: np.ndarray, device) -> np.ndarray:
    """recv_counts[r] = what rank r sends to me."""
    world = dist.get_world_size()
    s = torch.tensor(np.asarray(send_counts, dtype=np.int64), device=device)
    r = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(r, s)
    return r.cpu().numpy()


def all_to_all_rows(dist, t: torch.Tensor, send_counts: np.ndarray, recv_counts: np.ndarray) -> torch.Tensor:
    """t is grouped by destination rank (send_counts rows each); returns the rows received, grouped by source."""
    out = torch.empty((int(recv_counts.sum()),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_to_all_single(out, t.contiguous(), output_split_sizes=[int(x) for x in recv_counts], input_split_sizes=[int(x) for x in send_counts])
    return out


def pack_null_bits(flags: torch.Tensor) -> torch.Tensor:
    """uint8 0/1 flags -> pandrs bitmap (LSB first, bit set = NULL), padded to a multiple of 8 bytes."""
    n = flags.shape[0]
    nb = ((n + 7) // 8 + 7) // 8 * 8
    f = torch.zeros(nb * 8, dtype=torch.uint8, device=flags.device)
    f[:n] = flags.to(torch.uint8)
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.int32, device=flags.device)
    return (f.view(nb, 8).to(torch.int32) * w).sum(dim=1).to(torch.uint8)


# ---------------------------------------------------------------- CUDA backend: tensors <-> library
class CudaBackend:
    """Runs the compute steps through libpandrs_b200.so on this rank's GPU; tensors are torch CUDA tensors."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.device = torch.device("cuda", ctx.device)

    def col(self, dtype: int, data: torch.Tensor, nulls: Optional[torch.Tensor] = None, length: Optional[int] = None, null_alias: int = -1) -> Column:
        n = int(length if length is not None else data.shape[0])
        return Column(dtype, device_ptr=data.data_ptr() if data.numel() else 0, nulls_ptr=(nulls.data_ptr() if nulls is not None else None),
                      null_len=(int(nulls.numel()) if nulls is not None else 0), length=n, null_alias=null_alias, owner=(data, nulls))

    def _sync(self):
        """torch ops run on torch's current stream, the library on the context's: order them (the library
        synchronises its own stream before every call returns)."""
        torch.cuda.current_stream(self.device).synchronize()

    def tensor_from_ptr(self, ptr: int, count: int, dtype: torch.dtype) -> torch.Tensor:
        t = torch.empty(count, dtype=dtype, device=self.device)
        if count:
            self.ctx.memcpy(t.data_ptr(), ptr, t.numel() * t.element_size(), 2)
        return t

    def partial(self, keys: Sequence[Column], vals: Sequence[Column], filter, all_stats: bool):
        """-> (key tensors, key-null flag tensors, state tensors [G, 8] int64 per value column)"""
        self._sync()
        r = self.ctx.groupby_partial(keys, vals, filter=filter, all_stats=all_stats)
        try:
            G = r.n_groups
            kt = [self.tensor_from_ptr(r.key_dev(k), G, _TORCH_DTYPE[c.dtype]) for k, c in enumerate(keys)]
            kn = [self.tensor_from_ptr(r.key_null_dev(k), G, torch.uint8) for k in range(len(keys))]
            st = [self.tensor_from_ptr(r.states_dev(v), G * 8, torch.int64).view(G, 8) for v in range(len(vals))]
            rows = self.tensor_from_ptr(r.rows_dev(), G, torch.int64)
            if not vals:
                st = []
            return kt, kn, st, rows
        finally:
            r.close()

    def merge(self, key_dtypes, kt, kn, st, val_is_int, aggs, null_alias=None):
        n = int(kt[0].shape[0]) if kt else 0
        bits = [pack_null_bits(f) for f in kn]
        cols = [self.col(dt, t, b, length=n) for dt, t, b in zip(key_dtypes, kt, bits)]
        sts = [s.contiguous() for s in st]
        self._sync()
        return self.ctx.groupby_merge(cols, [s.data_ptr() for s in sts], val_is_int, n, aggs)

    def groupby(self, keys, vals, aggs, filter=None):
        self._sync()
        return self.ctx.groupby_agg(keys, vals, aggs, filter=filter)

    def hash_partition(self, keys: Sequence[Column], nparts: int):
        n = keys[0].len
        perm = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        self._sync()
        counts = self.ctx.hash_partition(keys, nparts, perm.data_ptr())
        return perm[:n], counts

    def gather(self, col: Column, perm: torch.Tensor) -> torch.Tensor:
        out = torch.empty(perm.shape[0], dtype=_TORCH_DTYPE[col.dtype], device=self.device)
        self._sync()
        if perm.shape[0]:
            self.ctx.gather(col, perm.data_ptr(), n=perm.shape[0], idx_dev=True, out_dev=out.data_ptr())
        return out

    def null_flags(self, col: Column, perm: torch.Tensor) -> Optional[torch.Tensor]:
        """uint8 flag per permuted row: 1 where the source value is NULL (None when the column has no bitmap)."""
        if not col.nulls_ptr:
            return None
        nb = (col.len + 7) // 8
        bits = self.tensor_from_ptr(col.nulls_ptr, nb, torch.uint8)
        return ((bits[perm >> 3].to(torch.int32) >> (perm & 7).to(torch.int32)) & 1).to(torch.uint8)

    def join(self, left: Column, right: Column, how: int):
        self._sync()
        j = self.ctx.join_pairs(left, right, how)
        try:
            li = self.tensor_from_ptr(j.left_dev(), j.n, torch.int64)
            ri = self.tensor_from_ptr(j.right_dev(), j.n, torch.int64)
            return li, ri
        finally:
            j.close()


# ---------------------------------------------------------------- the distributed operators
class DistGroupBy:
    def __init__(self, ctx_or_backend, dist):
        self.b = ctx_or_backend if hasattr(ctx_or_backend, "partial") else CudaBackend(ctx_or_backend)
        self.dist = dist
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()

    def groupby_agg_lowcard(self, keys: Sequence[Column], vals: Sequence[Column], aggs, filter=None):
        """Local partial aggregate + all_gather of the (small) state rows + merge on every rank.
        Every rank returns the full result.  Keys, NULL flags and states travel as ONE packed int64 matrix
        [groups, nkeys * 2 + 8 * nvals]: one size exchange and one all_gather per call."""
        all_stats = any(op in (N.MIN, N.MAX, N.STD, N.VAR) for _, op in aggs)
        kt, kn, st, _rows = self.b.partial(keys, vals, filter, all_stats)
        G = int(kt[0].shape[0]) if kt else 0
        nk, nv = len(kt), len(st)
        packed = torch.empty((G, 2 * nk + 8 * nv), dtype=torch.int64, device=kt[0].device)
        for i, t in enumerate(kt):
            packed[:, i] = t.view(torch.int64) if t.dtype == torch.float64 else t.to(torch.int64)
        for i, t in enumerate(kn):
            packed[:, nk + i] = t.to(torch.int64)
        for i, s in enumerate(st):
            packed[:, 2 * nk + 8 * i: 2 * nk + 8 * (i + 1)] = s
        allp, _ = all_gather_varlen(self.dist, packed)
        gk = [(allp[:, i].contiguous().view(torch.float64) if t.dtype == torch.float64 else allp[:, i].to(t.dtype)).contiguous() for i, t in enumerate(kt)]
        gn = [allp[:, nk + i].to(torch.uint8).contiguous() for i in range(nk)]
        gs = [allp[:, 2 * nk + 8 * i: 2 * nk + 8 * (i + 1)].contiguous() for i in range(nv)]
        return self.b.merge([c.dtype for c in keys], gk, gn, gs, [c.dtype == N.I64 for c in vals], aggs)

    def groupby_agg_shuffle(self, keys: Sequence[Column], vals: Sequence[Column], aggs):
        """Hash-partition rows by key, one all_to_all per column, local aggregate.
        Rank r returns the groups whose key hashes to r (results stay sharded)."""
        perm, send = self.b.hash_partition(keys, self.world)
        recv = exchange_counts(self.dist, send, perm.device)

        def shuffle(col: Column) -> Column:
            data = all_to_all_rows(self.dist, self.b.gather(col, perm), send, recv)
            nf = self.b.null_flags(col, perm)
            nulls = None
            if nf is not None:
                nulls = pack_null_bits(all_to_all_rows(self.dist, nf, send, recv))
            return self.b.col(col.dtype, data, nulls, length=int(recv.sum()), null_alias=col.null_alias)
        return self.b.groupby([shuffle(c) for c in keys], [shuffle(c) for c in vals], aggs)


class DistJoin:
    def __init__(self, ctx_or_backend, dist):
        self.b = ctx_or_backend if hasattr(ctx_or_backend, "partial") else CudaBackend(ctx_or_backend)
        self.dist = dist
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()

    def join_pairs(self, left: Column, right: Column, how: int, left_row0: int, right_row0: int):
        """Both sides hash-partitioned by key and exchanged; local join; pairs in GLOBAL row numbers
        (row0 = global number of this rank's first row).  NULL keys never match and are dropped on both sides
        (join.rs:112,152), so they need not travel."""
        out = []
        for col, row0 in ((left, left_row0), (right, right_row0)):
            perm, send = self.b.hash_partition([col], self.world)
            recv = exchange_counts(self.dist, send, perm.device)
            data = all_to_all_rows(self.dist, self.b.gather(col, perm), send, recv)
            ids = all_to_all_rows(self.dist, perm + row0, send, recv)
            nf = self.b.null_flags(col, perm)
            nulls = None if nf is None else pack_null_bits(all_to_all_rows(self.dist, nf, send, recv))
            out.append((self.b.col(col.dtype, data, nulls, length=int(recv.sum())), ids))
        (lc, lids), (rc, rids) = out
        li, ri = self.b.join(lc, rc, how)
        gl = lids[li]
        gr = torch.where(ri >= 0, rids[ri.clamp(min=0)], torch.full_like(ri, -1)) if rids.numel() else torch.full_like(ri, -1)
        return gl, gr

    def join_pairs_auto(self, left: Column, right: Column, how: int, left_row0: int, right_row0: int):
        """The exchange join when it was set up (setup_fused) and no padded region overflows, the all_to_all path
        otherwise.  Always returns (global left rows, global right rows) tensors; -1 stands for None."""
        if getattr(self, "fused", False) and getattr(self, "x", None) is not None and how in (N.INNER, N.LEFT):
            j = self.join_pairs_fused(left, right, how, left_row0, right_row0)
            if j is not None:
                try:
                    return (self.b.tensor_from_ptr(j.left_dev(), j.n, torch.int64), self.b.tensor_from_ptr(j.right_dev(), j.n, torch.int64))
                finally:
                    j.close()
        return self.join_pairs(left, right, how, left_row0, right_row0)

    # ---- fused partition + shuffle over NVLink peer memory (pdrs_xjoin_*)
    def setup_fused(self, max_left_rows: int, max_right_rows: int, total_right_rows: int) -> bool:
        """Allocates this rank's receive area, exchanges the CUDA IPC handles (all_gather) and maps the peers'
        areas.  Collective; every rank passes the same numbers.  Returns False (on every rank) when any rank
        could not set it up - the caller then keeps using join_pairs."""
        from .core import XJoin
        ctx = self.b.ctx
        dev = self.b.device
        ok, handle = 1, bytes(64)
        try:
            self.x = XJoin(ctx, self.rank, self.world, max_left_rows, max_right_rows, total_right_rows)
            handle = self.x.ipc_handle()
        except Exception as e:  # noqa: BLE001
            ok, self.x, self.fused_error = 0, None, str(e)
        mine = torch.tensor(list(handle) + [ok], dtype=torch.uint8, device=dev)
        allh = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(allh, mine)
        allh = [bytes(t.cpu().tolist()) for t in allh]
        if ok and all(h[64] for h in allh) and self.world > 1:
            try:
                self.x.attach_ipc([h[:64] for h in allh])
            except Exception as e:  # noqa: BLE001
                ok, self.fused_error = 0, str(e)
        flag = torch.tensor([ok if all(h[64] for h in allh) else 0], dtype=torch.int32, device=dev)
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
        self.fused = bool(flag.item())
        if not self.fused and getattr(self, "x", None) is not None:
            self.x.close()
            self.x = None
        return self.fused

    def join_pairs_fused(self, left: Column, right: Column, how: int, left_row0: int, right_row0: int, timings: Optional[dict] = None):
        """The partition pass of the radix join IS the shuffle: every rank's partition kernel stores its (key, row)
        runs straight into the destination rank's receive area through NVLink; after one barrier the local build /
        probe kernels run on the received sub-buckets.  Returns a JoinResult holding this rank's pairs in GLOBAL
        row numbers, or None (on every rank) when a sub-bucket overflowed anywhere (skewed keys)."""
        dev = self.b.device
        self.b._sync()
        self.dist.barrier()                        # every rank has finished reading its receive area (previous call)
        ok = 1
        try:
            self.x.shuffle(left, right, right_row0)
            if timings is not None:
                timings["shuffle_ms"] = self.b.ctx.stats()["total_ms"]
        except Exception as e:  # noqa: BLE001
            ok, self.fused_error = 0, str(e)
        info = torch.tensor([ok, left_row0], dtype=torch.int64, device=dev)
        allinfo = [torch.empty_like(info) for _ in range(self.world)]
        self.dist.all_gather(allinfo, info)        # doubles as the barrier: all stores into my area are complete
        allinfo = [t.cpu().tolist() for t in allinfo]
        if not all(a[0] for a in allinfo):
            return None
        # local() can fail on ONE rank alone (a radix bucket of the staged layout overflowed, slot limit, OOM): the ranks
        # must agree on the outcome before anybody returns, or the next collective deadlocks
        j, ok = None, 1
        try:
            j = self.x.local(how, [a[1] for a in allinfo])
            if timings is not None:
                timings["local_ms"] = self.b.ctx.stats()["total_ms"]
        except Exception as e:  # noqa: BLE001
            ok, self.fused_error = 0, str(e)
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
        if not int(flag.item()):
            if j is not None:
                j.close()
            return None
        return j
