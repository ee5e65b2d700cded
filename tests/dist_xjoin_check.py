"""Run under torchrun with >= 2 GPUs (launched by tests/test_gpu_multi.py): the fused partition + shuffle join
(pdrs_xjoin_*, CUDA IPC peer stores over NVLink) on world ranks against the oracle's join of the union of the rows."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as orc  # noqa: E402  (tests may use the oracle as the checker)
import pandrs_b200 as pb  # noqa: E402
from pandrs_b200.dist import DistJoin  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = pb.Context(device=local, stream=torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(77)                      # same data on every rank; each takes its slice
    nb, npr = 200_000, 1_500_000
    bk = rng.permutation(2 * nb)[:nb].astype(np.int64) * 1_000_003 - 5
    pk = (rng.integers(0, 2 * nb, npr) * 1_000_003 - 5).astype(np.int64)
    pnull = rng.random(npr) < 0.01
    lcut = [npr * r // world for r in range(world + 1)]
    rcut = [nb * r // world for r in range(world + 1)]
    dj = DistJoin(ctx, dist)
    cfg = None
    lc = ctx.upload(pb.Column.int64(pk[lcut[rank]:lcut[rank + 1]], pnull[lcut[rank]:lcut[rank + 1]]))
    rc = ctx.upload(pb.Column.int64(bk[rcut[rank]:rcut[rank + 1]]))
    # xjoin_mode 1 = fused (rank x radix bucket in one pass), 2 = staged (by rank, then the local radix partition), 0 = auto
    for mode, log_nb, how in ((1, 3, pb.INNER), (1, 3, pb.LEFT), (2, 3, pb.INNER), (2, 3, pb.LEFT), (0, 0, pb.INNER), (0, 0, pb.INNER)):
        if getattr(dj, "x", None) is None or (mode, log_nb) != cfg:     # the receive areas are reused call after call
            if getattr(dj, "x", None) is not None:
                dj.x.close()
            ctx.set_option("xjoin_mode", mode)
            ctx.set_option("join_log_nb", log_nb)
            assert dj.setup_fused(max(lcut[i + 1] - lcut[i] for i in range(world)), max(rcut[i + 1] - rcut[i] for i in range(world)), nb), getattr(dj, "fused_error", "")
            cfg = (mode, log_nb)
        j = dj.join_pairs_fused(lc, rc, how, lcut[rank], rcut[rank])
        assert j is not None, getattr(dj, "fused_error", "")
        li, ri = j.indices()
        j.close()
        mine = torch.tensor(np.stack([li, ri], 1), device="cuda")
        sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([mine.shape[0]], device="cuda"))
        m = max(int(s.item()) for s in sizes)
        pad = torch.zeros((m, 2), dtype=torch.int64, device="cuda")
        pad[: mine.shape[0]] = mine
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad)
        if rank == 0:
            got = torch.cat([b[: int(s.item())] for b, s in zip(bufs, sizes)]).cpu().numpy()
            wl, wr = orc.join(orc.Col(orc.I64, pk, orc.pack_bits(pnull)), orc.Col(orc.I64, bk), how)
            want = np.stack([wl, wr], 1)
            got = got[np.lexsort((got[:, 1], got[:, 0]))]
            want = want[np.lexsort((want[:, 1], want[:, 0]))]
            assert got.shape == want.shape and np.array_equal(got, want), (how, got.shape, want.shape)
    dj.x.close()
    ctx.close()
    dist.barrier()
    if rank == 0:
        print(f"xjoin parity ok on {world} GPUs")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
