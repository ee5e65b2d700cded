"""Host-clock phase marks of the collective groupby (pdrs_groupby_agg_dist, option "trace") at the strong-scaling shard size:
torchrun --nproc-per-node N tools/trace_dist.py [rows per GPU].  Every mark synchronises the stream, so the phases add up to more
than an untraced step; the untraced step time is printed beside them."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import pandrs_b200 as pb

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 125_000_000
ctx = pb.Context(device=local, stream=torch.cuda.current_stream().cuda_stream)
comm = pb.Comm(ctx, rank, world, pb.torch_broadcast_id(dist, torch.device("cuda", local)) if world > 1 else None)
keys = ctx.synth_keys(n, card=1000, row0=rank * n)
vals = ctx.synth_vals(n, null_per_million=50_000, row0=rank * n)
aggs = [(0, op) for op in (pb.SUM, pb.MEAN, pb.MIN, pb.MAX, pb.COUNT, pb.STD)]


def step():
    r = comm.groupby_agg([keys], [vals], aggs, result_mode=pb.Comm.REPLICATED)
    r.close()


for _ in range(5):
    step()
comm.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    step()
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / 50 * 1e3
if rank == 0:
    print(f"untraced step: {ms:.3f} ms  (local kernel {ctx.stats()['main_kernel_ms']:.3f} ms)", flush=True)
    ctx.set_option("trace", 1)
comm.barrier()
for i in range(3):
    if rank == 0:
        print(f"-- traced step {i}", file=sys.stderr, flush=True)
    step()
ctx.set_option("trace", 0)
comm.barrier()
if world > 1:
    dist.destroy_process_group()
