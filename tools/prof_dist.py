"""Where does the low-cardinality multi-GPU step spend its time?  One rank, collectives replaced by copies."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pandrs_b200 as pb
from pandrs_b200.dist import DistGroupBy, CudaBackend


class OneRank:
    @staticmethod
    def get_world_size(): return 1
    @staticmethod
    def get_rank(): return 0
    @staticmethod
    def all_gather(outs, t): outs[0].copy_(t)


torch.cuda.set_device(0)
stream = torch.cuda.current_stream()
ctx = pb.Context(device=0, stream=stream.cuda_stream)
n = 1_000_000_000
keys = ctx.synth_keys(n, card=1000)
vals = ctx.synth_vals(n, null_per_million=50_000)
aggs = [(0, op) for op in (pb.SUM, pb.MEAN, pb.MIN, pb.MAX, pb.COUNT, pb.STD)]
dgb = DistGroupBy(ctx, OneRank)
b = dgb.b


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def plain():
    r = ctx.groupby_agg([keys], [vals], aggs); r.close()


def dist():
    r = dgb.groupby_agg_lowcard([keys], [vals], aggs); r.close()


def partial_only():
    r = ctx.groupby_partial([keys], [vals], all_stats=True); r.close()


def partial_tensors():
    b.partial([keys], [vals], None, True)


print(f"groupby_agg            {timed(plain):7.3f} ms")
print(f"groupby_partial        {timed(partial_only):7.3f} ms")
print(f"backend.partial        {timed(partial_tensors):7.3f} ms   (+ D2D copies of keys / flags / states)")
print(f"dist lowcard (1 rank)  {timed(dist):7.3f} ms   (+ size exchange, all_gather, merge)")
