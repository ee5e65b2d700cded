"""configs[3] on one GPU: python tools/time_c4.py [rows] - Zipf(1.1) keys: dictionary key (1e4 groups), (i32, i64) (2.4e7 groups),
(i32, i64, dictionary) (1.65e8 groups), six aggregates of one f64 column; prints time, groups, algorithm, retries."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import pandrs_b200 as pb

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 500_000_000
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
ctx = pb.Context(0, stream=torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device=dev)
g.manual_seed(4242)
CH = 1 << 26


def zipf(domain, dtype, s=1.1):
    w = torch.arange(1, domain + 1, device=dev, dtype=torch.float64).pow(-s)
    cdf = (w.cumsum(0) / w.sum()).to(torch.float32)
    out = torch.empty(n, dtype=dtype, device=dev)
    for a in range(0, n, CH):
        b = min(n, a + CH)
        out[a:b] = torch.searchsorted(cdf, torch.rand(b - a, device=dev, generator=g)).clamp_(max=domain - 1).to(dtype)
    return out


k1, k2, k3 = zipf(1000, torch.int32), zipf(100_000, torch.int64), zipf(10_000, torch.int32)
v = torch.empty(n, dtype=torch.float64, device=dev)
for a in range(0, n, CH):
    b = min(n, a + CH)
    v[a:b] = torch.rand(b - a, device=dev, generator=g, dtype=torch.float64) * 1000
torch.cuda.synchronize()


def col(dtype, t):
    return pb.Column(dtype, device_ptr=t.data_ptr(), length=n, owner=t)


ALL6 = [(0, op) for op in (pb.SUM, pb.MEAN, pb.MIN, pb.MAX, pb.COUNT, pb.STD)]
ku = torch.randint(0, 24_000_000, (n,), device=dev, generator=g, dtype=torch.int64)
torch.cuda.synchronize()
for kv in filter(None, os.environ.get("PDRS_OPTS", "").split(",")):       # e.g. PDRS_OPTS=tsort_heavy=128 (experiments)
    ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
only = os.environ.get("C4_CASE")
for name, keys, bpr in (("uniform i64, 24M groups", [col(pb.I64, ku)], 16.0),("dictionary key", [col(pb.DICT_U32, k3)], 12.0), ("(i32, i64)", [col(pb.I32, k1), col(pb.I64, k2)], 20.0),
                        ("(i32, i64, dictionary)", [col(pb.I32, k1), col(pb.I64, k2), col(pb.DICT_U32, k3)], 24.0))[0 if len(sys.argv) < 4 else 1:]:
    if only and only not in name:
        continue
    for opt in ((0, 1), (2, 1)) if len(sys.argv) > 2 else ((0, 1),):
        ctx.set_option("part_hash", opt[0])
        best = 1e9
        for rep in range(3):
            ctx.set_option("timing", 2 if rep == 2 else 1)
            ctx.timer_begin()
            r = ctx.groupby_agg(keys, [col(pb.F64, v)], ALL6)
            G = r.n_groups
            r.close()
            best = min(best, ctx.timer_end())
        ctx.set_option("timing", 1)
        st = ctx.stats()
        print(f"{name:24s} part_hash={opt[0]}: {best:8.2f} ms  kernels {st['main_kernel_ms']:8.2f} ms  {G} groups  algo {st['groupby_algo_used']} retries {st['retries']} est {st['est_groups']}  = {bpr * n / best / 1e6 / 6504.1 * 100:.2f}% of roofline  spilled {st['spilled_rows']}", flush=True)
