"""Per-call times of consecutive pdrs_join_pairs calls (configs[2]): wall clock and CUDA events, with and without torch in the
process - to see whether the mean over steps that bench.py reports is one stable number or a mix of fast and slow calls."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pandrs_b200 as pb

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
if len(sys.argv) > 2:
    import torch
    torch.cuda.init()
    x = torch.zeros(1, device="cuda")
ctx = pb.Context(0)
nb = n // 10
build = ctx.synth_join_keys(nb, unique=True)
probe = ctx.synth_join_keys(n, domain=2 * nb)
ctx.sync()
for how in (pb.INNER, pb.LEFT, pb.INNER):
    for i in range(8):
        t0 = time.perf_counter()
        ctx.timer_begin()
        j = ctx.join_pairs(probe, build, how)
        t1 = time.perf_counter()
        m = j.n
        j.close()
        ms = ctx.timer_end()
        t2 = time.perf_counter()
        print(f"how {how} call {i}: events {ms:8.3f} ms  wall join {1e3 * (t1 - t0):8.3f}  close {1e3 * (t2 - t1):8.3f}  pairs {m}  stats {ctx.stats()['total_ms']:.3f}", flush=True)
