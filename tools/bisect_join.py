import os, sys, time
sys.path.insert(0, ".")
mode = sys.argv[1]
import torch
import pandrs_b200 as pb
torch.cuda.set_device(0)
stream = torch.cuda.current_stream()
ctx = pb.Context(device=0, stream=stream.cuda_stream)
if "json" in mode:
    import json
    peaks = json.load(open("MEASURED_PEAKS.json"))
nb = 100_000_000
if "row0" in mode:
    build = ctx.synth_join_keys(nb, unique=True, row0=0); probe = ctx.synth_join_keys(10 * nb, domain=2 * nb * 1, row0=0)
else:
    build = ctx.synth_join_keys(nb, unique=True); probe = ctx.synth_join_keys(10 * nb, domain=2 * nb)
ctx.sync()
def step():
    j = ctx.join_pairs(probe, build, pb.INNER); m = j.n; j.close()
for _ in range(3): step()
if "sync" in mode: torch.cuda.synchronize()
if "event" in mode:
    e0 = torch.cuda.Event(enable_timing=True); e0.record(stream)
walls = []
for i in range(8):
    t0 = time.perf_counter(); step()
    if "stats" in mode: ctx.stats()
    walls.append((time.perf_counter() - t0) * 1e3)
print(mode, " ".join("%.1f" % w for w in walls), flush=True)
