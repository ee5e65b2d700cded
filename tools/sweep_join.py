import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pandrs_b200 as pb
ctx = pb.Context(0)
n = 400_000_000
build = ctx.synth_join_keys(n // 10, unique=True)
probe = ctx.synth_join_keys(n, domain=2 * (n // 10))
for lognb in (3, 4, 5, 6, 7, 8, 9):
    for cps in (2, 4, 8):
        ctx.set_option("join_log_nb", lognb); ctx.set_option("join_ctas_per_sm", cps)
        best = 1e9
        for _ in range(2):
            j = ctx.join_pairs(probe, build, pb.INNER); st = ctx.stats(); j.close()
            best = min(best, st["main_kernel_ms"])
        print(f"log_nb={lognb} ctas/sm={cps} probe_ms={best:.2f} total_ms={st['total_ms']:.2f}", flush=True)
