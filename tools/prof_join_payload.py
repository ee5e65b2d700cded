"""Inner join 1B x 100M + gather of 2 build-side payload columns (BASELINE.json configs[2]) with the plain gather kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pandrs_b200 as pb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
ctx = pb.Context(0)
nb = n // 10
build = ctx.synth_join_keys(nb, unique=True)
probe = ctx.synth_join_keys(n, domain=2 * nb)
p1 = ctx.synth_keys(nb, card=1 << 40)          # i64 payload
p2 = ctx.synth_vals(nb)                        # f64 payload
for rep in range(2):
    ctx.timer_begin()
    j = ctx.join_pairs(probe, build, pb.INNER)
    t_join = ctx.timer_end()
    m = j.n
    o1, o2 = ctx.dev_alloc(m * 8), ctx.dev_alloc(m * 8)
    ctx.timer_begin()
    ctx.gather(p1, j.right_dev(), n=m, idx_dev=True, out_dev=o1)
    ctx.gather(p2, j.right_dev(), n=m, idx_dev=True, out_dev=o2)
    t_g = ctx.timer_end()
    print(f"pairs {m}: join {t_join:.2f} ms, 2 payload gathers {t_g:.2f} ms")
    ctx.dev_free(o1); ctx.dev_free(o2); j.close()
