"""Host-clock phase marks (option trace = 1) of pdrs_groupby_rows at 1000 and 10 M groups: python tools/trace_rows.py [rows]"""
import sys
sys.path.insert(0, ".")
import pandrs_b200 as pb
ctx = pb.Context(0)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
for card in (1000, 10_000_000):
    keys = ctx.synth_keys(n, card=card)
    r = ctx.groupby_rows([keys]); r.close()
    ctx.set_option("trace", 1)
    print("card", card, file=sys.stderr, flush=True)
    r = ctx.groupby_rows([keys]); r.close()
    ctx.set_option("trace", 0)
    ctx.free(keys)
