"""Static evidence from the built library: per kernel, the SASS mnemonics that matter for this path (bulk async copies =
UBLKCP + mbarrier SYNCS, 128-bit global loads, shared / global atomics, cache-hinted loads, L2 prefetches, warp match /
vote) and the register count from the ptxas logs.  Usage: python tools/sass_summary.py > profiles/sass_summary_r01.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pandrs_b200", "lib", "libpandrs_b200.so")
PATTERNS = collections.OrderedDict([
    ("UBLKCP", r"\bUBLKCP"), ("SYNCS(mbar)", r"\bSYNCS"), ("LDG.128", r"\bLDG\.E\.(\w+\.)*128"), ("LDG.64", r"\bLDG\.E\.(\w+\.)*64"),
    ("STG", r"\bSTG\."), ("ATOMS", r"\bATOMS"), ("ATOMG", r"\bATOMG"), ("RED", r"\bRED\."), ("CCTL/PREF", r"\bCCTL"),
    ("VOTE", r"\bVOTE"), ("MATCH", r"\bMATCH"), ("SHFL", r"\bSHFL"), ("BAR", r"\bBAR\.SYNC"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"),
])
WANT = ("gb_tsort_kernel", "gp_part_kernel", "jpart1_kernel", "join_build_kernel", "join_probe_emit_kernel", "gb_finalize_kernel",
        "gb_key_range_kernel", "jx_publish_counts_kernel", "gb_sample_kernel",
        # round 2
        "gb_few_kernel", "gp_hash_kernel", "gh_kernel", "rs_scatter_kernel", "rs_hist_kernel", "gr_assign_kernel", "gr_table_build_kernel", "de_hash_kernel", "de_verify_kernel",
        "validity_to_nulls_kernel", "dist_merge_kernel", "dist_pack_kernel", "jfat_probe_emit_kernel", "jfat_build_kernel", "gr_median_keys_kernel")


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for name, pat in PATTERNS.items():
            if re.search(pat, line):
                funcs[cur][name] += 1
        if re.search(r"^\s*/\*[0-9a-f]{4,}\*/", line):
            funcs[cur]["inst"] += 1
    regs = {}
    for log in glob.glob(os.path.join(ROOT, "pandrs_b200", "csrc", "build", "*.ptxas.log")):
        txt = open(log).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'.*?Used (\d+) registers", txt, re.S):
            regs[m.group(1)] = int(m.group(2))
    names = demangle(list(funcs))
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(funcs)} kernels, arch sm_100a (cuobjdump -sass); counts are static instruction counts")
    print("# kernel (first variant per family with the highest instruction count) | regs | inst | " + " | ".join(PATTERNS))
    best = {}
    for f, c in funcs.items():
        d = names.get(f, f)
        fam = next((w for w in WANT if w in d), None)
        if fam is None:
            continue
        key = fam + ("<part>" if fam == "gb_tsort_kernel" and ", 2, " in d else "")
        if key not in best or c["inst"] > best[key][1]["inst"]:
            best[key] = (f, c, d)
    for key, (f, c, d) in sorted(best.items()):
        print(f"{key:28s} | {regs.get(f, '?'):>4} | {c['inst']:6d} | " + " | ".join(f"{c[p]:4d}" for p in PATTERNS))
        print(f"    {d[:200]}")
    print("# tile-sort kernel variants <threads, value type, stats (GB_SUM / GB_ALL), groups per thread, key mode (0 dense/hashed i64, 2 partitions, 3 generic, 4 raw 4-byte keys), plain, team>:")
    for f, c in funcs.items():
        d = names.get(f, f)
        if re.search(r"gb_tsort_kernel<512, double, \d, \d, [0234], true, false>", d):
            print(f"{d[d.index('gb_tsort_kernel'):][:60]:62s} | {regs.get(f, '?'):>4} | {c['inst']:6d} | " + " | ".join(f"{c[p]:4d}" for p in PATTERNS))
    fams = collections.Counter(next((w for w in WANT if w in names.get(f, f)), "other") for f in funcs)
    print("# kernels per family:", dict(fams))


if __name__ == "__main__":
    main()
