"""Summarise an .ncu-rep: key raw metrics per kernel and the hottest source lines (needs -lineinfo + --import-source on).
usage: python tools/ncu_summary.py REPORT [--top N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:120])
    for k in KEYS:
        if k in hdr:
            print(f"  {k:90s} {r[hdr.index(k)]:>16s} {rows[1][hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fname, hdr, data, kern = "", None, [], ""
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) >= 2 and r[0] == "Function Name":
        kern = r[1]
    elif len(r) > 4 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] != "":
        try:
            d = dict(zip(hdr[4:], r[4:]))
            stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0")}
            data.append((int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0), int(d.get("L1 Wavefronts Shared") or 0), fname, r[0], r[1], stalls, kern))
        except ValueError:
            pass
tot = sum(d[0] for d in data) or 1
print(f"== hottest source lines (of {tot} samples; inst = warp instructions, wf = shared wavefronts)")
for d in sorted(data, key=lambda d: -d[0])[:top]:
    st = " ".join(f"{k}:{v}" for k, v in sorted(d[6].items(), key=lambda kv: -kv[1])[:3])
    print(f"  {100.0 * d[0] / tot:5.1f}%  inst {d[1]:>10d} wf {d[2]:>10d}  {d[3]}:{d[4]:>4s}  {d[5].strip()[:100]}   [{st}]")
