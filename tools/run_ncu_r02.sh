# ncu evidence of round 2 (one GPU): launch list of the bench command, then full captures of the dominant kernels
set -x
B="python bench.py --steps 3 --warmup 3 --no-extras --no-e2e --no-cpu --no-parity"
$B > gpurun_out/plain_bench_r02.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_bench_r02.csv $B > gpurun_out/ncu_bench_r02.log 2>&1
P="python tools/prof_groupby.py --rows 200000000 --reps 1"
$P > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gb_tsort_kernel -c 1 -f -o gpurun_out/prof_ts_r02 $P > gpurun_out/ncu_ts_r02.log 2>&1
F="python tools/time_c5.py 200000000"
$F > gpurun_out/plain_c5_r02.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gb_few_kernel -c 1 -f -o gpurun_out/prof_few_r02 $F > gpurun_out/ncu_few_r02.log 2>&1
R="python tools/trace_rows.py 200000000"
$R > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rs_scatter_kernel -c 1 -f -o gpurun_out/prof_rs_r02 $R > gpurun_out/ncu_rs_r02.log 2>&1
for r in ts few rs; do python tools/ncu_summary.py gpurun_out/prof_${r}_r02.ncu-rep --top 25 > gpurun_out/ncu_${r}_r02_summary.txt 2>&1; done
ls -la gpurun_out/*.ncu-rep | tail -4
rm -f gpurun_out/prof_ts_r02.ncu-rep gpurun_out/prof_few_r02.ncu-rep gpurun_out/prof_rs_r02.ncu-rep      # (> 64 MiB together: only the summaries travel back)
