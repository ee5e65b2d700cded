# 8-GPU evidence of the round: NCCL parity of the collective operators, then the four benchmark lines (weak, strong, join, configs[4])
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
free -g | head -2; nvidia-smi -L | head -8
timeout 600 $TR tests/dist_groupby_check.py > gpurun_out/dist_check_n${N}_r02b.log 2>&1; echo check_rc=$?; tail -16 gpurun_out/dist_check_n${N}_r02b.log
timeout 600 $TR bench.py --gpus $N --steps 30 --warmup 3 --e2e-steps 3 > gpurun_out/bench_n${N}_r02b.json 2> gpurun_out/bench_n${N}_r02b.err; echo rc=$?; tail -2 gpurun_out/bench_n${N}_r02b.err
timeout 300 $TR bench.py --gpus $N --steps 30 --warmup 3 --scaling strong --no-e2e --no-cpu > gpurun_out/bench_n${N}_strong_r02b.json 2> gpurun_out/bench_n${N}_strong_r02b.err; echo rc=$?; tail -2 gpurun_out/bench_n${N}_strong_r02b.err
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --metric join --no-cpu > gpurun_out/bench_join_n${N}_r02b.json 2> gpurun_out/bench_join_n${N}_r02b.err; echo rc=$?; tail -2 gpurun_out/bench_join_n${N}_r02b.err
timeout 500 $TR bench.py --gpus $N --steps 20 --warmup 3 --workload c5 --no-cpu > gpurun_out/bench_c5_n${N}_r02b.json 2> gpurun_out/bench_c5_n${N}_r02b.err; echo rc=$?; tail -2 gpurun_out/bench_c5_n${N}_r02b.err
for f in bench_n${N}_r02b bench_n${N}_strong_r02b bench_join_n${N}_r02b bench_c5_n${N}_r02b; do echo "== $f"; grep "^{" gpurun_out/$f.json | head -c 2600; echo; done
