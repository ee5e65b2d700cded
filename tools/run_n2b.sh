# 2-GPU checks of the second half of round 2: NCCL parity (with the join in rounds), strong-scaling phase trace, join + strong lines
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $TR tests/dist_groupby_check.py > gpurun_out/dist_check_n2_r02b.log 2>&1; echo check_rc=$?; tail -18 gpurun_out/dist_check_n2_r02b.log
timeout 300 $TR tools/trace_dist.py 125000000 > gpurun_out/trace_dist_n2.log 2>&1; echo trace_rc=$?; grep -v "^\*\|OMP_NUM" gpurun_out/trace_dist_n2.log | tail -40
timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 3 --scaling strong --rows 250000000 --no-e2e --no-cpu > gpurun_out/bench_n2_strong250_r02b.json 2> gpurun_out/bench_n2_strong250_r02b.err; echo rc=$?; tail -3 gpurun_out/bench_n2_strong250_r02b.err
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --metric join --no-cpu > gpurun_out/bench_join_n2_r02b.json 2> gpurun_out/bench_join_n2_r02b.err; echo rc=$?; tail -3 gpurun_out/bench_join_n2_r02b.err
