# 2-GPU checks of the exchange join after the build-once / probe-in-rounds change
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -q -x -k "xjoin or multi" 2>&1 | tail -4
timeout 600 $TR tests/dist_groupby_check.py > gpurun_out/dist_check_n2_r02c.log 2>&1; echo check_rc=$?; tail -8 gpurun_out/dist_check_n2_r02c.log
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --metric join --no-cpu > gpurun_out/bench_join_n2_r02c.json 2> gpurun_out/bench_join_n2_r02c.err; echo rc=$?; tail -3 gpurun_out/bench_join_n2_r02c.err
PDRS_OPTS=xjoin_round_rows=536870912 timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --metric join --no-cpu > gpurun_out/bench_join_n2_2rounds_r02c.json 2> gpurun_out/bench_join_n2_2rounds_r02c.err; echo rc=$?; tail -3 gpurun_out/bench_join_n2_2rounds_r02c.err
