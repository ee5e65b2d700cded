set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_multi_n2_r02.log 2>&1; echo multi_rc=$?; tail -30 gpurun_out/pytest_multi_n2_r02.log
timeout 600 $TR bench.py --gpus 2 --steps 30 --warmup 3 --dist-join > gpurun_out/bench_n2_r02.json 2> gpurun_out/bench_n2_r02.err; echo rc=$?; tail -3 gpurun_out/bench_n2_r02.err
timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 3 --scaling strong --no-e2e --no-cpu > gpurun_out/bench_n2_strong_r02.json 2> gpurun_out/bench_n2_strong_r02.err; echo rc=$?; tail -3 gpurun_out/bench_n2_strong_r02.err
timeout 400 $TR bench.py --gpus 2 --steps 10 --warmup 3 --metric join --no-cpu > gpurun_out/bench_join_n2_r02.json 2> gpurun_out/bench_join_n2_r02.err; echo rc=$?; tail -3 gpurun_out/bench_join_n2_r02.err
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 3 --workload c5 --rows 300000000 --no-cpu > gpurun_out/bench_c5_n2_r02.json 2> gpurun_out/bench_c5_n2_r02.err; echo rc=$?; tail -3 gpurun_out/bench_c5_n2_r02.err
for f in bench_n2_r02 bench_n2_strong_r02 bench_join_n2_r02 bench_c5_n2_r02; do echo "== $f"; head -c 2500 gpurun_out/$f.json; echo; done
