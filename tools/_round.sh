mkdir -p gpurun_out/r1r
timeout 120 python -m pytest tests/test_dp.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-rooflines --no-cpu-baseline > gpurun_out/r1r/bench_n2.json 2> gpurun_out/r1r/bench_n2.err; echo "rc=$?"; cut -c1-200 gpurun_out/r1r/bench_n2.json; grep -o '"e2e": {[^}]*}' gpurun_out/r1r/bench_n2.json; tail -3 gpurun_out/r1r/bench_n2.err
