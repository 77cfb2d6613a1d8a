mkdir -p gpurun_out/r1h
BCNN_B200_BENCH_WATCHDOG_S=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r1h/bench_n2.json 2> gpurun_out/r1h/bench_n2.err
echo "bench rc=$?"; head -c 900 gpurun_out/r1h/bench_n2.json; tail -5 gpurun_out/r1h/bench_n2.err
