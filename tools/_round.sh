mkdir -p gpurun_out/r1t
nvidia-smi -L | head -4
for N in 2 4; do
BCNN_B200_BENCH_WATCHDOG_S=200 timeout 260 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r1t/bench_n$N.json 2> gpurun_out/r1t/bench_n$N.err
echo "N=$N rc=$?"; head -c 330 gpurun_out/r1t/bench_n$N.json; echo
done
