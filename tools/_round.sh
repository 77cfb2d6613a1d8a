O=gpurun_out/r2zv; mkdir -p $O
timeout 600 python -m pytest tests/test_dp.py -m gpu -x -q > $O/dp_tests.log 2>&1; tail -2 $O/dp_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; head -c 300 $O/bench_n2.json; echo
