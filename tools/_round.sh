O=gpurun_out/r2zp; mkdir -p $O
timeout 900 python -m pytest tests/test_nhwc_bf16_gpu.py tests/test_kernels_gpu.py tests/test_nets_gpu.py -x -q > $O/t.log 2>&1; tail -3 $O/t.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; head -c 250 $O/bench_n1.json; echo; tail -2 $O/bench_n1.err
