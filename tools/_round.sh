O=gpurun_out/r2zz; mkdir -p $O
timeout 900 python -m pytest tests/test_nhwc_bf16_gpu.py tests/test_kernels_gpu.py -x -q > $O/t.log 2>&1; tail -2 $O/t.log
timeout 300 python tools/resident_sweep.py 256 5 "" sdw > $O/sweep.txt 2>&1; tail -4 $O/sweep.txt
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; head -c 250 $O/bench_n1.json; echo
