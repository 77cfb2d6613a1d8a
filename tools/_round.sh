mkdir -p gpurun_out/r1i
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r1i/gpu_tests.log 2>&1; tail -3 gpurun_out/r1i/gpu_tests.log
timeout 600 python tools/conv_sweep.py 256 5 > gpurun_out/r1i/conv_sweep_b256_pp.txt 2>&1
tail -n 3 gpurun_out/r1i/conv_sweep_b256_pp.txt
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r1i/bench_n1.json 2> gpurun_out/r1i/bench_n1.err; head -c 300 gpurun_out/r1i/bench_n1.json
