mkdir -p gpurun_out/r1i
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k conv > gpurun_out/r1i/gpu_tests_conv.log 2>&1; tail -15 gpurun_out/r1i/gpu_tests_conv.log
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r1i/gpu_tests.log 2>&1; tail -5 gpurun_out/r1i/gpu_tests.log
timeout 300 python bench.py > gpurun_out/r1i/bench_n1.json 2> gpurun_out/r1i/bench_n1.err; head -c 400 gpurun_out/r1i/bench_n1.json; tail -3 gpurun_out/r1i/bench_n1.err
