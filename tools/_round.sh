mkdir -p gpurun_out/r1r
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r1r/gpu_tests.log 2>&1; tail -3 gpurun_out/r1r/gpu_tests.log
timeout 400 python bench.py > gpurun_out/r1r/bench_n1.json 2> gpurun_out/r1r/bench_n1.err; head -c 250 gpurun_out/r1r/bench_n1.json; tail -2 gpurun_out/r1r/bench_n1.err
