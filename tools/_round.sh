mkdir -p gpurun_out/r1f
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1f/gpu_tests.log 2>&1; tail -15 gpurun_out/r1f/gpu_tests.log
timeout 600 python bench.py > gpurun_out/r1f/bench_n1.json 2> gpurun_out/r1f/bench_n1.err; head -c 400 gpurun_out/r1f/bench_n1.json
