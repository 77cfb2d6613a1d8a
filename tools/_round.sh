mkdir -p gpurun_out/r1j
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r1j/gpu_tests.log 2>&1; tail -3 gpurun_out/r1j/gpu_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r1j/bench_n1.json 2> gpurun_out/r1j/bench_n1.err; head -c 300 gpurun_out/r1j/bench_n1.json
