mkdir -p gpurun_out/r1o
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r1o/gpu_tests.log 2>&1; tail -4 gpurun_out/r1o/gpu_tests.log
timeout 200 python tools/dw_sweep.py 64 > gpurun_out/r1o/dw_sweep_b64.txt 2>&1; cat gpurun_out/r1o/dw_sweep_b64.txt
timeout 200 python bench.py --workload mobilenet --batch 64 --res 224 --no-rooflines --no-cpu-baseline > gpurun_out/r1o/bench_mobilenet.json 2>> gpurun_out/r1o/bench.err; cut -c1-200 gpurun_out/r1o/bench_mobilenet.json
