mkdir -p gpurun_out/r1u
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r1u/gpu_tests.log 2>&1; tail -6 gpurun_out/r1u/gpu_tests.log
timeout 200 python bench.py --workload yolo_tiny --batch 8 --res 416 --no-rooflines --no-cpu-baseline > gpurun_out/r1u/bench_yolo.json 2>> gpurun_out/r1u/bench.err; cut -c1-200 gpurun_out/r1u/bench_yolo.json; tail -3 gpurun_out/r1u/bench.err
