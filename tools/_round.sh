mkdir -p gpurun_out/r1n
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/r1n/gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r1n/gpu_tests.log; tail -25 gpurun_out/r1n/gpu_tests.log
timeout 200 python bench.py --workload yolo_tiny --batch 8 --res 416 --no-rooflines > gpurun_out/r1n/bench_yolo_b8.json 2> gpurun_out/r1n/bench_yolo.err; echo "rc=$?"; cut -c1-330 gpurun_out/r1n/bench_yolo_b8.json; grep -v "Yolo Avg" gpurun_out/r1n/bench_yolo.err | tail -3
