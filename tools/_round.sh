mkdir -p gpurun_out/r1m
timeout 200 python bench.py --workload yolo_tiny --batch 8 --res 416 --no-rooflines > gpurun_out/r1m/bench_yolo_b8.json 2> gpurun_out/r1m/bench_yolo.err; echo "rc=$?"; cut -c1-900 gpurun_out/r1m/bench_yolo_b8.json; grep -v "Yolo Avg" gpurun_out/r1m/bench_yolo.err | tail -5
timeout 100 python -m pytest tests -m gpu -x -q -k "graph_replay or yolo" 2>&1 | tail -3
