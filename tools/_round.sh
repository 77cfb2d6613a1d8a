mkdir -p gpurun_out/r1q
timeout 300 python bench.py > gpurun_out/r1q/bench_n1.json 2> gpurun_out/r1q/bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r1q/bench_n1.json; grep -o '"e2e": {[^}]*}' gpurun_out/r1q/bench_n1.json; grep -o '"gpu_launches": [0-9]*' gpurun_out/r1q/bench_n1.json; tail -3 gpurun_out/r1q/bench_n1.err
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
