mkdir -p gpurun_out/r1p
timeout 300 python bench.py > gpurun_out/r1p/bench_n1.json 2> gpurun_out/r1p/bench_n1.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r1p/bench_n1.json; grep -o '"e2e": {[^}]*}' gpurun_out/r1p/bench_n1.json; grep -o '"gpu_launches": [0-9]*' gpurun_out/r1p/bench_n1.json; tail -3 gpurun_out/r1p/bench_n1.err
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
