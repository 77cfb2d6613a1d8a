mkdir -p gpurun_out/r1n
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r1n/gpu_tests.log 2>&1; tail -4 gpurun_out/r1n/gpu_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-rooflines > gpurun_out/r1n/bench_n1.json 2> gpurun_out/r1n/bench_n1.err; head -c 300 gpurun_out/r1n/bench_n1.json; tail -3 gpurun_out/r1n/bench_n1.err
: > gpurun_out/r1n/bench_other_configs.jsonl
for cfg in "mnist 64 28" "cifar 128 32"; do
  set -- $cfg
  timeout 200 python bench.py --workload $1 --batch $2 --res $3 --no-rooflines --no-cpu-baseline >> gpurun_out/r1n/bench_other_configs.jsonl 2>> gpurun_out/r1n/bench_other.err
done
cut -c1-200 gpurun_out/r1n/bench_other_configs.jsonl
